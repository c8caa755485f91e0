// cuda_stub.h - host stand-in for the subset of the CUDA runtime API that pragma_b200.cu and
// fft_launch.cuh use (TEST BUILDS ONLY, -DPDSP_EMU).  "Device" memory is host memory, streams
// and events are synchronous no-ops, kernel launches run under the SIMT emulator.  This lets the
// C-ABI host logic (chunking, staging slots, plan cache, dispatch, locking) be exercised on a
// machine without a GPU.  It is not part of the product and is never on a product path.
#pragma once
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef enum {
  cudaSuccess = 0,
  cudaErrorInvalidValue = 1,
  cudaErrorInvalidDevice = 101,
  cudaErrorNotReady = 600,
  cudaErrorPeerAccessAlreadyEnabled = 704,
  cudaErrorLaunchOutOfResources = 701
} cudaError_t;
typedef struct pdsp_stub_stream* cudaStream_t;
typedef struct pdsp_stub_event* cudaEvent_t;
struct cudaDeviceProp {
  int major, minor, multiProcessorCount;
};
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes {
  cudaMemoryType type;
};
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostAllocMapped = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaGetLastError();
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaMalloc(void** p, size_t bytes);
cudaError_t cudaFree(void* p);
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned flags);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned flags);
cudaError_t cudaMemset(void* p, int value, size_t bytes);
inline cudaError_t cudaMemsetAsync(void* p, int value, size_t bytes, cudaStream_t) { return cudaMemset(p, value, bytes); }
cudaError_t cudaStreamQuery(cudaStream_t s);
cudaError_t cudaEventQuery(cudaEvent_t e);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaDeviceCanAccessPeer(int* can, int dev, int peer);
cudaError_t cudaDeviceEnablePeerAccess(int peer, unsigned flags);
cudaError_t cudaMemcpyPeerAsync(void* dst, int dst_dev, const void* src, int src_dev, size_t bytes, cudaStream_t s);
cudaError_t cudaDeviceGetPCIBusId(char* id, int len, int dev);
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind,
                              cudaStream_t s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p);
// stub-only: counters the tests read to see which copy path (direct DMA vs staged) was taken
extern "C" long long pdsp_stub_counter(int which);  // 0: H2D copies, 1: D2H copies, 2: launches, 3: live device allocations

template <typename K>
cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) {
  return cudaSuccess;
}
template <typename K>
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int threads, size_t) {
  *n = threads >= 512 ? 1 : 2;
  return cudaSuccess;
}
