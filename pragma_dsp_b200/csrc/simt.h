// simt.h - the handful of SIMT primitives the FFT kernels use, behind one namespace.
//
// Product build (nvcc, sm_100a): thin wrappers over the CUDA intrinsics / PTX.
// Test build (g++ -DPDSP_EMU, tests/simt_emu): the same kernel source runs on a host SIMT
// emulator - one OS thread per CUDA thread, barriers for bar.sync/__syncwarp, a mailbox for
// shuffles - so index math and synchronisation can be verified against the oracle in the
// build container, which has no GPU.  The emulator is test infrastructure: it is never
// compiled into libpragma_b200.so and the product has no CPU execution path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(PDSP_EMU)

#define PDSP_DEVICE __device__ __forceinline__
#define PDSP_DEVICE_NOINLINE static __device__ __noinline__
#define PDSP_GLOBAL __global__
#define PDSP_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#define PDSP_KERNEL_LIMITS(threads, regs) __launch_bounds__(threads) __maxnreg__(regs)
#define PDSP_RESTRICT __restrict__
#define PDSP_UNROLL _Pragma("unroll")
#define PDSP_LAUNCH(kernel, grid, threads, smem, stream, ...) kernel<<<(grid), (threads), (smem), (stream)>>>(__VA_ARGS__)

namespace simt {
PDSP_DEVICE int tid() { return (int)threadIdx.x; }
PDSP_DEVICE int bid() { return (int)blockIdx.x; }
PDSP_DEVICE int nblocks() { return (int)gridDim.x; }
PDSP_DEVICE int nthreads() { return (int)blockDim.x; }
PDSP_DEVICE void sync_warp() { __syncwarp(); }
PDSP_DEVICE void sync_block() { __syncthreads(); }
// Named barrier over `nthreads` (multiple of 32) threads; ids 1..15 (0 is __syncthreads).
PDSP_DEVICE void sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <typename T>
PDSP_DEVICE T shfl(T v, int src_lane, int width) {
  return __shfl_sync(0xffffffffu, v, src_lane, width);
}
template <typename T>
PDSP_DEVICE T shfl_xor(T v, int mask, int width) {
  return __shfl_xor_sync(0xffffffffu, v, mask, width);
}
PDSP_DEVICE bool any(bool pred) { return __any_sync(0xffffffffu, pred) != 0; }
PDSP_DEVICE unsigned char* smem() {
  extern __shared__ __align__(16) unsigned char pdsp_smem_[];
  return pdsp_smem_;
}
template <typename T>
PDSP_DEVICE T ldg(const T* p) {
  return __ldg(p);
}
}  // namespace simt

#else  // ------------------------------------------------------------------ host emulation

#include <math.h>
#include <string.h>

#include <functional>

#define PDSP_DEVICE inline
#define PDSP_DEVICE_NOINLINE static inline
#define PDSP_GLOBAL
#define PDSP_LAUNCH_BOUNDS(t, b)
#define PDSP_KERNEL_LIMITS(threads, regs)
#define PDSP_RESTRICT __restrict__
#define PDSP_UNROLL
#define PDSP_LAUNCH(kernel, grid, threads, smem, stream, ...) \
  simt::emu_launch((grid), (threads), (smem), [=] { kernel(__VA_ARGS__); })

namespace simt {
// runs `body` once per emulated CUDA thread, CTA after CTA (tests/simt_emu/emu_runtime.cc)
void emu_launch(int grid, int threads, size_t smem_bytes, const std::function<void()>& body);
struct EmuThread {
  int tid, bid, nblocks, nthreads;
  unsigned char* smem;
  void* block;  // EmuBlock*
};
extern thread_local EmuThread emu_self;
void emu_sync_warp();
void emu_sync_block();
void emu_sync_named(int id, int nthreads);
void emu_shfl(const void* in, void* out, int bytes, int src_lane, int width, bool is_xor);

inline int tid() { return emu_self.tid; }
inline int bid() { return emu_self.bid; }
inline int nblocks() { return emu_self.nblocks; }
inline int nthreads() { return emu_self.nthreads; }
inline void sync_warp() { emu_sync_warp(); }
inline void sync_block() { emu_sync_block(); }
inline void sync_named(int id, int nthreads) { emu_sync_named(id, nthreads); }
template <typename T>
inline T shfl(T v, int src_lane, int width) {
  T r;
  emu_shfl(&v, &r, (int)sizeof(T), src_lane, width, false);
  return r;
}
template <typename T>
inline T shfl_xor(T v, int mask, int width) {
  T r;
  emu_shfl(&v, &r, (int)sizeof(T), mask, width, true);
  return r;
}
inline bool any(bool pred) {  // warp vote through the shuffle mailbox: OR over the warp's lanes
  int acc = pred ? 1 : 0;
  for (int m = 16; m >= 1; m >>= 1) acc |= shfl_xor(acc, m, 32);
  return acc != 0;
}
inline unsigned char* smem() { return emu_self.smem; }
template <typename T>
inline T ldg(const T* p) {
  return *p;
}
}  // namespace simt

// CUDA math / bit intrinsics used by the kernels
inline int __double2hiint(double d) {
  int64_t b;
  memcpy(&b, &d, 8);
  return (int)(b >> 32);
}
inline int __float_as_int(float f) {
  int b;
  memcpy(&b, &f, 4);
  return b;
}
inline long long __double_as_longlong(double d) {
  long long b;
  memcpy(&b, &d, 8);
  return b;
}
#endif
