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
#include <cuda.h>  // CUtensorMap (type only; the driver entry point is resolved at run time)

#define PDSP_DEVICE __device__ __forceinline__
#define PDSP_DEVICE_NOINLINE static __device__ __noinline__
#define PDSP_GLOBAL __global__
#define PDSP_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#define PDSP_KERNEL_LIMITS(threads, regs) __launch_bounds__(threads) __maxnreg__(regs)
#define PDSP_RESTRICT __restrict__
#define PDSP_UNROLL _Pragma("unroll")
#define PDSP_LAUNCH(kernel, grid, threads, smem, stream, ...) kernel<<<(grid), (threads), (smem), (stream)>>>(__VA_ARGS__)
#define PDSP_GRID_CONSTANT __grid_constant__

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

// ---- TMA (cp.async.bulk.tensor) + mbarrier, the sm_90+/sm_100 bulk-copy path (SASS: UTMALDG, SYNCS)
typedef CUtensorMap TensorMap2D;  // encoded on the host with cuTensorMapEncodeTiled, passed __grid_constant__
PDSP_DEVICE unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
PDSP_DEVICE void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
PDSP_DEVICE void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 2-D tile load: box (tensor map's boxDim) at element coordinates (x = inner, y = outer) -> dense smem
PDSP_DEVICE void tma_load_2d(void* smem_dst, const TensorMap2D* map, int x, int y, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
// 1-D bulk copy global -> shared (cp.async.bulk, SASS UBLKCP): `bytes` and both addresses multiples of 16; completion
// is counted on the mbarrier like a tensor load's
PDSP_DEVICE void bulk_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// rank-3 forms (second-generation large-FFT passes): tile load, and tile STORE shared -> global (SASS UTMASTG) as a bulk
// async-group operation: commit after issuing, wait_read before the source shared memory is reused, wait_all before exit
typedef CUtensorMap TensorMap;
PDSP_DEVICE void tma_load_3d(void* smem_dst, const TensorMap* map, int x, int y, int z, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}
PDSP_DEVICE void tma_store_3d(const TensorMap* map, int x, int y, int z, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(x), "r"(y), "r"(z),
               "r"(smem_u32(smem_src))
               : "memory");
}
PDSP_DEVICE void tma_prefetch_3d(const TensorMap* map, int x, int y, int z) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z) : "memory");
}
PDSP_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
PDSP_DEVICE void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
PDSP_DEVICE void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// L2 prefetch of the same box (no shared-memory destination, no barrier)
PDSP_DEVICE void tma_prefetch_2d(const TensorMap2D* map, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
}
PDSP_DEVICE void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// order generic-proxy accesses to shared memory before later async-proxy (TMA) accesses to the same bytes
PDSP_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// bulk prefetch of [p, p + bytes) into L2 (one instruction; p and bytes multiples of 16)
PDSP_DEVICE void prefetch_l2_bulk(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// per-line prefetch hints (experiment PDSP_NEXT_PREFETCH in the framed kernels): one 128-byte line per lane
PDSP_DEVICE void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
PDSP_DEVICE void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// completion doorbell (low-latency host entry points): the last CTA to finish writes a sequence number into
// host-mapped pinned memory; the host spins on it instead of going through a stream synchronisation
PDSP_DEVICE void fence_system() { __threadfence_system(); }
PDSP_DEVICE unsigned atomic_add(unsigned* p, unsigned v) { return atomicAdd(p, v); }
// inter-CTA hand-over inside one launch (fused large-FFT passes): release = bulk stores complete, proxy fence, gpu fence,
// then the counter increment; acquire = ld.acquire.gpu on the counter
PDSP_DEVICE void fence_gpu() { __threadfence(); }
PDSP_DEVICE void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
PDSP_DEVICE unsigned load_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
PDSP_DEVICE void store_volatile(unsigned* p, unsigned v) { *reinterpret_cast<volatile unsigned*>(p) = v; }
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
#define PDSP_GRID_CONSTANT
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
// ---- emulated TMA: the "tensor map" is a plain description; a load copies the box synchronously and
// completes the barrier's transaction count; waits spin on the phase bit like the hardware's try_wait.
struct TensorMap2D {
  const void* base;
  long long dim0, dim1;       // extent in elements (inner, outer)
  long long stride1_bytes;    // byte stride between consecutive outer indices
  int box0, box1, esize;
};
void emu_mbar_init(unsigned long long* bar);
void emu_mbar_expect_tx(unsigned long long* bar, unsigned bytes);
void emu_mbar_complete_tx(unsigned long long* bar, unsigned bytes);
void emu_mbar_wait(unsigned long long* bar, unsigned parity);
inline void mbar_init(unsigned long long* bar, int) { emu_mbar_init(bar); }
inline void mbar_expect_tx(unsigned long long* bar, unsigned bytes) { emu_mbar_expect_tx(bar, bytes); }
inline void tma_load_2d(void* smem_dst, const TensorMap2D* m, int x, int y, unsigned long long* bar) {
  char* dst = static_cast<char*>(smem_dst);
  for (int r = 0; r < m->box1; ++r) {
    for (int c = 0; c < m->box0; ++c) {
      const long long gx = x + c, gy = y + r;
      char* d = dst + ((size_t)r * m->box0 + c) * m->esize;
      if (gx < m->dim0 && gy < m->dim1)
        memcpy(d, static_cast<const char*>(m->base) + gy * m->stride1_bytes + gx * m->esize, (size_t)m->esize);
      else
        memset(d, 0, (size_t)m->esize);  // out-of-bounds elements read as zero
    }
  }
  emu_mbar_complete_tx(bar, (unsigned)(m->box0 * m->box1 * m->esize));
}
inline void bulk_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  memcpy(smem_dst, gsrc, bytes);
  emu_mbar_complete_tx(bar, bytes);
}
inline void tma_prefetch_2d(const TensorMap2D*, int, int) {}
// rank-3 description: element (x, y, z) at base + x*esize + y*stride_bytes[0] + z*stride_bytes[1]
struct TensorMap {
  void* base;
  long long dim[3];
  long long stride_bytes[2];
  int box[3], esize;
};
inline void tma_load_3d(void* smem_dst, const TensorMap* m, int x, int y, int z, unsigned long long* bar) {
  char* dst = static_cast<char*>(smem_dst);
  size_t i = 0;
  for (int c = 0; c < m->box[2]; ++c)
    for (int b = 0; b < m->box[1]; ++b)
      for (int a = 0; a < m->box[0]; ++a, ++i) {
        const long long gx = x + a, gy = y + b, gz = z + c;
        char* d = dst + i * (size_t)m->esize;
        if (gx < m->dim[0] && gy < m->dim[1] && gz < m->dim[2])
          memcpy(d, static_cast<const char*>(m->base) + gx * m->esize + gy * m->stride_bytes[0] + gz * m->stride_bytes[1], (size_t)m->esize);
        else
          memset(d, 0, (size_t)m->esize);
      }
  emu_mbar_complete_tx(bar, (unsigned)(m->box[0] * m->box[1] * m->box[2] * m->esize));
}
inline void tma_store_3d(const TensorMap* m, int x, int y, int z, const void* smem_src) {
  const char* src = static_cast<const char*>(smem_src);
  size_t i = 0;
  for (int c = 0; c < m->box[2]; ++c)
    for (int b = 0; b < m->box[1]; ++b)
      for (int a = 0; a < m->box[0]; ++a, ++i) {
        const long long gx = x + a, gy = y + b, gz = z + c;
        if (gx < m->dim[0] && gy < m->dim[1] && gz < m->dim[2])  // out-of-bounds elements are not written
          memcpy(static_cast<char*>(m->base) + gx * m->esize + gy * m->stride_bytes[0] + gz * m->stride_bytes[1], src + i * (size_t)m->esize,
                 (size_t)m->esize);
      }
}
inline void tma_prefetch_3d(const TensorMap*, int, int, int) {}
inline void bulk_commit() {}
inline void bulk_wait_read() {}
inline void bulk_wait_all() {}
inline void mbar_wait(unsigned long long* bar, unsigned parity) { emu_mbar_wait(bar, parity); }
inline void fence_proxy_async() {}
inline void prefetch_l2_bulk(const void*, unsigned) {}
inline void prefetch_l1(const void*) {}
inline void prefetch_l2(const void*) {}
inline bool any(bool pred) {  // warp vote through the shuffle mailbox: OR over the warp's lanes
  int acc = pred ? 1 : 0;
  for (int m = 16; m >= 1; m >>= 1) acc |= shfl_xor(acc, m, 32);
  return acc != 0;
}
inline void fence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline unsigned atomic_add(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline void fence_gpu() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void fence_proxy_async_all() {}
inline unsigned load_acquire(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void store_volatile(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
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
