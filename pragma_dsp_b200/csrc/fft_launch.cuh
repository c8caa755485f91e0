// fft_launch.cuh - per-size kernel configuration and persistent-grid launchers.
#pragma once
#ifdef PDSP_EMU
#include "cuda_stub.h"  // tests/simt_emu: host stand-in for the runtime API (test builds only)
#else
#include <cuda_runtime.h>
#endif

#include "fft_config.h"
#include "fft_kernels.cuh"

namespace pdsp {

struct LaunchCtx {
  int device;
  int sm_count;
  cudaStream_t stream;
  // returns the device table of per-pass twiddles for an M = 2^log2m point schedule with full-pass
  // radix 2^rb (FftEngine::tw_offset layout), building and caching it on first use; null on failure
  const void* (*pass_twiddles)(void* owner, bool f64, int log2m, int rb);
  void* owner;
};

constexpr int kMaxDevices = 16;

template <typename KernelT>
inline cudaError_t persistent_grid(KernelT kern, int threads, size_t smem, const LaunchCtx& lc, int* bps_cache,
                                   long long groups, int* grid_out) {
  if (lc.device < 0 || lc.device >= kMaxDevices) return cudaErrorInvalidDevice;
  if (bps_cache[lc.device] == 0) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    int bps = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, threads, smem);
    if (e != cudaSuccess) return e;
    if (bps < 1) return cudaErrorLaunchOutOfResources;
    bps_cache[lc.device] = bps;
  }
  long long cap = (long long)bps_cache[lc.device] * lc.sm_count;
  *grid_out = (int)(groups < cap ? groups : cap);
  return cudaSuccess;
}

template <typename T, int LOG2M, int MODE, int VAR = 0>
cudaError_t launch_r2c_t(const R2CParams& p, const LaunchCtx& lc) {
  using C = KCfg<T, LOG2M, VAR>;
  using E = FftEngine<T, LOG2M, C::LOG2P, C::MAXRB>;
  constexpr int THREADS = C::THREADS;
  constexpr int SLOTS = THREADS / E::TF;
  constexpr size_t SMEM = r2c_smem_bytes<T, E, MODE, SLOTS>();
  static_assert((MODE & MD_STAGED) == 0 || E::TF <= 32, "staged loads: frames no wider than a warp");
  auto kern = [] {
    if constexpr (C::MAXREG > 0)
      return r2c_kernel_mr<T, LOG2M, C::LOG2P, C::MAXRB, THREADS, C::MAXREG, MODE>;
    else
      return r2c_kernel<T, LOG2M, C::LOG2P, C::MAXRB, THREADS, C::MINB, MODE>;
  }();
  static int bps[kMaxDevices] = {0};
  if (p.batch <= 0) return cudaSuccess;
  int grid = 0;
  cudaError_t e = persistent_grid(kern, THREADS, SMEM, lc, bps, (p.batch + SLOTS - 1) / SLOTS, &grid);
  if (e != cudaSuccess) return e;
  R2CParams q = p;
  q.tw = E::TW_ELEMS ? lc.pass_twiddles(lc.owner, sizeof(T) == 8, LOG2M, E::RB) : nullptr;
  if (E::TW_ELEMS && !q.tw) return cudaErrorInvalidValue;
  PDSP_LAUNCH(kern, grid, THREADS, SMEM, lc.stream, q);
  return cudaGetLastError();
}

template <typename T, int LOG2M>
cudaError_t launch_c2c_t(const C2CParams& p, const LaunchCtx& lc) {
  using C = KCfg<T, LOG2M>;
  using E = FftEngine<T, LOG2M, C::LOG2P, C::MAXRB>;
  constexpr int THREADS = C::THREADS;
  constexpr int SLOTS = THREADS / E::TF;
  constexpr size_t SMEM = E::NEEDS_SMEM ? sizeof(cx<T>) * E::SMEM_ELEMS * SLOTS : 0;
  auto kern = c2c_kernel<T, LOG2M, C::LOG2P, C::MAXRB, THREADS, C::MINB>;
  static int bps[kMaxDevices] = {0};
  if (p.batch <= 0) return cudaSuccess;
  int grid = 0;
  cudaError_t e = persistent_grid(kern, THREADS, SMEM, lc, bps, (p.batch + SLOTS - 1) / SLOTS, &grid);
  if (e != cudaSuccess) return e;
  C2CParams q = p;
  q.tw = E::TW_ELEMS ? lc.pass_twiddles(lc.owner, sizeof(T) == 8, LOG2M, E::RB) : nullptr;
  if (E::TW_ELEMS && !q.tw) return cudaErrorInvalidValue;
  PDSP_LAUNCH(kern, grid, THREADS, SMEM, lc.stream, q);
  return cudaGetLastError();
}

// The output combinations that get a compile-time specialised kernel (everything else, and every
// ragged/unaligned call, runs the generic kernel).
#define PDSP_SPEC_MODES(X)                                                                          \
  X(MD_AMP) X(MD_AMP | MD_PEAK) X(MD_PEAK) X(MD_CPLX) X(MD_AMP | MD_PHASE) X(MD_AMP | MD_PHASE | MD_PEAK)                    \
  X(MD_AMP | MD_TWO) X(MD_AMP | MD_PHASE | MD_PEAK | MD_TWO)                                                               \
  X(MD_AMP | MD_PAD) X(MD_AMP | MD_PEAK | MD_PAD) X(MD_AMP | MD_PHASE | MD_PEAK | MD_PAD)

// ... and those that also exist with bulk-staged sample loads (MD_STAGED), for the sizes staged_supported() names
#define PDSP_STAGED_MODES(X) \
  X(MD_AMP) X(MD_AMP | MD_PEAK) X(MD_PEAK) X(MD_CPLX) X(MD_AMP | MD_PHASE) X(MD_AMP | MD_PHASE | MD_PEAK)

// staged kernels are an opt-in experiment (pdsp_ctx_tune "staged"), built for the headline size only
template <typename T, int LOG2M>
constexpr bool staged_supported() {
  using C = KCfg<T, LOG2M>;
  return LOG2M == kVariantLog2M && C::TF <= 32;
}

inline bool mode_is_specialised(int mode) {
  switch (mode) {
#define X(m) \
  case (m):  \
    return true;
    PDSP_SPEC_MODES(X)
#undef X
    default:
      return false;
  }
}

// One translation unit instantiates a contiguous range [LO, HI] of sizes (see inst.cu).
// `mode` = 0 (generic) or one of PDSP_SPEC_MODES; the caller guarantees the specialised
// kernels' preconditions (whole vector-aligned frames).
template <typename T, int L, int LO, int HI>
cudaError_t r2c_case(int mode, const R2CParams& p, const LaunchCtx& lc) {
  if constexpr (L >= LO && L <= HI) {
    if constexpr (L >= kMinSpecLog2M) {
      if constexpr (staged_supported<T, L>()) {
        switch (mode) {
#define X(m)              \
  case ((m) | MD_STAGED): \
    return launch_r2c_t<T, L, ((m) | MD_STAGED)>(p, lc);
          PDSP_STAGED_MODES(X)
#undef X
          default:
            break;
        }
      }
      switch (mode & ~MD_STAGED) {  // no staged form of this mode / size: the direct-load kernel
#define X(m) \
  case (m):  \
    return launch_r2c_t<T, L, (m)>(p, lc);
        PDSP_SPEC_MODES(X)
#undef X
        default:
          break;
      }
    }
    return launch_r2c_t<T, L, MD_GENERIC>(p, lc);
  } else {
    return cudaErrorInvalidValue;
  }
}
template <typename T, int L, int LO, int HI>
cudaError_t c2c_case(const C2CParams& p, const LaunchCtx& lc) {
  if constexpr (L >= LO && L <= HI)
    return launch_c2c_t<T, L>(p, lc);
  else
    return cudaErrorInvalidValue;
}

typedef cudaError_t (*r2c_group_fn)(int log2m, int mode, const R2CParams& p, const LaunchCtx& lc);
typedef cudaError_t (*c2c_group_fn)(int log2m, const C2CParams& p, const LaunchCtx& lc);

}  // namespace pdsp
