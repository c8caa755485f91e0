// emu.cpp - host SIMT emulator for the kernel sources in pragma_dsp_b200/csrc (TEST ONLY).
//
// Compiles fft_kernels.cuh with g++ -DPDSP_EMU and runs every CUDA thread of a CTA as an OS
// thread: bar.sync / __syncwarp become counting barriers, shuffles go through a per-warp
// mailbox, dynamic shared memory is a per-CTA heap buffer.  It exists so that index math,
// synchronisation and the epilogue logic of the real kernels can be checked against the oracle
// in the build container (no GPU).  Arithmetic differs from the device only by FMA contraction.
// Never linked into libpragma_b200.so; the product has no CPU path.

#define PDSP_EMU 1
#include "../../pragma_dsp_b200/csrc/fft_config.h"
#include "../../pragma_dsp_b200/csrc/fft_kernels.cuh"

using namespace pdsp;

template <typename T, int LOG2M, int MODE, int VAR = 0>
static int run_r2c_m(const R2CParams& p, int nblocks) {
  using C = KCfg<T, LOG2M, VAR>;
  using E = FftEngine<T, LOG2M, C::LOG2P, C::MAXRB>;
  constexpr int THREADS = C::THREADS, SLOTS = THREADS / E::TF;
  constexpr size_t SMEM = r2c_smem_bytes<T, E, MODE, SLOTS>() + sizeof(cx<T>);
  if constexpr ((MODE & MD_STAGED) != 0 && E::TF > 32) return -3;
  else
  simt::emu_launch(nblocks, THREADS, SMEM, [&] { r2c_kernel<T, LOG2M, C::LOG2P, C::MAXRB, THREADS, C::MINB, MODE>(p); });
  return 0;
}
// mode: 0 generic, else a specialised mode (only instantiated for the sizes the tests use)
template <typename T, int LOG2M>
static int run_r2c(const R2CParams& p, int nblocks, int mode) {
  if constexpr (LOG2M == 5 || LOG2M == 7 || LOG2M == 9 || LOG2M == 11) {
    switch (mode) {
      case MD_AMP: return run_r2c_m<T, LOG2M, MD_AMP>(p, nblocks);
      case MD_AMP | MD_PEAK: return run_r2c_m<T, LOG2M, MD_AMP | MD_PEAK>(p, nblocks);
      case MD_PEAK: return run_r2c_m<T, LOG2M, MD_PEAK>(p, nblocks);
      case MD_CPLX: return run_r2c_m<T, LOG2M, MD_CPLX>(p, nblocks);
      case MD_AMP | MD_PHASE | MD_PEAK: return run_r2c_m<T, LOG2M, MD_AMP | MD_PHASE | MD_PEAK>(p, nblocks);
      case MD_AMP | MD_TWO: return run_r2c_m<T, LOG2M, MD_AMP | MD_TWO>(p, nblocks);
      case MD_AMP | MD_PAD: return run_r2c_m<T, LOG2M, MD_AMP | MD_PAD>(p, nblocks);
      case MD_AMP | MD_PEAK | MD_PAD: return run_r2c_m<T, LOG2M, MD_AMP | MD_PEAK | MD_PAD>(p, nblocks);
      case MD_AMP | MD_PHASE | MD_PEAK | MD_PAD: return run_r2c_m<T, LOG2M, MD_AMP | MD_PHASE | MD_PEAK | MD_PAD>(p, nblocks);
      case MD_AMP | MD_PHASE | MD_PEAK | MD_TWO: return run_r2c_m<T, LOG2M, MD_AMP | MD_PHASE | MD_PEAK | MD_TWO>(p, nblocks);
      // bulk-staged sample loads (frames that fit a warp)
      case MD_AMP | MD_STAGED: return run_r2c_m<T, LOG2M, MD_AMP | MD_STAGED>(p, nblocks);
      case MD_AMP | MD_PEAK | MD_STAGED: return run_r2c_m<T, LOG2M, MD_AMP | MD_PEAK | MD_STAGED>(p, nblocks);
      case MD_PEAK | MD_STAGED: return run_r2c_m<T, LOG2M, MD_PEAK | MD_STAGED>(p, nblocks);
      case MD_CPLX | MD_STAGED: return run_r2c_m<T, LOG2M, MD_CPLX | MD_STAGED>(p, nblocks);
      case MD_AMP | MD_PHASE | MD_PEAK | MD_STAGED: return run_r2c_m<T, LOG2M, MD_AMP | MD_PHASE | MD_PEAK | MD_STAGED>(p, nblocks);
      default: break;
    }
  }
  if (mode != 0) return -2;
  return run_r2c_m<T, LOG2M, MD_GENERIC>(p, nblocks);
}
template <typename T, int LOG2M>
static int run_c2c(const C2CParams& p, int nblocks) {
  using C = KCfg<T, LOG2M>;
  using E = FftEngine<T, LOG2M, C::LOG2P, C::MAXRB>;
  constexpr int THREADS = C::THREADS, SLOTS = THREADS / E::TF;
  constexpr size_t SMEM = sizeof(cx<T>) * E::SMEM_ELEMS * SLOTS;
  simt::emu_launch(nblocks, THREADS, SMEM, [&] { c2c_kernel<T, LOG2M, C::LOG2P, C::MAXRB, THREADS, C::MINB>(p); });
  return 0;
}

// The instantiations are split over several translation units (-DEMU_PART=k, built in parallel by
// tests/simt_emu/__init__.py): parts 0-4 hold the r2c kernels of a group of sizes, 5 / 6 the fp64 / fp32 tuning
// variants, 7 the c2c kernels, the configuration query and the dispatchers.
#ifndef EMU_PART
#error "compile with -DEMU_PART=0..7"
#endif
#define EMU_SIZES(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11)
#if EMU_PART == 0
#define EMU_PART_SIZES(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6)
#define EMU_PART_NAME emu_r2c_part0
#elif EMU_PART == 1
#define EMU_PART_SIZES(X) X(7) X(8)
#define EMU_PART_NAME emu_r2c_part1
#elif EMU_PART == 2
#define EMU_PART_SIZES(X) X(9)
#define EMU_PART_NAME emu_r2c_part2
#elif EMU_PART == 3
#define EMU_PART_SIZES(X) X(10)
#define EMU_PART_NAME emu_r2c_part3
#elif EMU_PART == 4
#define EMU_PART_SIZES(X) X(11)
#define EMU_PART_NAME emu_r2c_part4
#endif

#if EMU_PART <= 4
extern "C" int EMU_PART_NAME(int f64, int log2m, const R2CParams* p, int nblocks, int mode) {
  switch (log2m) {
#define X(L) \
  case L:    \
    return f64 ? run_r2c<double, L>(*p, nblocks, mode) : run_r2c<float, L>(*p, nblocks, mode);
    EMU_PART_SIZES(X)
#undef X
  }
  return -1;
}
#endif

#if EMU_PART == 7
extern "C" int emu_r2c_part0(int, int, const R2CParams*, int, int);
extern "C" int emu_r2c_part1(int, int, const R2CParams*, int, int);
extern "C" int emu_r2c_part2(int, int, const R2CParams*, int, int);
extern "C" int emu_r2c_part3(int, int, const R2CParams*, int, int);
extern "C" int emu_r2c_part4(int, int, const R2CParams*, int, int);
extern "C" int emu_r2c(int f64, int log2m, const R2CParams* p, int nblocks, int mode) {
  if (log2m < 0 || log2m > 11) return -1;
  if (log2m <= 6) return emu_r2c_part0(f64, log2m, p, nblocks, mode);
  if (log2m <= 8) return emu_r2c_part1(f64, log2m, p, nblocks, mode);
  if (log2m == 9) return emu_r2c_part2(f64, log2m, p, nblocks, mode);
  if (log2m == 10) return emu_r2c_part3(f64, log2m, p, nblocks, mode);
  return emu_r2c_part4(f64, log2m, p, nblocks, mode);
}
extern "C" int emu_c2c(int f64, int log2m, const C2CParams* p, int nblocks) {
  switch (log2m) {
#define X(L) \
  case L:    \
    return f64 ? run_c2c<double, L>(*p, nblocks) : run_c2c<float, L>(*p, nblocks);
    EMU_SIZES(X)
#undef X
  }
  return -1;
}
#endif

#if EMU_PART == 5 || EMU_PART == 6
template <typename T, int VAR>
static int run_var(const R2CParams& p, int nblocks, int mode) {
  switch (mode) {
    case MD_AMP: return run_r2c_m<T, kVariantLog2M, MD_AMP, VAR>(p, nblocks);
    case MD_AMP | MD_PEAK: return run_r2c_m<T, kVariantLog2M, MD_AMP | MD_PEAK, VAR>(p, nblocks);
    case MD_PEAK: return run_r2c_m<T, kVariantLog2M, MD_PEAK, VAR>(p, nblocks);
  }
  return -2;
}
#if EMU_PART == 5
extern "C" int emu_r2c_var_f64(int var, const R2CParams* p, int nblocks, int mode) {
  using VT = double;
#else
extern "C" int emu_r2c_var_f32(int var, const R2CParams* p, int nblocks, int mode) {
  using VT = float;
#endif
  switch (var) {
#define X(V) \
  case V:    \
    return run_var<VT, V>(*p, nblocks, mode);
    X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17)
#undef X
  }
  return -1;
}
#endif

#if EMU_PART == 7
extern "C" int emu_r2c_var_f64(int, const R2CParams*, int, int);
extern "C" int emu_r2c_var_f32(int, const R2CParams*, int, int);
extern "C" int emu_r2c_var(int f64, int var, const R2CParams* p, int nblocks, int mode) {
  return f64 ? emu_r2c_var_f64(var, p, nblocks, mode) : emu_r2c_var_f32(var, p, nblocks, mode);
}
// kernel configuration of (type, size, variant): lets the harness build the per-pass twiddle table
template <typename T, int LOG2M, int VAR>
static void cfg_of(int* out) {
  using C = KCfg<T, LOG2M, VAR>;
  using E = FftEngine<T, LOG2M, C::LOG2P, C::MAXRB>;
  out[0] = C::LOG2P;
  out[1] = E::RB;
  out[2] = E::TW_ELEMS;
}
extern "C" int emu_cfg(int f64, int log2m, int var, int* out) {
  if (var) {
    if (log2m != kVariantLog2M) return -1;
    switch (var) {
#define X(V) \
  case V:    \
    f64 ? cfg_of<double, kVariantLog2M, V>(out) : cfg_of<float, kVariantLog2M, V>(out); \
    return 0;
      X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17)
#undef X
    }
    return -1;
  }
  switch (log2m) {
#define X(L) \
  case L:    \
    f64 ? cfg_of<double, L, 0>(out) : cfg_of<float, L, 0>(out); \
    return 0;
    EMU_SIZES(X)
#undef X
  }
  return -1;
}
extern "C" int emu_params_size(int which) { return which == 0 ? (int)sizeof(R2CParams) : (int)sizeof(C2CParams); }
#endif  // EMU_PART == 7
