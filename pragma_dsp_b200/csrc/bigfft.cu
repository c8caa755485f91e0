// bigfft.cu - instantiations and launcher of the multi-pass large-N FFT kernel (bigfft_kernels.cuh).
#include "bigfft_kernels.cuh"
#include "fft_launch.cuh"

namespace pdsp {

int big_pass_c(int log2l);

template <typename T, int LOG2L, int IO>
static cudaError_t launch_big_io(const BigPassParams& p, const LaunchCtx& lc) {
  using B = BigCfg<T, LOG2L>;
  using E = FftEngine<T, LOG2L, B::LOG2P, B::MAXRB>;
  constexpr int THREADS = B::TF * B::C;
  constexpr size_t SMEM = sizeof(cx<T>) * (size_t)(E::SMEM_ELEMS | 1) * B::C;
  auto kern = bigfft_pass_kernel<T, LOG2L, B::LOG2P, B::MAXRB, B::C, IO>;
  static int bps[kMaxDevices] = {0};
  if (p.n_groups <= 0 || p.n_frames <= 0) return cudaSuccess;
  int grid = 0;
  cudaError_t e = persistent_grid(kern, THREADS, SMEM, lc, bps, p.n_groups * p.n_frames, &grid);
  if (e != cudaSuccess) return e;
  BigPassParams q = p;
  q.tw = lc.pass_twiddles(lc.owner, sizeof(T) == 8, LOG2L, E::RB);
  if (!q.tw) return cudaErrorInvalidValue;
  PDSP_LAUNCH(kern, grid, THREADS, SMEM, lc.stream, q);
  return cudaGetLastError();
}

template <typename T, int LOG2L>
static cudaError_t launch_big_t(const BigPassParams& p, const LaunchCtx& lc) {
  switch ((p.in_cplx ? 1 : 0) | (p.out_cplx ? 2 : 0)) {
    case 0: return launch_big_io<T, LOG2L, 0>(p, lc);
    case 1: return launch_big_io<T, LOG2L, 1>(p, lc);
    case 2: return launch_big_io<T, LOG2L, 2>(p, lc);
    default: return launch_big_io<T, LOG2L, 3>(p, lc);
  }
}

template <typename T, int LOG2L, int IO>
static cudaError_t launch_big_tma_io(const BigPassParams& p, const simt::TensorMap2D& tm_re, const simt::TensorMap2D& tm_im,
                                     const LaunchCtx& lc) {
  using B = BigCfg<T, LOG2L>;
  using E = FftEngine<T, LOG2L, B::LOG2P, B::MAXRB>;
  using S = BigTmaSmem<T, LOG2L, B::LOG2P, B::MAXRB, B::C>;
  constexpr int THREADS = B::TF * B::C;
  auto kern = bigfft_pass_tma_kernel<T, LOG2L, B::LOG2P, B::MAXRB, B::C, S::STAGES, IO>;
  static int bps[kMaxDevices] = {0};
  if (p.n_groups <= 0 || p.n_frames <= 0) return cudaSuccess;
  int grid = 0;
  cudaError_t e = persistent_grid(kern, THREADS, S::BYTES, lc, bps, p.n_groups * p.n_frames, &grid);
  if (e != cudaSuccess) return e;
  BigPassParams q = p;
  q.tw = lc.pass_twiddles(lc.owner, sizeof(T) == 8, LOG2L, E::RB);
  if (!q.tw) return cudaErrorInvalidValue;
  PDSP_LAUNCH(kern, grid, THREADS, S::BYTES, lc.stream, q, tm_re, tm_im);
  return cudaGetLastError();
}

template <typename T, int LOG2L>
static cudaError_t launch_big_tma_t(const BigPassParams& p, const simt::TensorMap2D& tm_re, const simt::TensorMap2D& tm_im,
                                    const LaunchCtx& lc) {
  // never the last pass: the output kind equals the work-buffer kind
  switch ((p.in_cplx ? 1 : 0) | (p.out_cplx ? 2 : 0)) {
    case 0: return launch_big_tma_io<T, LOG2L, 0>(p, tm_re, tm_im, lc);
    case 2: return launch_big_tma_io<T, LOG2L, 2>(p, tm_re, tm_im, lc);
    case 3: return launch_big_tma_io<T, LOG2L, 3>(p, tm_re, tm_im, lc);
    default: return cudaErrorInvalidValue;
  }
}

// box of the TMA tile of a pass: {C columns, min(L, 256) rows}
void big_pass_tma_box(int log2l, int* cols, int* rows) {
  *cols = big_pass_c(log2l);
  *rows = (1 << log2l) < 256 ? (1 << log2l) : 256;
}

cudaError_t launch_big_pass_tma(bool f64, int log2l, const BigPassParams& p, const simt::TensorMap2D& tm_re,
                                const simt::TensorMap2D& tm_im, const LaunchCtx& lc) {
  switch (log2l) {
#define X(L) \
  case L:    \
    return f64 ? launch_big_tma_t<double, L>(p, tm_re, tm_im, lc) : launch_big_tma_t<float, L>(p, tm_re, tm_im, lc);
    X(6) X(7) X(8) X(9) X(10)
#undef X
    default:
      return cudaErrorInvalidValue;
  }
}

int big_pass_c(int log2l) {
  switch (log2l) {
    case 6: return BigCfg<double, 6>::C;
    case 7: return BigCfg<double, 7>::C;
    case 8: return BigCfg<double, 8>::C;
    case 9: return BigCfg<double, 9>::C;
    default: return BigCfg<double, 10>::C;
  }
}

cudaError_t launch_big_pass(bool f64, int log2l, const BigPassParams& p, const LaunchCtx& lc) {
  switch (log2l) {
#define X(L) \
  case L:    \
    return f64 ? launch_big_t<double, L>(p, lc) : launch_big_t<float, L>(p, lc);
    X(6) X(7) X(8) X(9) X(10)
#undef X
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace pdsp
