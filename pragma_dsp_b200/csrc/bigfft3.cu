// bigfft3.cu - instantiations and launcher of the third-generation ("pipeline") large-FFT pass kernel
// (bigfft3_kernels.cuh): passes of 256 / 512 / 1024 points over the interleaved work buffer.
#include "bigfft3_kernels.cuh"
#include "fft_launch.cuh"

namespace pdsp {

template <typename T, int LOG2L, int IO>
static cudaError_t launch_pipe_io(const BigPassParams& p, const simt::TensorMap2D& tm_re, const simt::TensorMap2D& tm_im,
                                  const LaunchCtx& lc) {
  using B = BigCfg<T, LOG2L>;
  using E = FftEngine<T, LOG2L, B::LOG2P, B::MAXRB>;
  using S = BigPipeSmem<T, LOG2L, B::LOG2P, B::MAXRB, B::C>;
  constexpr int THREADS = B::TF * B::C;
  auto kern = bigfft_pipe_kernel<T, LOG2L, B::LOG2P, B::MAXRB, B::C, IO>;
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
static cudaError_t launch_pipe_t(bool last, const BigPassParams& p, const simt::TensorMap2D& tm_re, const simt::TensorMap2D& tm_im,
                                 const LaunchCtx& lc) {
  if (last) return p.in_cplx && !p.out_cplx ? launch_pipe_io<T, LOG2L, 5>(p, tm_re, tm_im, lc) : cudaErrorInvalidValue;
  if (!p.out_cplx) return cudaErrorInvalidValue;
  return p.in_cplx ? launch_pipe_io<T, LOG2L, 3>(p, tm_re, tm_im, lc) : launch_pipe_io<T, LOG2L, 2>(p, tm_re, tm_im, lc);
}

// whether a pass of 2^log2l points of this precision has a pipeline form (fp64 only: the fp32 tiles are half the size
// and already leave room for two CTAs per SM in the earlier generations)
bool big_pipe_supported(bool f64, int log2l) { return f64 && log2l >= kPipeMinLog2L && log2l <= kPipeMaxLog2L; }

// `last`: contiguous rows in, transposed planar output (tensor maps unused); otherwise a strided pass with the
// first generation's tile box ({C, min(L, 256)} per plane, {2*C, min(L, 256)/2} over the interleaved buffer)
cudaError_t launch_big_pipe(bool f64, int log2l, bool last, const BigPassParams& p, const simt::TensorMap2D& tm_re,
                            const simt::TensorMap2D& tm_im, const LaunchCtx& lc) {
  if (!f64) return cudaErrorInvalidValue;
  switch (log2l) {
    case 8: return launch_pipe_t<double, 8>(last, p, tm_re, tm_im, lc);
    case 9: return launch_pipe_t<double, 9>(last, p, tm_re, tm_im, lc);
    case 10: return launch_pipe_t<double, 10>(last, p, tm_re, tm_im, lc);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace pdsp
