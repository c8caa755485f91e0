// bigfft2.cu - instantiations and launcher of the second-generation large-FFT pass kernel (bigfft2_kernels.cuh).
#include "bigfft2_kernels.cuh"
#include "fft_launch.cuh"

namespace pdsp {

template <typename T, int LOG2L, int IO>
static cudaError_t launch_tile_io(const BigTileParams& p, const simt::TensorMap* maps, const LaunchCtx& lc) {
  using B = BigCfg2<T, LOG2L>;
  using E = FftEngine<T, LOG2L, B::LOG2P, B::MAXRB>;
  constexpr int THREADS = B::TF * B::C;
  auto kern = bigfft_tile_kernel<T, LOG2L, B::LOG2P, B::MAXRB, B::C, IO>;
  static int bps[kMaxDevices] = {0};
  if (p.n_groups <= 0 || p.n_frames <= 0) return cudaSuccess;
  int grid = 0;
  cudaError_t e = persistent_grid(kern, THREADS, B::SMEM, lc, bps, p.n_groups * p.n_frames, &grid);
  if (e != cudaSuccess) return e;
  BigTileParams q = p;
  q.tw = lc.pass_twiddles(lc.owner, sizeof(T) == 8, LOG2L, E::RB);
  if (!q.tw) return cudaErrorInvalidValue;
  PDSP_LAUNCH(kern, grid, THREADS, B::SMEM, lc.stream, q, maps[0], maps[1], maps[2], maps[3]);
  return cudaGetLastError();
}

template <typename T, int LOG2L>
static cudaError_t launch_tile_t(int io, const BigTileParams& p, const simt::TensorMap* maps, const LaunchCtx& lc) {
  switch (io) {
    case 0: return launch_tile_io<T, LOG2L, 0>(p, maps, lc);  // planar -> planar (planar work planes)
    case 2: return launch_tile_io<T, LOG2L, 2>(p, maps, lc);  // planar -> interleaved (first pass)
    case 3: return launch_tile_io<T, LOG2L, 3>(p, maps, lc);  // interleaved -> interleaved (middle pass)
    case 4: return launch_tile_io<T, LOG2L, 4>(p, maps, lc);  // last pass, planar in
    case 5: return launch_tile_io<T, LOG2L, 5>(p, maps, lc);  // last pass, interleaved in
    default: return cudaErrorInvalidValue;
  }
}

// sequences per CTA of a pass of length 2^log2l
int big2_pass_c(int log2l) {
  switch (log2l) {
    case 6: return BigCfg2<double, 6>::C;
    case 7: return BigCfg2<double, 7>::C;
    case 8: return BigCfg2<double, 8>::C;
    case 9: return BigCfg2<double, 9>::C;
    default: return BigCfg2<double, 10>::C;
  }
}

// maps: {in_re, in_im, out_re, out_im} (unused entries may repeat a valid map)
cudaError_t launch_big_tile(bool f64, int log2l, int io, const BigTileParams& p, const simt::TensorMap* maps, const LaunchCtx& lc) {
  switch (log2l) {
#define X(L) \
  case L:    \
    return f64 ? launch_tile_t<double, L>(io, p, maps, lc) : launch_tile_t<float, L>(io, p, maps, lc);
    X(6) X(7) X(8) X(9) X(10)
#undef X
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace pdsp
