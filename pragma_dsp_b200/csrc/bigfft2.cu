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

// ---- fused middle + last pass (bigfft_fused_kernel): fp64, interleaved work buffer, both passes of 256 or 512 points
template <int LA, int LB>
static cudaError_t launch_fused_t(const BigTileParams& pa, const BigTileParams& pb, const BigFusedSync& fs, const simt::TensorMap* ma,
                                  const simt::TensorMap* mb, const LaunchCtx& lc) {
  using A = BigCfg2<double, LA>;
  using B = BigCfg2<double, LB>;
  using EA = FftEngine<double, LA, A::LOG2P, A::MAXRB>;
  using EB = FftEngine<double, LB, B::LOG2P, B::MAXRB>;
  static_assert(A::LOG2P == B::LOG2P, "one CTA shape for both passes");
  constexpr int THREADS = A::TF * A::C;
  constexpr size_t SMEM = A::SMEM > B::SMEM ? A::SMEM : B::SMEM;
  auto kern = bigfft_fused_kernel<double, LA, LB, A::LOG2P, A::MAXRB, B::MAXRB, A::C, B::C>;
  static int bps[kMaxDevices] = {0};
  const long long total = fs.n_groups * (fs.n1g + fs.n2g);
  if (total <= 0) return cudaSuccess;
  int grid = 0;
  cudaError_t e = persistent_grid(kern, THREADS, SMEM, lc, bps, total, &grid);
  if (e != cudaSuccess) return e;
#ifdef PDSP_EMU
  grid = 1;  // the emulator runs CTAs one after another: a single CTA takes the items in list order (producers first)
#endif
  BigTileParams qa = pa, qb = pb;
  qa.tw = lc.pass_twiddles(lc.owner, true, LA, EA::RB);
  qb.tw = lc.pass_twiddles(lc.owner, true, LB, EB::RB);
  if (!qa.tw || !qb.tw) return cudaErrorInvalidValue;
  PDSP_LAUNCH(kern, grid, THREADS, SMEM, lc.stream, qa, qb, fs, ma[0], ma[2], mb[2], mb[3]);
  return cudaGetLastError();
}
bool big_fused_supported(bool f64, int la, int lb) {
#ifdef PDSP_EMU
  if (f64 && la == 6 && lb == 6) return true;  // test build only: the same list / counter logic at a size the emulator can run
#endif
  return f64 && (la == 8 || la == 9) && (lb == 8 || lb == 9);
}
// ma / mb: the {in_re, in_im, out_re, out_im} maps of the middle / last pass as launch_big_tile takes them
cudaError_t launch_big_fused(int la, int lb, const BigTileParams& pa, const BigTileParams& pb, const BigFusedSync& fs,
                             const simt::TensorMap* ma, const simt::TensorMap* mb, const LaunchCtx& lc) {
#ifdef PDSP_EMU
  if (la == 6 && lb == 6) return launch_fused_t<6, 6>(pa, pb, fs, ma, mb, lc);
#endif
  if (la == 8 && lb == 8) return launch_fused_t<8, 8>(pa, pb, fs, ma, mb, lc);
  if (la == 9 && lb == 8) return launch_fused_t<9, 8>(pa, pb, fs, ma, mb, lc);
  if (la == 8 && lb == 9) return launch_fused_t<8, 9>(pa, pb, fs, ma, mb, lc);
  if (la == 9 && lb == 9) return launch_fused_t<9, 9>(pa, pb, fs, ma, mb, lc);
  return cudaErrorInvalidValue;
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
