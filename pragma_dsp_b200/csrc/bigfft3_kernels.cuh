// bigfft3_kernels.cuh - K2, third generation ("pipeline" passes): one pass of the multi-pass large-N FFT with the NEXT
// tile landing in shared memory while the current one is transformed and stored.
//
// The first generation (bigfft_kernels.cuh) runs one 512-thread CTA per SM whose tile (8192 complex points: 128 KB in
// fp64) aliases the exchange buffers, so load -> transform -> store of a tile are strictly serial: a 2^20 pass spends
// ~11 us per tile of which ~3 us is arithmetic.  The second generation (bigfft2_kernels.cuh) overlaps two CTAs of half
// tiles, which pays for passes of <= 256 points but halves the row length of the strided tiles (2^20 = 1024 x 1024:
// 32/64-byte rows, no faster).  Here the tile keeps its full width and the overlap happens inside one CTA:
//
//   * the exchange buffer of the Stockham engine holds one scalar plane at a time (FftEngine::fft<.., SPLITX>): 70 KB
//     instead of 139 KB, which leaves room for a landing zone of a whole tile (128 KB) beside it;
//   * as soon as the tile has been copied from the landing zone into registers, one thread re-arms the mbarrier and
//     issues the TMA loads of the CTA's next tile (cp.async.bulk.tensor box loads for the strided passes; for the last
//     pass, whose rows are contiguous, one cp.async.bulk per row) - they complete during the transform, the twiddle
//     multiply and the (fire-and-forget) stores of the current tile;
//   * results leave from registers with the sequence index fastest across lanes, as in the first generation:
//     C * 16-byte segments into the interleaved work buffer, C * 8-byte segments for the transposed
//     (digit-reversed) output of the last pass.
//
// Same pass structure and parameter block as the first generation (BigPassParams); the reference runs the same
// transform as log2(N) strided in-place sweeps (/root/reference/src/core/fft.ts:116-140).
#pragma once
#include "bigfft_kernels.cuh"

namespace pdsp {

// shared-memory plan for (T, LOG2L, C)
template <typename T, int LOG2L, int LOG2P, int MAXRB, int C>
struct BigPipeSmem {
  using E = FftEngine<T, LOG2L, LOG2P, MAXRB>;
  static constexpr int L = E::M;
  // scalar slot of one sequence: the engine's padded indexing, rounded up so that the sequences sharing a wavefront
  // (128 / sizeof(T) lanes; C of them differ in the sequence index) start in distinct banks
  static constexpr int W = 128 / (int)sizeof(T);
  static constexpr int RES = W / C > 0 ? W / C : 1;
  static constexpr int NEED = L + (L >> E::PAD_SHIFT) + 1;
  static constexpr int SLOT = NEED + ((RES - NEED % W) % W + W) % W;
  static constexpr size_t PLANE = sizeof(T) * (size_t)L * C;
  static constexpr size_t ROW_PITCH = sizeof(cx<T>) * (size_t)L + 16;  // last pass: rows 16 bytes apart modulo 128
  static constexpr size_t LZ = (2 * PLANE > ROW_PITCH * C ? 2 * PLANE : ROW_PITCH * C);
  static constexpr size_t LZ_AL = (LZ + 127) & ~(size_t)127;
  static constexpr size_t EXCH = sizeof(T) * (size_t)SLOT * C;
  static constexpr size_t BAR_OFF = (LZ_AL + EXCH + 15) & ~(size_t)15;
  static constexpr size_t BYTES = BAR_OFF + 16 + 128;
};

// IO bit 0: input is the interleaved work buffer; bit 1: output is; bit 2: last pass (contiguous rows in, transposed out)
template <typename T, int LOG2L, int LOG2P, int MAXRB, int C, int IO>
PDSP_GLOBAL void PDSP_LAUNCH_BOUNDS(((1 << LOG2L) >> LOG2P) * C, 1)
    bigfft_pipe_kernel(const BigPassParams p, const PDSP_GRID_CONSTANT simt::TensorMap2D tm_re,
                       const PDSP_GRID_CONSTANT simt::TensorMap2D tm_im) {
  constexpr bool IN_CPLX = (IO & 1) != 0, OUT_CPLX = (IO & 2) != 0, LAST = (IO & 4) != 0;
  static_assert(!LAST || IN_CPLX, "the last pass reads the interleaved work buffer");
  using E = FftEngine<T, LOG2L, LOG2P, MAXRB>;
  using S = BigPipeSmem<T, LOG2L, LOG2P, MAXRB, C>;
  constexpr int L = E::M, P = E::P, TF = E::TF;
  constexpr int BOX_ROWS = L < 256 ? L : 256;
  constexpr int NBOX = L / BOX_ROWS;
  constexpr size_t PLANE = S::PLANE;
  const int tid = simt::tid();
  const int c = tid % C;
  const int t = tid / C;
  unsigned char* base = simt::smem();
  base += (128 - (reinterpret_cast<uintptr_t>(base) & 127)) & 127;
  T* xb = reinterpret_cast<T*>(base + S::LZ_AL) + (size_t)c * S::SLOT;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(base + S::BAR_OFF);
  T* PDSP_RESTRICT ore = static_cast<T*>(p.out_re);
  T* PDSP_RESTRICT oim = static_cast<T*>(p.out_im);
  cx<T>* PDSP_RESTRICT ocx = static_cast<cx<T>*>(p.out_re);
  const cx<T>* PDSP_RESTRICT icx = static_cast<const cx<T>*>(p.in_re);
  const T scale = (T)p.scale;
  const bool has_im = p.in_im != nullptr && !IN_CPLX;
  const long long total = p.n_groups * p.n_frames;
  const long long n_hi = p.n_groups / p.n_lo;

  auto issue = [&](long long w) {  // one thread: arm the barrier, launch the loads of tile w into the landing zone
    const long long fb = w / p.n_groups, g = w % p.n_groups;
    const long long g_hi = g / p.n_lo, g_lo = g % p.n_lo;
    if constexpr (LAST) {
      const long long in_base = fb * p.in_frame + g_hi * p.in_hi + g_lo * p.in_lo;
      simt::mbar_expect_tx(bar, (unsigned)(sizeof(cx<T>) * (size_t)L * C));
      for (int cc = 0; cc < C; ++cc)
        simt::bulk_load_1d(base + (size_t)cc * S::ROW_PITCH, icx + in_base + cc * p.in_c, (unsigned)(sizeof(cx<T>) * L), bar);
    } else {
      const int x = (int)(g_lo * C);
      const int y = (int)((fb * n_hi + g_hi) * L);
      simt::mbar_expect_tx(bar, (unsigned)((has_im || IN_CPLX ? 2 : 1) * PLANE));
      if constexpr (IN_CPLX) {  // [L][C] cx<T>, as 2*NBOX boxes of {2*C scalars, BOX_ROWS/2 rows}
        for (int j = 0; j < 2 * NBOX; ++j)
          simt::tma_load_2d(base + (size_t)j * (BOX_ROWS / 2) * C * sizeof(cx<T>), &tm_re, 2 * x, y + j * (BOX_ROWS / 2), bar);
      } else {
        for (int j = 0; j < NBOX; ++j) {
          simt::tma_load_2d(base + (size_t)j * BOX_ROWS * C * sizeof(T), &tm_re, x, y + j * BOX_ROWS, bar);
          if (has_im) simt::tma_load_2d(base + PLANE + (size_t)j * BOX_ROWS * C * sizeof(T), &tm_im, x, y + j * BOX_ROWS, bar);
        }
      }
    }
  };

  if (tid == 0) simt::mbar_init(bar, 1);
  simt::sync_block();
  if (tid == 0 && simt::bid() < total) issue(simt::bid());
  unsigned phase = 0u;

  for (long long w = simt::bid(); w < total; w += simt::nblocks()) {
    simt::mbar_wait(bar, phase);
    phase ^= 1u;
    cx<T> v[P];
    if constexpr (LAST) {
      const cx<T>* row = reinterpret_cast<const cx<T>*>(base + (size_t)c * S::ROW_PITCH);
      static_for<0, P>([&](auto qi) { v[decltype(qi)::value] = row[t + TF * decltype(qi)::value]; });
    } else {
      const T* tre = reinterpret_cast<const T*>(base);
      const T* tim = reinterpret_cast<const T*>(base + PLANE);
      static_for<0, P>([&](auto qi) {
        constexpr int q = decltype(qi)::value;
        const int e = t + TF * q;
        if constexpr (IN_CPLX) {
          v[q] = reinterpret_cast<const cx<T>*>(tre)[e * C + c];
        } else {
          const T re = tre[e * C + c];
          const T im = has_im ? tim[e * C + c] : (T)0;
          v[q] = p.swap_in ? cx<T>{im, re} : cx<T>{re, im};
        }
      });
    }
    simt::sync_block();  // the landing zone has been read by every thread: refill it with the next tile
    if (tid == 0 && w + simt::nblocks() < total) issue(w + simt::nblocks());

    E::template fft<true, typename E::NoHook, true>(v, t, reinterpret_cast<cx<T>*>(xb), static_cast<const cx<T>*>(p.tw), 0, 1);

    const long long fb = w / p.n_groups, g = w % p.n_groups;
    const long long g_hi = g / p.n_lo, g_lo = g % p.n_lo;
    if (p.tw_hi != nullptr) {
      // W_NT^{k*i}, k = t + TF*q: start at W^{t*i}, step by W^{TF*i}
      const cx<T>* PDSP_RESTRICT hi = static_cast<const cx<T>*>(p.tw_hi);
      const cx<T>* PDSP_RESTRICT lo = static_cast<const cx<T>*>(p.tw_lo);
      const long long i = g_lo * C + c;
      cx<T> wv = big_twiddle(hi, lo, (long long)t * i, p.log_b);
      const cx<T> step = big_twiddle(hi, lo, (long long)TF * i, p.log_b);
      static_for<0, P>([&](auto qi) {
        constexpr int q = decltype(qi)::value;
        v[q] = cmul(v[q], wv);
        if constexpr (q + 1 < P) wv = cmul(wv, step);
      });
    }
    const long long out_base = fb * p.out_frame + g_hi * p.out_hi + g_lo * p.out_lo;
    static_for<0, P>([&](auto qi) {
      constexpr int q = decltype(qi)::value;
      const long long a = out_base + c * p.out_c + (long long)(t + TF * q) * p.out_e;
      if constexpr (OUT_CPLX) {
        ocx[a] = v[q];
      } else {
        const T x = v[q].x * scale, y = v[q].y * scale;
        ore[a] = p.swap_out ? y : x;
        oim[a] = p.swap_out ? x : y;
      }
    });
  }
}

constexpr int kPipeMinLog2L = 8, kPipeMaxLog2L = 10;  // tiles of 8192 points: C = 32 / 16 / 8 sequences

}  // namespace pdsp
