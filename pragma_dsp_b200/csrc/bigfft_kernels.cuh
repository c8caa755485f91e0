// bigfft_kernels.cuh - K2: one pass of the multi-pass ("four-step" / "six-step") FFT used for
// transforms too long for one CTA (N > 8192 complex points; BASELINE config C4: N = 2^20, 2^24).
//
// N = L1*L2(*L3).  Pass j transforms the middle axis of the view [O][Lj][I] (stride I) in place,
// then multiplies by the Cooley-Tukey twiddle W_{Lj*I}^{k*i}; the last pass (I = 1) transforms
// contiguous rows and writes them digit-reversed, i.e. transposed, so the output is in natural
// order.  The reference runs the same transform as log2(N) strided in-place sweeps
// (/root/reference/src/core/fft.ts:116-140).
//
// A CTA handles C adjacent sequences at once with the sequence index fastest across lanes
// (tid = t*C + c), so that every global access moves C*sizeof(T)-byte contiguous segments in both
// planes - the strided "column" gather of pass j and the transposed scatter of the last pass are
// coalesced without a separate transpose kernel.  The per-sequence FFT is the same in-register
// Stockham engine as K1 (fft_core.cuh) with block-wide barriers.
#pragma once
#include "fft_core.cuh"

namespace pdsp {

struct BigPassParams {
  const void* in_re;  // T planes; in_im may be null (real input: imaginary plane of zeros)
  const void* in_im;
  void* out_re;
  void* out_im;
  long long n_groups;  // CTA work items per transform; group g = g_hi * n_lo + g_lo
  long long n_lo;
  long long n_frames;                  // transforms in this launch (work item = frame * n_groups + g)
  long long in_frame, out_frame;       // element strides between consecutive transforms
  long long in_hi, in_lo, in_c, in_e;      // element strides of (g_hi, g_lo, sequence c, element e)
  long long out_hi, out_lo, out_c, out_e;  // same for the output index k
  const void* tw;                          // per-pass twiddles of the L-point schedule (set by the launcher)
  const void* tw_hi;                       // two-level inter-pass twiddle W_NT^m = hi[m >> log_b] * lo[m & (B-1)];
  const void* tw_lo;                       // null on the last pass
  int log_b;
  int stage_in;  // 1: sequences are contiguous rows -> load cooperatively through shared memory
  int swap_in;   // inverse transform = swap(FFT(swap(x))) / N
  int swap_out;
  double scale;  // applied on the way out (1/N on the last pass of an inverse)
  int l2_prefetch;  // 1: pull the CTA's next tile into L2 while the current one is transformed
  // Intermediate passes exchange data through an INTERLEAVED work buffer (cx<T> elements, pointer in in_re /
  // out_re, strides still in complex elements): a tile row is then 2*C*sizeof(T) contiguous bytes instead of
  // two segments of half that, which is what the strided rows' DRAM efficiency depends on.  Caller-facing
  // input (first pass) and output (last pass) stay planar like the reference's ComplexArray.
  int in_cplx, out_cplx;
};

template <typename T>
PDSP_DEVICE cx<T> big_twiddle(const cx<T>* PDSP_RESTRICT hi, const cx<T>* PDSP_RESTRICT lo, long long m, int log_b) {
  const cx<T> a = ldg_cx(hi + (m >> log_b));
  const cx<T> b = ldg_cx(lo + (m & ((1LL << log_b) - 1)));
  return cmul(a, b);
}

// IO: bit 0 = the input is the interleaved work buffer, bit 1 = the output is (compile time: a run-time test
// inside the unrolled load / store loops cost 25-50 % of the kernel)
template <typename T, int LOG2L, int LOG2P, int MAXRB, int C, int IO>
PDSP_GLOBAL void PDSP_LAUNCH_BOUNDS(((1 << LOG2L) >> LOG2P) * C, (512 / (((1 << LOG2L) >> LOG2P) * C) > 0 ? 512 / (((1 << LOG2L) >> LOG2P) * C) : 1))
    bigfft_pass_kernel(const BigPassParams p) {
  constexpr bool IN_CPLX = (IO & 1) != 0, OUT_CPLX = (IO & 2) != 0;
  using E = FftEngine<T, LOG2L, LOG2P, MAXRB>;
  constexpr int L = E::M, P = E::P, TF = E::TF;
  constexpr int THREADS = TF * C;
  constexpr int SLOT = E::SMEM_ELEMS | 1;  // odd stride: neighbouring sequences start in neighbouring banks
  const int tid = simt::tid();
  const int c = tid % C;
  const int t = tid / C;
  cx<T>* smem = reinterpret_cast<cx<T>*>(simt::smem());
  cx<T>* sm = smem + (size_t)c * SLOT;
  const cx<T>* PDSP_RESTRICT tw = static_cast<const cx<T>*>(p.tw);
  const T* PDSP_RESTRICT ire = static_cast<const T*>(p.in_re);
  const T* PDSP_RESTRICT iim = static_cast<const T*>(p.in_im);
  T* PDSP_RESTRICT ore = static_cast<T*>(p.out_re);
  T* PDSP_RESTRICT oim = static_cast<T*>(p.out_im);
  const cx<T>* PDSP_RESTRICT icx = static_cast<const cx<T>*>(p.in_re);  // when p.in_cplx
  cx<T>* PDSP_RESTRICT ocx = static_cast<cx<T>*>(p.out_re);              // when p.out_cplx
  const T scale = (T)p.scale;

  for (long long w = simt::bid(); w < p.n_groups * p.n_frames; w += simt::nblocks()) {
    const long long fb = w / p.n_groups, g = w % p.n_groups;
    const long long g_hi = g / p.n_lo, g_lo = g % p.n_lo;
    const long long in_base = fb * p.in_frame + g_hi * p.in_hi + g_lo * p.in_lo;
    const long long out_base = fb * p.out_frame + g_hi * p.out_hi + g_lo * p.out_lo;
    cx<T> v[P];
    if (p.stage_in) {
      // rows are contiguous in memory: read them with the element index fastest, park in shared memory
      for (int idx = tid; idx < C * L; idx += THREADS) {
        const int cc = idx >> LOG2L, e = idx & (L - 1);
        const long long a = in_base + cc * p.in_c + e * p.in_e;
        if constexpr (IN_CPLX) {
          smem[(size_t)cc * SLOT + E::pad(e)] = ldg_cx(icx + a);
        } else {
          const T re = ire[a];
          const T im = iim != nullptr ? iim[a] : (T)0;
          smem[(size_t)cc * SLOT + E::pad(e)] = p.swap_in ? cx<T>{im, re} : cx<T>{re, im};
        }
      }
      simt::sync_block();
      static_for<0, P>([&](auto q) { v[decltype(q)::value] = sm[E::pad(t + TF * decltype(q)::value)]; });
      simt::sync_block();
    } else {
      static_for<0, P>([&](auto qi) {
        constexpr int q = decltype(qi)::value;
        const long long a = in_base + c * p.in_c + (long long)(t + TF * q) * p.in_e;
        if constexpr (IN_CPLX) {
          v[q] = ldg_cx(icx + a);
        } else {
          const T re = ire[a];
          const T im = iim != nullptr ? iim[a] : (T)0;
          v[q] = p.swap_in ? cx<T>{im, re} : cx<T>{re, im};
        }
      });
    }

    if (p.l2_prefetch && w + simt::nblocks() < p.n_groups * p.n_frames) {
      // this CTA is the only one on its SM and load -> transform -> store do not overlap: have the next tile
      // waiting in L2.  One bulk prefetch per contiguous segment (C elements per row, or a whole row).
      const long long w2 = w + simt::nblocks();
      const long long fb2 = w2 / p.n_groups, g2 = w2 % p.n_groups;
      const long long nb = fb2 * p.in_frame + (g2 / p.n_lo) * p.in_hi + (g2 % p.n_lo) * p.in_lo;
      const int nseg = p.stage_in ? C : L;                                  // segments per plane
      const long long seg_stride = p.stage_in ? p.in_c : p.in_e;            // elements between segments
      const unsigned seg_bytes = (unsigned)((p.stage_in ? L : C) * sizeof(T)) * (IN_CPLX ? 2u : 1u);
      for (int idx = tid; idx < (IN_CPLX ? 1 : 2) * nseg; idx += THREADS) {
        const long long off = nb + (long long)(idx % nseg) * seg_stride;
        const void* a;
        if constexpr (IN_CPLX) {
          a = icx + off;
        } else {
          const T* plane = idx < nseg ? ire : iim;
          if (plane == nullptr) continue;
          a = plane + off;
        }
        if (((reinterpret_cast<uintptr_t>(a) | seg_bytes) & 15u) == 0) simt::prefetch_l2_bulk(a, seg_bytes);
      }
    }

    E::template fft<true>(v, t, sm, tw, 0, 1);

    if (p.tw_hi != nullptr) {
      // W_NT^{k*i}, k = t + TF*q: start at W^{t*i}, step by W^{TF*i}
      const cx<T>* PDSP_RESTRICT hi = static_cast<const cx<T>*>(p.tw_hi);
      const cx<T>* PDSP_RESTRICT lo = static_cast<const cx<T>*>(p.tw_lo);
      const long long i = g_lo * C + c;
      cx<T> w = big_twiddle(hi, lo, (long long)t * i, p.log_b);
      const cx<T> step = big_twiddle(hi, lo, (long long)TF * i, p.log_b);
      static_for<0, P>([&](auto qi) {
        constexpr int q = decltype(qi)::value;
        v[q] = cmul(v[q], w);
        if constexpr (q + 1 < P) w = cmul(w, step);
      });
    }

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

// TMA-staged form of a non-last pass.  The C x L tile of each plane (C adjacent columns of the view
// [O*L rows][I cols], row stride I) is fetched by cp.async.bulk.tensor 2-D box loads ({C, <=256 rows} per
// box) issued by one thread and tracked by an mbarrier; with STAGES = 2 the next work item's tile is in
// flight while the current one is transformed, with STAGES = 1 (tile + exchange buffers do not both fit)
// the tile aliases the exchange buffers.  Results are stored with the same coalesced per-thread stores
// as the plain kernel.  Requires frames to be contiguous (in_frame = N) and the inner index to be the
// tensor's contiguous dimension - i.e. any pass but the last.
template <typename T, int LOG2L, int LOG2P, int MAXRB, int C, int STAGES, int IO>
PDSP_GLOBAL void PDSP_LAUNCH_BOUNDS(((1 << LOG2L) >> LOG2P) * C, (512 / (((1 << LOG2L) >> LOG2P) * C) > 0 ? 512 / (((1 << LOG2L) >> LOG2P) * C) : 1))
    bigfft_pass_tma_kernel(const BigPassParams p, const PDSP_GRID_CONSTANT simt::TensorMap2D tm_re,
                           const PDSP_GRID_CONSTANT simt::TensorMap2D tm_im) {
  using E = FftEngine<T, LOG2L, LOG2P, MAXRB>;
  constexpr int L = E::M, P = E::P, TF = E::TF;
  constexpr int SLOT = E::SMEM_ELEMS | 1;
  constexpr int BOX_ROWS = L < 256 ? L : 256;
  constexpr int NBOX = L / BOX_ROWS;
  constexpr size_t PLANE = sizeof(T) * (size_t)L * C;         // one plane of one tile
  constexpr size_t TILES = 2 * PLANE * STAGES;                // re + im, per stage
  constexpr size_t EXCH = sizeof(cx<T>) * (size_t)SLOT * C;
  constexpr size_t EXCH_OFF = STAGES == 2 ? TILES : 0;        // STAGES == 1: the tile aliases the exchange buffers
  constexpr size_t BAR_OFF = ((STAGES == 2 ? TILES + EXCH : (TILES > EXCH ? TILES : EXCH)) + 15) & ~(size_t)15;
  const int tid = simt::tid();
  const int c = tid % C;
  const int t = tid / C;
  // 128-byte aligned carve-up of dynamic shared memory (TMA destinations)
  unsigned char* base = simt::smem();
  base += (128 - (reinterpret_cast<uintptr_t>(base) & 127)) & 127;
  cx<T>* sm = reinterpret_cast<cx<T>*>(base + EXCH_OFF) + (size_t)c * SLOT;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + BAR_OFF);
  T* PDSP_RESTRICT ore = static_cast<T*>(p.out_re);
  T* PDSP_RESTRICT oim = static_cast<T*>(p.out_im);
  cx<T>* PDSP_RESTRICT ocx = static_cast<cx<T>*>(p.out_re);  // when p.out_cplx
  const T scale = (T)p.scale;
  constexpr bool in_cplx = (IO & 1) != 0;  // tm_re then maps the interleaved buffer as [rows][2*I], {2*C, BOX_ROWS/2} box
  constexpr bool OUT_CPLX = (IO & 2) != 0;
  const bool has_im = p.in_im != nullptr && !in_cplx;
  const long long total = p.n_groups * p.n_frames;
  const long long n_hi = p.n_groups / p.n_lo;  // O

  auto prefetch = [&](long long w) {  // one thread: L2 prefetch of a tile's boxes
    const long long fb = w / p.n_groups, g = w % p.n_groups;
    const long long g_hi = g / p.n_lo, g_lo = g % p.n_lo;
    const int x = (int)(g_lo * C);
    const int y = (int)((fb * n_hi + g_hi) * L);
    if constexpr (in_cplx) {
      for (int j = 0; j < 2 * NBOX; ++j) simt::tma_prefetch_2d(&tm_re, 2 * x, y + j * (BOX_ROWS / 2));
      return;
    }
    for (int j = 0; j < NBOX; ++j) {
      simt::tma_prefetch_2d(&tm_re, x, y + j * BOX_ROWS);
      if (has_im) simt::tma_prefetch_2d(&tm_im, x, y + j * BOX_ROWS);
    }
  };
  auto issue = [&](long long w, int stage) {  // one thread: arm the barrier and launch the tile's box loads
    const long long fb = w / p.n_groups, g = w % p.n_groups;
    const long long g_hi = g / p.n_lo, g_lo = g % p.n_lo;
    const int x = (int)(g_lo * C);
    const int y = (int)((fb * n_hi + g_hi) * L);
    simt::mbar_expect_tx(&bars[stage], (unsigned)((has_im || in_cplx ? 2 : 1) * PLANE));
    unsigned char* tile = base + (size_t)stage * 2 * PLANE;
    if constexpr (in_cplx) {  // interleaved rows: the tile is [L][C] cx<T>, fetched as 2*NBOX boxes of BOX_ROWS/2 rows
      for (int j = 0; j < 2 * NBOX; ++j)
        simt::tma_load_2d(tile + (size_t)j * (BOX_ROWS / 2) * C * sizeof(cx<T>), &tm_re, 2 * x, y + j * (BOX_ROWS / 2),
                          &bars[stage]);
      return;
    }
    for (int j = 0; j < NBOX; ++j) {
      simt::tma_load_2d(tile + (size_t)j * BOX_ROWS * C * sizeof(T), &tm_re, x, y + j * BOX_ROWS, &bars[stage]);
      if (has_im)
        simt::tma_load_2d(tile + PLANE + (size_t)j * BOX_ROWS * C * sizeof(T), &tm_im, x, y + j * BOX_ROWS, &bars[stage]);
    }
  };

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) simt::mbar_init(&bars[s], 1);
  }
  simt::sync_block();
  unsigned phase[2] = {0u, 0u};
  if (STAGES == 2 && tid == 0 && simt::bid() < total) issue(simt::bid(), 0);

  long long it = 0;
  for (long long w = simt::bid(); w < total; w += simt::nblocks(), ++it) {
    const int stage = STAGES == 2 ? (int)(it & 1) : 0;
    if (STAGES == 2) {
      // prefetch the next tile into the other stage (its previous contents were consumed one iteration ago)
      if (tid == 0 && w + simt::nblocks() < total) issue(w + simt::nblocks(), stage ^ 1);
    } else {
      // the tile shares memory with the exchange buffers of the previous item: order those generic-proxy
      // accesses before the async-proxy writes of the bulk copy
      simt::fence_proxy_async();
      simt::sync_block();
      if (tid == 0) {
        issue(w, 0);
        if (p.l2_prefetch && w + simt::nblocks() < total) prefetch(w + simt::nblocks());
      }
    }
    simt::mbar_wait(&bars[stage], phase[stage]);
    phase[stage] ^= 1u;

    const T* tre = reinterpret_cast<const T*>(base + (size_t)stage * 2 * PLANE);
    const T* tim = reinterpret_cast<const T*>(base + (size_t)stage * 2 * PLANE + PLANE);
    cx<T> v[P];
    static_for<0, P>([&](auto qi) {
      constexpr int q = decltype(qi)::value;
      const int e = t + TF * q;
      if constexpr (in_cplx) {
        v[q] = reinterpret_cast<const cx<T>*>(tre)[e * C + c];
      } else {
        const T re = tre[e * C + c];
        const T im = has_im ? tim[e * C + c] : (T)0;
        v[q] = p.swap_in ? cx<T>{im, re} : cx<T>{re, im};
      }
    });
    simt::sync_block();  // tile consumed: it may be refilled (STAGES 2) / overwritten by the exchanges (STAGES 1)

    E::template fft<true>(v, t, sm, static_cast<const cx<T>*>(p.tw), 0, 1);

    const long long fb = w / p.n_groups, g = w % p.n_groups;
    const long long g_hi = g / p.n_lo, g_lo = g % p.n_lo;
    if (p.tw_hi != nullptr) {
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

// shared-memory plan of the TMA kernel for (T, LOG2L, C)
template <typename T, int LOG2L, int LOG2P, int MAXRB, int C>
struct BigTmaSmem {
  using E = FftEngine<T, LOG2L, LOG2P, MAXRB>;
  static constexpr size_t PLANE = sizeof(T) * (size_t)E::M * C;
  static constexpr size_t EXCH = sizeof(cx<T>) * (size_t)(E::SMEM_ELEMS | 1) * C;
  // two stages only while two CTAs still fit an SM (more resident warps beat a deeper pipeline here:
  // profiles/r1/README.md), otherwise one stage aliased onto the exchange buffers
  static constexpr int STAGES = (4 * PLANE + EXCH + 512 <= 100 * 1024) ? 2 : 1;
  static constexpr size_t BYTES = (STAGES == 2 ? 4 * PLANE + EXCH : (2 * PLANE > EXCH ? 2 * PLANE : EXCH)) + 16 + 32 + 128;
};

// Pass configuration for a sub-transform length 2^LOG2L.
template <typename T, int LOG2L>
struct BigCfg {
#ifndef PDSP_BIG_P16_FROM
#define PDSP_BIG_P16_FROM 8  // smallest log2(L) whose threads hold 16 points (8 below)
#endif
  static constexpr int LOG2P = LOG2L >= PDSP_BIG_P16_FROM ? 4 : 3;
  static constexpr int MAXRB = LOG2P == 4 ? 4 : 3;
  static constexpr int TF = (1 << LOG2L) >> LOG2P;
  // sequences per CTA = contiguous elements per row of the strided tile: as many as 512 threads and
  // 148 KB of exchange buffers allow (32 -> 256-byte rows of doubles; DRAM efficiency of the column
  // gather grows with the row length, profiles/r1/README.md)
#ifndef PDSP_BIG_C_SMALL
#define PDSP_BIG_C_SMALL 32
#endif
#ifndef PDSP_BIG_C_L10
#define PDSP_BIG_C_L10 8  // experiment: 4 = half tiles, 256-thread CTAs, two per SM
#endif
  static constexpr int C = LOG2L == 10 ? PDSP_BIG_C_L10 : (LOG2L == 9 ? 16 : PDSP_BIG_C_SMALL);
};
constexpr int kBigMinLog2L = 6, kBigMaxLog2L = 10;

}  // namespace pdsp
