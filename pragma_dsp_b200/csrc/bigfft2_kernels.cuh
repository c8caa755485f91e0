// bigfft2_kernels.cuh - K2, second generation: one pass of the multi-pass large-N FFT (N > 8192 complex points;
// BASELINE config C4: N = 2^20, 2^24) with TMA on BOTH sides of the tile.
//
// N = L1*L2(*L3) as in bigfft_kernels.cuh.  A pass transforms the axis of length L of a 3-D view of the data; a CTA works
// on a tile of C adjacent sequences:
//
//   * column passes (every pass but the last): the tile is C adjacent columns of the view [O][L][I] - fetched by
//     cp.async.bulk.tensor box loads into shared memory, transformed in place (the landing zone is also the exchange
//     buffer of the Stockham engine), multiplied by the Cooley-Tukey twiddle, written back to shared memory in tile
//     order and sent to global memory by cp.async.bulk.tensor box STORES (SASS UTMASTG);
//   * the last pass reads C contiguous rows with coalesced per-thread loads (its input was just written by the previous
//     pass: L2 hits) and writes its output TRANSPOSED (digit-reversed: X[k1 + L1*k2 (+ L1*L2*k3)]) with the same box
//     stores - the transpose of the four-step algorithm is a TMA store of a {C, rows} box, not a scatter of C*8-byte
//     segments from the LSU.
//
// Global traffic no longer passes through the LSU (the first generation spent 2/7 of its L1 wavefronts on it); tiles are
// half the size of the first generation's (C = 4 for L = 1024, 16 for L = 256), so that TWO CTAs share an SM and the
// memory phases of one overlap the arithmetic of the other; stores drain asynchronously while the CTA loads its next
// tile.  The reference runs the same transform as log2(N) strided in-place sweeps (/root/reference/src/core/fft.ts:116-140).
#pragma once
#include "bigfft_kernels.cuh"

namespace pdsp {

struct BigTileParams {
  long long n_groups;  // tiles per transform; group g = g_hi * n_lo + g_lo
  long long n_lo;
  long long n_frames;  // transforms in this launch (work item = frame * n_groups + g)
  // TMA coordinates of a tile's first box: coord[d] = g_lo * c_lo[d] + g_hi * c_hi[d] + frame * c_fr[d]; successive
  // boxes of the tile advance dimension box_dim by box_step
  int in_lo[3], in_hi[3], in_fr[3], in_box_dim, in_box_step, in_boxes;
  int out_lo[3], out_hi[3], out_fr[3], out_box_dim, out_box_step, out_boxes;
  unsigned in_box_bytes, out_box_bytes;  // bytes one box moves (per plane)
  // last pass only: per-thread staged row loads (element strides as in BigPassParams)
  const void* in_re;
  const void* in_im;
  long long in_frame, in_g_hi, in_g_lo, in_c;
  const void* tw;     // per-pass twiddles of the L-point schedule (set by the launcher)
  const void* tw_hi;  // two-level inter-pass twiddle, null on the last pass
  const void* tw_lo;
  int log_b;
  int swap_in, swap_out;  // inverse transform = swap(FFT(swap(x))) / N
  double scale;
  int has_im;  // planar input with an imaginary plane (0: real input, zeros)
  int l2_prefetch;  // 1: pull the CTA's next tile into L2 while the current one is transformed (the mbarrier wait on
                    // the tile load was half of all stall samples without it: profiles/r2/ncu_digest_c4_2e24.txt)
};

// IO bit 0: input is the interleaved work buffer; bit 1: output is; bit 2: last pass (row tile in, transposed box out)
//
// One work item (tile) of a pass.  `w` is the item's index in the pass' own list, `w_next` the item of the SAME pass this CTA
// will take next (or -1: nothing to prefetch), `phase` the CTA's mbarrier phase bit (carried across items of any kind).
struct BigNoHook {
  PDSP_DEVICE void operator()() const {}
};
// after_load(): called by every thread once the item's input is in registers (the fused kernel publishes the previous
// item's stores there: by then they have long completed, so waiting for them costs nothing)
template <typename T, int LOG2L, int LOG2P, int MAXRB, int C, int IO, class Hook = BigNoHook>
PDSP_DEVICE void big_tile_item(const BigTileParams& p, const simt::TensorMap* tm_in_re, const simt::TensorMap* tm_in_im,
                               const simt::TensorMap* tm_out_re, const simt::TensorMap* tm_out_im, unsigned char* base,
                               unsigned long long* bar, unsigned& phase, long long w, long long w_next, Hook after_load = Hook{}) {
  constexpr bool IN_CPLX = (IO & 1) != 0, OUT_CPLX = (IO & 2) != 0, LAST = (IO & 4) != 0;
  // bit 3 (last pass of the fused kernel): the input rows were written by other CTAs of this very launch - read them
  // through the L2 (ld.global.cg), not through the non-coherent path.  Compile time: a run-time test inside the unrolled
  // row loads is what cost the first generation 25-50 %.
  constexpr bool COHERENT_IN = (IO & 8) != 0;
  using E = FftEngine<T, LOG2L, LOG2P, MAXRB>;
  constexpr int L = E::M, P = E::P, TF = E::TF;
  constexpr int THREADS = TF * C;
  constexpr int SLOT = E::SMEM_ELEMS | 1;  // odd stride: neighbouring sequences start in neighbouring banks
  constexpr size_t PLANE = sizeof(T) * (size_t)L * C;
  const int tid = simt::tid();
  const int c = tid % C;
  const int t = tid / C;
  cx<T>* smem = reinterpret_cast<cx<T>*>(base);
  cx<T>* sm = smem + (size_t)c * SLOT;
  const T scale = (T)p.scale;
  {
    const long long fb = w / p.n_groups, g = w % p.n_groups;
    const long long g_hi = g / p.n_lo, g_lo = g % p.n_lo;
    cx<T> v[P];
    if constexpr (!LAST) {
      // ---- tile in: TMA box loads into the (free) region.  The previous item's stores must have finished READING it.
      if (tid == 0) {
        simt::bulk_wait_read();
        simt::fence_proxy_async();
        int co[3];
        for (int d = 0; d < 3; ++d) co[d] = (int)(g_lo * p.in_lo[d] + g_hi * p.in_hi[d] + fb * p.in_fr[d]);
        const int planes = IN_CPLX ? 1 : (p.has_im ? 2 : 1);
        simt::mbar_expect_tx(bar, (unsigned)planes * p.in_box_bytes * (unsigned)p.in_boxes);
        for (int j = 0; j < p.in_boxes; ++j) {
          simt::tma_load_3d(base + (size_t)j * p.in_box_bytes, tm_in_re, co[0], co[1], co[2], bar);
          if (!IN_CPLX && p.has_im) simt::tma_load_3d(base + PLANE + (size_t)j * p.in_box_bytes, tm_in_im, co[0], co[1], co[2], bar);
          co[p.in_box_dim] += p.in_box_step;
        }
        if (p.l2_prefetch && w_next >= 0) {
          const long long fb2 = w_next / p.n_groups, g2 = w_next % p.n_groups;
          const long long h2 = g2 / p.n_lo, l2 = g2 % p.n_lo;
          for (int d = 0; d < 3; ++d) co[d] = (int)(l2 * p.in_lo[d] + h2 * p.in_hi[d] + fb2 * p.in_fr[d]);
          for (int j = 0; j < p.in_boxes; ++j) {
            simt::tma_prefetch_3d(tm_in_re, co[0], co[1], co[2]);
            if (!IN_CPLX && p.has_im) simt::tma_prefetch_3d(tm_in_im, co[0], co[1], co[2]);
            co[p.in_box_dim] += p.in_box_step;
          }
        }
      }
      simt::mbar_wait(bar, phase);
      phase ^= 1u;
      const T* tre = reinterpret_cast<const T*>(base);
      const T* tim = reinterpret_cast<const T*>(base + PLANE);
      static_for<0, P>([&](auto qi) {
        constexpr int q = decltype(qi)::value;
        const int e = t + TF * q;
        if constexpr (IN_CPLX) {
          v[q] = reinterpret_cast<const cx<T>*>(tre)[e * C + c];
        } else {
          const T re = tre[e * C + c];
          const T im = p.has_im ? tim[e * C + c] : (T)0;
          v[q] = p.swap_in ? cx<T>{im, re} : cx<T>{re, im};
        }
      });
      simt::sync_block();  // tile consumed: the exchanges may overwrite it
      after_load();
    } else {
      // ---- last pass: C contiguous rows, element index fastest across the CTA, parked in the padded exchange layout.
      // (the region may still be read by the previous item's stores)
      if (tid == 0) simt::bulk_wait_read();
      simt::sync_block();
      const long long in_base = fb * p.in_frame + g_hi * p.in_g_hi + g_lo * p.in_g_lo;
      const T* PDSP_RESTRICT ire = static_cast<const T*>(p.in_re);
      const T* PDSP_RESTRICT iim = static_cast<const T*>(p.in_im);
      const cx<T>* PDSP_RESTRICT icx = static_cast<const cx<T>*>(p.in_re);
      PDSP_UNROLL
      for (int k = 0; k < P; ++k) {
        const int idx = tid + k * THREADS;
        const int cc = idx >> LOG2L, e = idx & (L - 1);
        const long long a = in_base + cc * p.in_c + e;
        if constexpr (IN_CPLX) {
          smem[(size_t)cc * SLOT + E::pad(e)] = COHERENT_IN ? ldcg_cx(icx + a) : ldg_cx(icx + a);
        } else {
          const T re = ire[a];
          const T im = iim != nullptr ? iim[a] : (T)0;
          smem[(size_t)cc * SLOT + E::pad(e)] = p.swap_in ? cx<T>{im, re} : cx<T>{re, im};
        }
      }
      simt::sync_block();
      static_for<0, P>([&](auto q) { v[decltype(q)::value] = sm[E::pad(t + TF * decltype(q)::value)]; });
      if (p.l2_prefetch && w_next >= 0 && tid < C) {
        // the next tile's rows (C runs of L contiguous elements), one bulk L2 prefetch per row
        const long long fb2 = w_next / p.n_groups, g2 = w_next % p.n_groups;
        const long long a2 = fb2 * p.in_frame + (g2 / p.n_lo) * p.in_g_hi + (g2 % p.n_lo) * p.in_g_lo + tid * p.in_c;
        const unsigned row_bytes = (unsigned)(L * sizeof(T)) * (IN_CPLX ? 2u : 1u);
        if constexpr (IN_CPLX) {
          simt::prefetch_l2_bulk(icx + a2, row_bytes);
        } else {
          simt::prefetch_l2_bulk(ire + a2, row_bytes);
          if (iim != nullptr) simt::prefetch_l2_bulk(iim + a2, row_bytes);
        }
      }
      simt::sync_block();
      after_load();
    }

    E::template fft<true>(v, t, sm, static_cast<const cx<T>*>(p.tw), 0, 1);

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

    // ---- tile out: results to shared memory in tile order [k][c] (the engine's last pass left the region free), then
    // box stores by one thread.  They drain while this CTA fetches its next tile.
    simt::sync_block();  // every thread is past the last exchange read
    {
      T* sre = reinterpret_cast<T*>(base);
      T* sim = reinterpret_cast<T*>(base + PLANE);
      static_for<0, P>([&](auto qi) {
        constexpr int q = decltype(qi)::value;
        const int k = t + TF * q;
        if constexpr (OUT_CPLX) {
          reinterpret_cast<cx<T>*>(sre)[k * C + c] = v[q];
        } else {
          const T x = v[q].x * scale, y = v[q].y * scale;
          sre[k * C + c] = p.swap_out ? y : x;
          sim[k * C + c] = p.swap_out ? x : y;
        }
      });
    }
    simt::fence_proxy_async();  // generic-proxy writes before the async-proxy reads of the bulk stores
    simt::sync_block();
    if (tid == 0) {
      int co[3];
      for (int d = 0; d < 3; ++d) co[d] = (int)(g_lo * p.out_lo[d] + g_hi * p.out_hi[d] + fb * p.out_fr[d]);
      for (int j = 0; j < p.out_boxes; ++j) {
        simt::tma_store_3d(tm_out_re, co[0], co[1], co[2], base + (size_t)j * p.out_box_bytes);
        if (!OUT_CPLX) simt::tma_store_3d(tm_out_im, co[0], co[1], co[2], base + PLANE + (size_t)j * p.out_box_bytes);
        co[p.out_box_dim] += p.out_box_step;
      }
      simt::bulk_commit();
    }
  }
}

// shared-memory carve-up shared by the kernels below: the 128-byte aligned region and the mbarrier behind it
template <typename T, int LOG2L, int LOG2P, int MAXRB, int C>
struct BigTileSmem {
  using E = FftEngine<T, LOG2L, LOG2P, MAXRB>;
  static constexpr size_t PLANE = sizeof(T) * (size_t)E::M * C;
  static constexpr size_t EXCH = sizeof(cx<T>) * (size_t)(E::SMEM_ELEMS | 1) * C;
  static constexpr size_t REGION = EXCH > 2 * PLANE ? EXCH : 2 * PLANE;
  static constexpr size_t BAR_OFF = (REGION + 15) & ~(size_t)15;
};

template <typename T, int LOG2L, int LOG2P, int MAXRB, int C, int IO>
PDSP_GLOBAL void PDSP_LAUNCH_BOUNDS(((1 << LOG2L) >> LOG2P) * C, 2)
    bigfft_tile_kernel(const BigTileParams p, const PDSP_GRID_CONSTANT simt::TensorMap tm_in_re,
                       const PDSP_GRID_CONSTANT simt::TensorMap tm_in_im, const PDSP_GRID_CONSTANT simt::TensorMap tm_out_re,
                       const PDSP_GRID_CONSTANT simt::TensorMap tm_out_im) {
  using S = BigTileSmem<T, LOG2L, LOG2P, MAXRB, C>;
  const int tid = simt::tid();
  // one 128-byte aligned region: landing zone of the tile, exchange buffer, staging of the outgoing tile
  unsigned char* base = simt::smem();
  base += (128 - (reinterpret_cast<uintptr_t>(base) & 127)) & 127;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(base + S::BAR_OFF);
  const long long total = p.n_groups * p.n_frames;
  if (tid == 0) simt::mbar_init(bar, 1);
  simt::sync_block();
  unsigned phase = 0u;
  for (long long w = simt::bid(); w < total; w += simt::nblocks()) {
    const long long w_next = w + simt::nblocks() < total ? w + simt::nblocks() : -1;
    big_tile_item<T, LOG2L, LOG2P, MAXRB, C, IO>(p, &tm_in_re, &tm_in_im, &tm_out_re, &tm_out_im, base, bar, phase, w, w_next);
  }
  if (tid == 0) simt::bulk_wait_all();  // global writes complete before the kernel ends
}

// ---- Fused passes 1 + 2 of a three-pass transform in ONE persistent launch.
// The last pass works on tiles of C2 rows with adjacent k1 (so that its transposed output leaves as C2*8-byte segments),
// i.e. it needs the middle pass' results of C2 whole k1 blocks - a "k1 group".  The fused work list alternates, group by
// group, the group's n1g middle-pass tiles and its n2g last-pass tiles; CTAs take items in order (bid, bid + grid, ...), a
// last-pass item first waits until the group's counter shows all n1g middle tiles stored.  A group's slice of the work
// buffer (C2 k1 blocks: 16 MB for 2^24) is then read back while it is still in the L2: the work buffer crosses HBM once
// between passes 0 and 1 only, with none of the ramp and tail of the launch-per-group form (profiles/r2/README.md).
// No deadlock: every producer of an item precedes it in the list, and the grid is never larger than what is resident.
struct BigFusedSync {
  unsigned* counters;   // one per k1 group, zeroed before the launch
  unsigned* error;      // set when a wait gave up (watchdog) - the host reports it
  long long n_groups;   // k1 groups
  long long n1g, n2g;   // middle-pass / last-pass items per group
  long long skew;       // list order: a group's last-pass tiles follow its middle tiles by this many groups of middle tiles (>= 1)
};
template <typename T, int LOG2LA, int LOG2LB, int LOG2P, int MAXRBA, int MAXRBB, int CA, int CB>
PDSP_GLOBAL void PDSP_LAUNCH_BOUNDS(((1 << LOG2LA) >> LOG2P) * CA, 2)
    bigfft_fused_kernel(const BigTileParams pa, const BigTileParams pb, const BigFusedSync fs,
                        const PDSP_GRID_CONSTANT simt::TensorMap a_in, const PDSP_GRID_CONSTANT simt::TensorMap a_out,
                        const PDSP_GRID_CONSTANT simt::TensorMap b_out_re, const PDSP_GRID_CONSTANT simt::TensorMap b_out_im) {
  static_assert((((1 << LOG2LA) >> LOG2P) * CA) == (((1 << LOG2LB) >> LOG2P) * CB), "both passes run on the same CTA shape");
  using SA = BigTileSmem<T, LOG2LA, LOG2P, MAXRBA, CA>;
  using SB = BigTileSmem<T, LOG2LB, LOG2P, MAXRBB, CB>;
  constexpr size_t BAR_OFF = SA::BAR_OFF > SB::BAR_OFF ? SA::BAR_OFF : SB::BAR_OFF;
  const int tid = simt::tid();
  unsigned char* base = simt::smem();
  base += (128 - (reinterpret_cast<uintptr_t>(base) & 127)) & 127;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(base + BAR_OFF);
  const long long per_group = fs.n1g + fs.n2g;
  const long long total = fs.n_groups * per_group;
  if (tid == 0) simt::mbar_init(bar, 1);
  simt::sync_block();
  unsigned phase = 0u;
  long long pending = -1;  // thread 0: group whose counter still owes this CTA's last middle-pass tile
  // List order: M0, M1, L0, M2, L1, ..., M(n-1), L(n-2), L(n-1) (M = a group's middle tiles, L = its last-pass tiles): a
  // last-pass item then follows its producers by a whole group of other work, so in steady state nobody waits (with
  // M0, L0, M1, L1, ... a third of the CTAs sat out a tile time in every round: 2^24 430 us against 327 pass by pass).
  auto decode = [&](long long f, long long* G, long long* r) -> bool {  // true: middle-pass item
    const long long S = fs.skew < fs.n_groups ? fs.skew : fs.n_groups;  // groups of middle tiles ahead of the last-pass tiles
    if (f < S * fs.n1g) {
      *G = f / fs.n1g, *r = f % fs.n1g;
      return true;
    }
    const long long fp = f - S * fs.n1g, pair = fp / per_group, rem = fp % per_group;
    if (pair >= fs.n_groups - S) {  // tail: the last S groups' last-pass tiles
      const long long ft = fp - (fs.n_groups - S) * per_group;
      *G = fs.n_groups - S + ft / fs.n2g, *r = ft % fs.n2g;
      return false;
    }
    if (rem < fs.n1g) {
      *G = pair + S, *r = rem;
      return true;
    }
    *G = pair, *r = rem - fs.n1g;
    return false;
  };
  for (long long f = simt::bid(); f < total; f += simt::nblocks()) {
    long long G, r;
    const bool middle = decode(f, &G, &r);
    // publish this CTA's previous middle tile once its box stores have completed (release).  Done from inside the item,
    // after its input has arrived: the stores have had a whole load latency to drain, so the wait is free - at the top
    // of the item it exposed the drain of every middle tile.  (No deadlock: the wait below is on an EARLIER group's
    // counter than the one this CTA still owes, and every CTA reaches its hook without waiting on anyone but producers
    // of earlier groups.)
    auto publish = [&]() {
      if (tid == 0 && pending >= 0) {
        simt::bulk_wait_all();
        simt::fence_proxy_async_all();
        simt::fence_gpu();
        simt::atomic_add(fs.counters + pending, 1u);
        pending = -1;
      }
    };
    if (tid == 0) {
      if (!middle && pending == G) publish();  // owing the very group it is about to wait for: settle first
      if (!middle) {  // last-pass item: the group's middle tiles must all be in (acquire)
        unsigned long long spins = 0;
        while (simt::load_acquire(fs.counters + G) < (unsigned)fs.n1g) {
          if (++spins > (1ull << 22)) {  // watchdog (seconds): never hang the device on a lost signal
            simt::store_volatile(fs.error, 1u);
            break;
          }
        }
      }
    }
    if (middle) {
      const long long w = G * fs.n1g + r;
      // next middle-pass item of this CTA (for the L2 prefetch of its tile), if its next item is one
      // the CTA's next MIDDLE item, whatever lies between (its tile comes from HBM: have it waiting in the L2)
      long long w2 = -1;
      for (long long f2 = f + simt::nblocks(), k = 0; f2 < total && k < 4; f2 += simt::nblocks(), ++k) {
        long long G2, r2;
        if (decode(f2, &G2, &r2)) {
          w2 = G2 * fs.n1g + r2;
          break;
        }
      }
      big_tile_item<T, LOG2LA, LOG2P, MAXRBA, CA, 3>(pa, &a_in, &a_in, &a_out, &a_out, base, bar, phase, w, w2, publish);
      if (tid == 0) pending = G;
    } else {
      const long long w = G * fs.n2g + r;
      big_tile_item<T, LOG2LB, LOG2P, MAXRBB, CB, 5 | 8>(pb, &b_out_re, &b_out_im, &b_out_re, &b_out_im, base, bar, phase, w, -1, publish);
    }
  }
  if (tid == 0) {
    simt::bulk_wait_all();  // global writes complete before the kernel ends
    if (pending >= 0) {
      simt::fence_proxy_async_all();
      simt::fence_gpu();
      simt::atomic_add(fs.counters + pending, 1u);
    }
  }
}

// Tile configuration of the second generation: half the sequences of BigCfg, two CTAs per SM
template <typename T, int LOG2L>
struct BigCfg2 {
  static constexpr int LOG2P = BigCfg<T, LOG2L>::LOG2P;
  static constexpr int MAXRB = BigCfg<T, LOG2L>::MAXRB;
  static constexpr int TF = (1 << LOG2L) >> LOG2P;
#ifndef PDSP_BIG2_THREADS
#define PDSP_BIG2_THREADS 256
#endif
  static constexpr int C = PDSP_BIG2_THREADS / TF > 0 ? PDSP_BIG2_THREADS / TF : 1;
  using E = FftEngine<T, LOG2L, LOG2P, MAXRB>;
  static constexpr size_t PLANE = sizeof(T) * (size_t)E::M * C;
  static constexpr size_t EXCH = sizeof(cx<T>) * (size_t)(E::SMEM_ELEMS | 1) * C;
  static constexpr size_t SMEM = (((EXCH > 2 * PLANE ? EXCH : 2 * PLANE) + 15) & ~(size_t)15) + 16 + 128;
};

}  // namespace pdsp
