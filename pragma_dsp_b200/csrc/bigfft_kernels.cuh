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
};

template <typename T>
PDSP_DEVICE cx<T> big_twiddle(const cx<T>* PDSP_RESTRICT hi, const cx<T>* PDSP_RESTRICT lo, long long m, int log_b) {
  const cx<T> a = ldg_cx(hi + (m >> log_b));
  const cx<T> b = ldg_cx(lo + (m & ((1LL << log_b) - 1)));
  return cmul(a, b);
}

template <typename T, int LOG2L, int LOG2P, int MAXRB, int C>
PDSP_GLOBAL void PDSP_LAUNCH_BOUNDS(((1 << LOG2L) >> LOG2P) * C, 1) bigfft_pass_kernel(const BigPassParams p) {
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
        const T re = ire[a];
        const T im = iim != nullptr ? iim[a] : (T)0;
        smem[(size_t)cc * SLOT + E::pad(e)] = p.swap_in ? cx<T>{im, re} : cx<T>{re, im};
      }
      simt::sync_block();
      static_for<0, P>([&](auto q) { v[decltype(q)::value] = sm[E::pad(t + TF * decltype(q)::value)]; });
      simt::sync_block();
    } else {
      static_for<0, P>([&](auto qi) {
        constexpr int q = decltype(qi)::value;
        const long long a = in_base + c * p.in_c + (long long)(t + TF * q) * p.in_e;
        const T re = ire[a];
        const T im = iim != nullptr ? iim[a] : (T)0;
        v[q] = p.swap_in ? cx<T>{im, re} : cx<T>{re, im};
      });
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
      const T x = v[q].x * scale, y = v[q].y * scale;
      ore[a] = p.swap_out ? y : x;
      oim[a] = p.swap_out ? x : y;
    });
  }
}

// Pass configuration for a sub-transform length 2^LOG2L.
template <typename T, int LOG2L>
struct BigCfg {
  static constexpr int LOG2P = LOG2L >= 9 ? 4 : 3;
  static constexpr int MAXRB = 3;
  static constexpr int TF = (1 << LOG2L) >> LOG2P;
  // sequences per CTA: 16 (128-byte segments of doubles); 8 for L = 1024 (512 threads, <= 148 KB smem)
  static constexpr int C = LOG2L == 10 ? 8 : 16;
};
constexpr int kBigMinLog2L = 6, kBigMaxLog2L = 10;

}  // namespace pdsp
