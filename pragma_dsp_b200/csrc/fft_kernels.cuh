// fft_kernels.cuh - the batched hot-path kernels built on FftEngine.
//
//   r2c_kernel : one launch = buildFrame (zero-pad/truncate, f32->compute widen) + applyWindow +
//                Radix2Fft.forward on a real frame + magnitude/phase + amplitude scaling + findPeak
//                i.e. the whole of spectrum() (/root/reference/src/public/spectrum.ts:107-142) and of
//                Radix2Fft.forward (/root/reference/src/core/fft.ts:77-79) for a batch of frames.
//                The N real samples are packed as M = N/2 complex points, transformed, and split
//                with the Hermitian post-pass; the partner bin Z[M-k] comes from a warp shuffle
//                when a frame fits a warp, from shared memory otherwise.
//   c2c_kernel : Radix2Fft.forwardComplex / inverse (/root/reference/src/core/fft.ts:81-87) on planar
//                (split real/imag) arrays - the reference's ComplexArray layout (:1-4).
#pragma once
#include "fft_core.cuh"

namespace pdsp {

enum : int { DT_F32 = 0, DT_F64 = 1 };

template <typename T>
struct PeakRec;
template <>
struct PeakRec<double> {  // 32 B
  int32_t index;
  int32_t pad;
  double frequency, amplitude, phase;
};
template <>
struct PeakRec<float> {  // 16 B
  int32_t index;
  float frequency, amplitude, phase;
};

// Completion doorbell of a launch (null flag = none): every CTA makes its results visible system-wide and bumps `count`
// (device memory, zero between launches); the last one resets it and stores `seq` to `flag`, a word in host-mapped pinned
// memory the host spins on.  Used by the single-call fast path (Radix2Fft.forward / spectrum() on one frame), where a
// stream synchronisation would cost as much as the transform.
struct Doorbell {
  unsigned* flag;
  unsigned* count;
  unsigned seq;
};
PDSP_DEVICE void ring_doorbell(const Doorbell& d) {
  if (d.flag == nullptr) return;
  simt::fence_system();  // this thread's results are visible to the host before anything ordered after the barrier
  simt::sync_block();
  if (simt::tid() == 0) {
    // one-CTA launches (a single frame) ring at once; otherwise the CTA that completes the count does.  Every CTA's
    // fence precedes its increment, so when the last increment is observed all results are out (fence cumulativity).
    const unsigned nb = (unsigned)simt::nblocks();
    if (nb == 1u || simt::atomic_add(d.count, 1u) == nb - 1u) {
      if (nb != 1u) *d.count = 0u;  // device memory, next used by a later launch on the same stream
      simt::store_volatile(d.flag, d.seq);
    }
  }
}

struct R2CParams {
  // input: frame f starts at samples + f*hop (elements), frame_len samples are valid
  const void* samples;
  int sample_dtype;  // DT_F32 / DT_F64
  int vec_ok;        // pairs (2e, 2e+1) may be loaded as one aligned vector
  int frame_len;
  long long hop;
  long long batch;
  const void* window;  // T[N] or nullptr (rect)
  const void* tw;      // cx<T>[E::TW_ELEMS]: per-pass twiddles of the M-point schedule (set by the launcher)
  const void* post;    // cx<T>[M/2+1]: (wi/2, -wr/2) of W_N^k
  // outputs, any may be null.  All in T.
  void* out_re;  // complex spectrum, planar; pitch = cbins
  void* out_im;
  int cfull;     // 1: all N bins (mirror written as conjugate), 0: N/2+1 bins
  void* amp;     // |X| * scale; pitch = bins
  void* phase;   // atan2(im, re); pitch = bins
  void* peaks;   // PeakRec<T>[batch]
  int two_sided;      // bins = N (mirror bins written) else N/2+1
  int shift;          // two-sided rows are stored fftShift-ed: bin k lands at (k + N/2) mod N (fourier.ts:122-134)
  double scale_edge;  // amplitude scale of DC and Nyquist
  double scale_mid;   // amplitude scale of every other bin
  double bin_hz;      // sampleRate / N
  // fused gather of the peak records over NVLink peer memory: every finished record of frame f is also
  // stored to peer[g] + (peer_offset + f) records, g < n_peers (mapped peer buffers, one per rank incl. self)
  void* peer[8];
  int n_peers;
  long long peer_offset;
  Doorbell door;
};

struct C2CParams {
  const void* in_re;  // T planar, frame stride N; in_im may be null (zero imaginary plane)
  const void* in_im;
  void* out_re;
  void* out_im;
  long long batch;
  const void* tw;  // cx<T>[E::TW_ELEMS]: per-pass twiddles (set by the launcher)
  int inverse;     // conjugate transform and multiply by 1/N
  Doorbell door;
};

#ifndef PDSP_FUSE_WINDOW
#define PDSP_FUSE_WINDOW 0
#endif
#ifndef PDSP_NEXT_PREFETCH
#define PDSP_NEXT_PREFETCH 0
#endif
template <typename T>
PDSP_DEVICE T t_sqrt(T v);
template <>
PDSP_DEVICE float t_sqrt<float>(float v) {
  return sqrtf(v);
}
template <>
PDSP_DEVICE double t_sqrt<double>(double v) {
  return sqrt(v);
}
// atan2 for the phase rows.  libdevice's atan2 is ~230 instructions with several slow-path branches (it
// made the STFT workload C3 compute-bound at 0.24 of the roofline); this one is branch-free except for
// a rarely taken escape to libdevice for arguments near the ends of the exponent range:
//   reduce to t = num/den with |t| <= tan(pi/8) using one division (a rotation by pi/4 is folded into the
//   choice of numerator/denominator), odd minimax polynomial in t (degree 25 for fp64: 3e-18 relative;
//   degree 11 for fp32), then undo the octant/quadrant folds.  Measured max error vs libm: see tests.
// Polynomial coefficients live in the constant bank so that each DFMA takes its coefficient as a c[][] operand;
// as literals every call site re-materialised them with two UMOVs apiece (14 % of the C3 kernel's issue slots).
#if defined(__CUDACC__) && !defined(PDSP_EMU)
static __constant__ double kAtanC[12] = {
#else
static const double kAtanC[12] = {
#endif
    0.016285756855221028, -0.034570561981427744, 0.04551593220626549,  -0.05230454270650244,
    0.05878928997834775,  -0.06666424885738255,  0.07692296375032143,  -0.09090908753500877,
    0.11111111105155447,  -0.14285714285659828,  0.19999999999999804,  -0.3333333333333333};
// branch-free core; `rare` is set when the caller must redo the value with libdevice's atan2
PDSP_DEVICE double fast_atan2_core(double y, double x, bool& rare) {
  const double ax = fabs(x), ay = fabs(y);
  const double mx = fmax(ax, ay), mn = fmin(ax, ay);
  const unsigned ex = ((unsigned)__double2hiint(mx) >> 20) & 0x7ffu;
  rare = ex - 64u >= 1920u && mx != 0.0;  // |max| outside [2^-959, 2^961) (incl. inf / NaN)
  const bool big = mn > 0.41421356237309503 * mx;          // above tan(pi/8): atan(z) = pi/4 + atan((z-1)/(z+1))
  const double num = big ? mn - mx : mn;
  const double den = big ? mn + mx : mx;
  double t;
#if defined(__CUDACC__) && !defined(PDSP_EMU)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  r = fma(fma(-den, r, 1.0), r, r);  // 2^-23 -> 2^-46
  t = num * r;
  t = fma(fma(-den, t, num), r, t);  // residual correction: ~2^-90, i.e. correctly rounded but for the last bit
#else
  t = num / den;
#endif
  if (mx == 0.0) t = 0.0;
  const double u = t * t;
  // even / odd split (two half-length Horner chains in u^2): the serial depth drops from 12 to 7 DFMAs
  const double u2 = u * u;
  double po = kAtanC[0], pe = kAtanC[1];
  po = fma(po, u2, kAtanC[2]);
  pe = fma(pe, u2, kAtanC[3]);
  po = fma(po, u2, kAtanC[4]);
  pe = fma(pe, u2, kAtanC[5]);
  po = fma(po, u2, kAtanC[6]);
  pe = fma(pe, u2, kAtanC[7]);
  po = fma(po, u2, kAtanC[8]);
  pe = fma(pe, u2, kAtanC[9]);
  po = fma(po, u2, kAtanC[10]);
  pe = fma(pe, u2, kAtanC[11]);
  const double p = fma(po, u, pe);
  double a = fma(t * u, p, t);
  if (big) a += 0.78539816339744831;
  if (ay > ax) a = 1.5707963267948966 - a;
  if (__double2hiint(x) < 0) a = 3.141592653589793 - a;  // sign bit, so that atan2(+-0, -0) = +-pi
  return copysign(a, y);
}
PDSP_DEVICE double fast_atan2(double y, double x) {
  bool rare;
  const double a = fast_atan2_core(y, x, rare);
  return rare ? atan2(y, x) : a;
}
PDSP_DEVICE float fast_atan2_core(float y, float x, bool& rare) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const unsigned ex = ((unsigned)__float_as_int(mx) >> 23) & 0xffu;
  rare = ex - 16u >= 224u && mx != 0.0f;  // |max| outside [2^-111, 2^113)
  const bool big = mn > 0.41421356f * mx;
  const float num = big ? mn - mx : mn;
  const float den = big ? mn + mx : mx;
  float t;
#if defined(__CUDACC__) && !defined(PDSP_EMU)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
  t = num * r;
  t = fmaf(fmaf(-den, t, num), r, t);
#else
  t = num / den;
#endif
  if (mx == 0.0f) t = 0.0f;
  const float u = t * t;
  float p = -0.06451927870512009f;
  p = fmaf(p, u, 0.10743731260299683f);
  p = fmaf(p, u, -0.14263956248760223f);
  p = fmaf(p, u, 0.19999539852142334f);
  p = fmaf(p, u, -0.3333333134651184f);
  float a = fmaf(t * u, p, t);
  if (big) a += 0.78539816f;
  if (ay > ax) a = 1.57079633f - a;
  if (__float_as_int(x) < 0) a = 3.14159265f - a;
  return copysignf(a, y);
}
PDSP_DEVICE float fast_atan2(float y, float x) {
  bool rare;
  const float a = fast_atan2_core(y, x, rare);
  return rare ? atan2f(y, x) : a;
}
PDSP_DEVICE_NOINLINE float t_atan2(float y, float x) { return fast_atan2(y, x); }
PDSP_DEVICE_NOINLINE double t_atan2(double y, double x) { return fast_atan2(y, x); }
// two independent arguments per call: the two dependency chains interleave (the per-bin phase of the two
// streams of a thread), which a single noinline call per bin cannot offer the scheduler
template <typename T>
struct Pair2 {
  T a, b;
};
// inlined form: the scheduler can interleave the two chains with each other and with the neighbouring bins
// (specialised fp64 kernels: C3 0.578 -> 0.559 ms); the generic kernels stay small with the out-of-line form, and the
// fp32 kernels measured 3-4 % slower inlined
template <typename T>
PDSP_DEVICE Pair2<T> atan2_x2_inline(T y0, T x0, T y1, T x1) {
  bool r0, r1;
  Pair2<T> o{fast_atan2_core(y0, x0, r0), fast_atan2_core(y1, x1, r1)};
  if (r0 || r1) o = Pair2<T>{t_atan2(y0, x0), t_atan2(y1, x1)};
  return o;
}
PDSP_DEVICE_NOINLINE Pair2<float> t_atan2_x2(float y0, float x0, float y1, float x1) { return atan2_x2_inline(y0, x0, y1, x1); }
PDSP_DEVICE_NOINLINE Pair2<double> t_atan2_x2(double y0, double x0, double y1, double x1) {
  return atan2_x2_inline(y0, x0, y1, x1);
}
PDSP_DEVICE_NOINLINE float t_hypot_slow(float x, float y) { return hypotf(x, y); }
PDSP_DEVICE_NOINLINE double t_hypot_slow(double x, double y) { return hypot(x, y); }

// sqrt without the IEEE slow path.  fp32: MUFU.SQRT (sqrt.approx, <= 1 ulp-ish, 2^-23 relative).
// fp64: MUFU.RSQ64H seed + one third-order (Halley) step, 5 DP instructions, branch- and select-free; exact 0 for +0; inf/NaN and sums outside
// the normal range are caught by the caller's exponent tracker and redone with hypot().
#if defined(__CUDACC__) && !defined(PDSP_EMU)
PDSP_DEVICE float fast_sqrt(float s) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));  // one MUFU.SQRT; denormal sums flush to 0
  return r;
}
PDSP_DEVICE double fast_sqrt(double s) {
  // MUFU.RSQ64H reads the high word only: clamping it (one integer max, ALU pipe) to the smallest normal keeps the seed
  // finite for s = +0, so that g = s * y is an exact 0 there without a select; relative error of y about 2^-20
  double y;
  const double sc = __hiloint2double(max(__double2hiint(s), 0x00100000), 0);
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(sc));
  const double g = s * y;              // ~ sqrt(s)
  const double e = fma(-g, y, 1.0);    // 1 - s*y^2
  const double q = fma(0.375, e, 0.5);
  return fma(g * e, q, g);             // g * (1 + e/2 + 3e^2/8): third-order step, error ~ e^3 < 2^-58
}
#else
PDSP_DEVICE float fast_sqrt(float s) { return sqrtf(s); }
PDSP_DEVICE double fast_sqrt(double s) { return sqrt(s); }
#endif

// |re + i*im| for one value: fast_sqrt inside the safe exponent range, hypot() outside it
PDSP_DEVICE double t_mag_checked(double re, double im) {
  const double ss = re * re + im * im;
  const unsigned hi = (unsigned)__double2hiint(ss);
  if (hi >= 0x7fd00000u || hi < 0x00400000u) return t_hypot_slow(re, im);
  return fast_sqrt(ss);
}
PDSP_DEVICE float t_mag_checked(float re, float im) { return fast_sqrt(re * re + im * im); }

// a[i] for a run-time i over a register array (a switch: only the owning lane of one frame takes one arm)
template <typename T, int NQ>
PDSP_DEVICE cx<T> pick_reg(const cx<T> (&a)[NQ], int i) {
  static_assert(NQ <= 16, "at most 32 points per thread");
  cx<T> r = a[0];
  switch (i) {
#define PDSP_PICK(J) \
  case J:            \
    if constexpr (NQ > J) r = a[J]; \
    break;
    PDSP_PICK(1) PDSP_PICK(2) PDSP_PICK(3) PDSP_PICK(4) PDSP_PICK(5) PDSP_PICK(6) PDSP_PICK(7) PDSP_PICK(8)
    PDSP_PICK(9) PDSP_PICK(10) PDSP_PICK(11) PDSP_PICK(12) PDSP_PICK(13) PDSP_PICK(14) PDSP_PICK(15)
#undef PDSP_PICK
    default:
      break;
  }
  return r;
}

// findPeak ordering (/root/reference/src/public/spectrum.ts:74-105): strict '>' scanning upward,
// so among equal values the lowest index wins; a candidate needs v > 0; NaN never wins.
template <typename T>
PDSP_DEVICE bool peak_better(T v, int k, T bv, int bk) {
  return (v > bv) || (v == bv && v > (T)0 && k < bk);
}

// Kernel specialisation.  MD_GENERIC keeps every choice a runtime (warp-uniform) flag and handles
// ragged frames (zero-pad / truncate / unaligned / tail predicates).  Any other value fixes the
// outputs at compile time and assumes whole, vector-aligned frames (frame_len >= N, vec_ok), which
// is what the batched entry points see for the BASELINE workloads; dispatch falls back to
// MD_GENERIC whenever those assumptions do not hold.
enum : int {
  MD_GENERIC = 0,
  MD_AMP = 1,     // scaled amplitude / raw magnitude rows
  MD_PHASE = 2,   // phase rows
  MD_PEAK = 4,    // per-frame findPeak record
  MD_CPLX = 8,    // complex spectrum, all N bins (Radix2Fft.forward)
  MD_TWO = 16,    // two-sided amplitude / phase rows (N bins, mirror bins written); one-sided (N/2+1) otherwise
  MD_PAD = 32,    // frames shorter than N, zero-padded (spectrum()'s default for lengths that are not a power of two)
  MD_STAGED = 64, // software-pipelined sample loads: one lane per frame fetches the NEXT frame of its slot with a bulk
                  // asynchronous copy (cp.async.bulk + mbarrier) into the slot's shared-memory buffer while the current
                  // frame is transformed; needs 16-byte aligned frames (base and hop) and sample type no wider than T
};

// Shared-memory plan of r2c_kernel: per frame slot the exchange buffer of the engine, which in staged mode doubles as the
// landing zone of the next frame's raw samples (N samples of at most sizeof(T) bytes = M complex elements) and is kept
// 16-byte aligned; staged mode appends one mbarrier per slot.
template <typename T, typename E, int MODE>
constexpr int r2c_slot_elems() {
  int e = E::NEEDS_SMEM ? E::SMEM_ELEMS : 0;
  if ((MODE & MD_STAGED) != 0) {
    if (e < E::M) e = E::M;
    while ((e * (int)sizeof(cx<T>)) % 16 != 0) ++e;
  }
  return e;
}
template <typename T, typename E, int MODE, int SLOTS>
constexpr size_t r2c_smem_bytes() {
  return sizeof(cx<T>) * (size_t)r2c_slot_elems<T, E, MODE>() * SLOTS + ((MODE & MD_STAGED) != 0 ? 8 * SLOTS + 8 : 0);
}

template <typename T, typename S>
PDSP_DEVICE cx<T> load_pair(const S* PDSP_RESTRICT s, int i0) {
  const cx<S> pr = *reinterpret_cast<const cx<S>*>(s + i0);
  return cx<T>{(T)pr.x, (T)pr.y};
}
// Sample loads from GLOBAL memory (read exactly once) with a cache policy (experiment switch PDSP_SAMPLE_LD): 0 plain
// ld.global (allocates in L1: 128 KB of samples stream through the L1 of an SM per 16 resident frames, and 6.7 % of the
// twiddle / window table sectors then miss it - profiles/r2/ncu_summary_north_star.json), 1 ld.global.cs (streaming),
// 2 ld.global.L1::no_allocate, 3 ld.global.lu (last use).
#ifndef PDSP_SAMPLE_LD
#define PDSP_SAMPLE_LD 0
#endif
#ifndef PDSP_OUT_ST
#define PDSP_OUT_ST 0  // amplitude row stores: 0 plain st.global, 1 st.global.cs (streaming)
#endif
template <typename T, typename S>
PDSP_DEVICE cx<T> load_pair_global(const S* PDSP_RESTRICT s, int i0) {
#if defined(__CUDACC__) && !defined(PDSP_EMU) && PDSP_SAMPLE_LD
  cx<S> pr;
  const cx<S>* a = reinterpret_cast<const cx<S>*>(s + i0);
#if PDSP_SAMPLE_LD == 1
#define PDSP_LDQ "ld.global.cs"
#elif PDSP_SAMPLE_LD == 2
#define PDSP_LDQ "ld.global.L1::no_allocate"
#else
#define PDSP_LDQ "ld.global.lu"
#endif
  if constexpr (sizeof(S) == 8) {
    asm volatile(PDSP_LDQ ".v2.f64 {%0, %1}, [%2];" : "=d"(pr.x), "=d"(pr.y) : "l"(a));
  } else {
    asm volatile(PDSP_LDQ ".v2.f32 {%0, %1}, [%2];" : "=f"(pr.x), "=f"(pr.y) : "l"(a));
  }
#undef PDSP_LDQ
  return cx<T>{(T)pr.x, (T)pr.y};
#else
  return load_pair<T>(s, i0);
#endif
}
template <typename T>
PDSP_DEVICE void store_stream(T* a, T v) {
#if defined(__CUDACC__) && !defined(PDSP_EMU) && PDSP_OUT_ST
  __stcs(a, v);
#else
  *a = v;
#endif
}

template <typename T, int LOG2M, int LOG2P, int MAXRB, int THREADS, int MODE>
PDSP_DEVICE void r2c_body(const R2CParams& p) {
  using E = FftEngine<T, LOG2M, LOG2P, MAXRB>;
  constexpr int M = E::M, P = E::P, TF = E::TF, N = 2 * M;
  constexpr int SLOTS = THREADS / TF;
  static_assert(THREADS % TF == 0 && SLOTS >= 1, "CTA must hold whole frames");
  constexpr bool GEN = MODE == MD_GENERIC;
  constexpr bool PHASE = GEN || (MODE & MD_PHASE) != 0;
  // Frames wider than a warp: the frame's threads are laid out so that the owners of bins k and M-k always
  // share a warp (lanes l and l^16), and the Hermitian post-pass can shuffle instead of taking a third trip
  // through shared memory.  Warp w of the frame holds logical threads 16w..16w+15 on lanes 0-15 and their
  // partners TF-16w..TF-16w-15 on lanes 16-31; t = 0 and t = TF/2 pair with themselves (warp 0, lanes 0 and 16).
  constexpr bool PAIRED = TF > 32;
  constexpr bool STAGED = !GEN && (MODE & MD_STAGED) != 0;
  static_assert(!STAGED || (MODE & MD_PAD) == 0, "staged loads fetch whole frames");
  constexpr int SLOT_ELEMS = r2c_slot_elems<T, E, MODE>();
  // The amplitude scale is a power of two (1, 1/N or 2/N): in the specialised kernels without a complex output it is
  // folded into the Hermitian post-pass constants, so X arrives scaled (bit-identical to scaling |X| afterwards, short
  // of subnormal results) and the per-bin multiply disappears; DC / Nyquist take one extra factor s_edge / s_mid.
  constexpr bool FOLD = !GEN && (MODE & MD_CPLX) == 0;
  // peak-only kernels rank bins by |X|^2 (no square root per bin); the winner's amplitude is computed
  // once, in the finishing loop.  Frames spanning several warps keep the linear key.
  constexpr bool KEYSQ = MODE == MD_PEAK && TF <= 32;
  // window multiply fused into the first butterfly stage (specialised fp64 kernels with one radix-P butterfly per thread)
  constexpr bool FUSE_WIN = PDSP_FUSE_WINDOW && !GEN && !STAGED && sizeof(T) == 8 && E::CAN_PRE0 && P >= 4;

  const int tid = simt::tid();
  const int slot = tid / TF;
  const int tl = tid % TF;  // position inside the frame's thread group
  int t = tl;               // logical thread: holds elements t + TF*q
  if constexpr (PAIRED) {
    const int w = tl >> 5, l = tl & 31;
    t = l < 16 ? 16 * w + l : (tl == 16 ? TF / 2 : TF - (16 * w + (l - 16)));
  }
  cx<T>* sm = reinterpret_cast<cx<T>*>(simt::smem()) + (size_t)slot * SLOT_ELEMS;
  const cx<T>* PDSP_RESTRICT tw = static_cast<const cx<T>*>(p.tw);
  const cx<T>* PDSP_RESTRICT post = static_cast<const cx<T>*>(p.post);
  const T* PDSP_RESTRICT win = static_cast<const T*>(p.window);
  const int lim = p.frame_len < N ? p.frame_len : N;
  const bool two_sided = GEN ? p.two_sided != 0 : (MODE & MD_TWO) != 0;
  const bool cfull = GEN ? p.cfull != 0 : true;
  const int bins = two_sided ? N : M + 1;
  const int sh = (two_sided && p.shift) ? M : 0;  // fftShift fused into the two-sided stores (row rotation by N/2)
  const int cbins = cfull ? N : M + 1;
  const T s_edge = (T)p.scale_edge, s_mid = (T)p.scale_mid;
  [[maybe_unused]] const T edge_ratio = (T)(p.scale_edge / p.scale_mid);  // 1/2 (one-sided) or 1
  [[maybe_unused]] const T edge_key = edge_ratio * edge_ratio;
  const bool want_cplx = GEN ? p.out_re != nullptr : (MODE & MD_CPLX) != 0;
  const bool want_amp = GEN ? p.amp != nullptr : (MODE & MD_AMP) != 0;
  const bool want_phase = GEN ? p.phase != nullptr : (MODE & MD_PHASE) != 0;
  const bool want_peak = GEN ? p.peaks != nullptr : (MODE & MD_PEAK) != 0;
  const bool need_mag = want_amp || want_peak;

  // Each CTA owns a contiguous run of frame groups (not a grid-stride interleave): its peak records are then
  // contiguous too, so the finishing loop below writes them - locally and to the peers over NVLink - as
  // whole 128-byte lines (at 8 GPUs the interleaved form made the kernel 43 % slower).
  const long long n_groups = (p.batch + SLOTS - 1) / SLOTS;
  const long long groups_per_cta = (n_groups + simt::nblocks() - 1) / simt::nblocks();
  const long long g_begin = (long long)simt::bid() * groups_per_cta;
  const long long g_end = g_begin + groups_per_cta < n_groups ? g_begin + groups_per_cta : n_groups;
  // fp64 kernels with one warp per frame: 32-bit trip count and a running frame index - the 64-bit (g, g_end) pair was spilled and re-read
  // every iteration and the loop test waited on that local-memory load (8 % of the north-star kernel's stall
  // samples; 0.715 -> 0.727).  The fp32 kernels keep the 64-bit form, which allocates better there (-1.4 % otherwise).
  constexpr bool IT32 = sizeof(T) == 8 && TF <= 32;  // wider fp64 frames measured 1-4 % slower with it (N=2048 all-bins forward)
  using Cnt = typename std::conditional<IT32, int, long long>::type;
  const Cnt it_begin = IT32 ? (Cnt)0 : (Cnt)g_begin;
  const Cnt n_iter = IT32 ? (Cnt)(g_end > g_begin ? g_end - g_begin : 0) : (Cnt)g_end;
  long long f_run = g_begin * SLOTS + slot;

  // ---- staged mode: the slot's buffer receives frame i+1 (one bulk asynchronous copy, issued by the frame's first lane
  // as soon as frame i's last shared-memory exchange is over) while frame i goes through the remaining passes, the
  // post-pass and the stores; an mbarrier per slot counts the bytes in.  The global-load latency that the direct form
  // exposes at the top of every iteration (a quarter of the fp64 kernel's stall cycles, profiles/r1) is off the path.
  [[maybe_unused]] unsigned long long* bar = nullptr;
  [[maybe_unused]] unsigned stage_phase = 0u;
  [[maybe_unused]] const unsigned stage_bytes = (unsigned)N * (p.sample_dtype == DT_F32 ? 4u : 8u);
  [[maybe_unused]] auto stage_issue = [&](long long fi) {  // first lane of the frame only
    const long long ff = fi < p.batch ? fi : p.batch - 1;   // a tail slot re-reads the last frame
    const char* src = static_cast<const char*>(p.samples) + (size_t)(ff * p.hop) * (p.sample_dtype == DT_F32 ? 4 : 8);
    simt::fence_proxy_async();  // the lanes' generic-proxy accesses to the buffer (ordered by the sync before this call)
    simt::mbar_expect_tx(bar, stage_bytes);
    simt::bulk_load_1d(sm, src, stage_bytes, bar);
  };
  if constexpr (STAGED) {
    bar = reinterpret_cast<unsigned long long*>(simt::smem() + ((sizeof(cx<T>) * (size_t)SLOT_ELEMS * SLOTS + 7) & ~(size_t)7)) + slot;
    if (tl == 0) simt::mbar_init(bar, 1);
    simt::sync_block();
    if (tl == 0 && it_begin < n_iter) stage_issue(IT32 ? f_run : (long long)it_begin * SLOTS + slot);
  }

  for (Cnt it = it_begin; it < n_iter; ++it, f_run += SLOTS) {
    const long long f = IT32 ? f_run : (long long)it * SLOTS + slot;
    const bool valid = f < p.batch;

    // ---- buildFrame + applyWindow fused into the load (spectrum.ts:36-43, fourier.ts:54-67)
    cx<T> v[P];
    if constexpr (!GEN) {
      // whole aligned frames: one vector load per complex point; a tail slot re-reads the last frame
      auto load_frame = [&](auto* s, auto from_global) {
        constexpr bool G = decltype(from_global)::value;
        if constexpr ((MODE & MD_PAD) != 0) {
          // buildFrame's zero padding (spectrum.ts:36-43): pairs at or beyond frame_len read as 0, the pair that
          // straddles an odd frame_len would keep its first sample (generic kernel only).  A separate compile-time mode: the same test as a
          // run-time branch around these loads cost the whole-frame kernels 2-5 %.
          // (keeping the vector loads unconditional and zeroing by selects measured far worse: 0.68 -> 0.38.  The host
          // sends odd frame lengths - one pair would straddle the end - to the generic kernel.)
          static_for<0, P>([&](auto qi) {
            const int i0 = 2 * (t + TF * decltype(qi)::value);
            cx<T> pr{(T)0, (T)0};
            if (i0 < lim) pr = G ? load_pair_global<T>(s, i0) : load_pair<T>(s, i0);
            v[decltype(qi)::value] = pr;
          });
        } else {
          static_for<0, P>([&](auto qi) {
            const int i0 = 2 * (t + TF * decltype(qi)::value);
            v[decltype(qi)::value] = G ? load_pair_global<T>(s, i0) : load_pair<T>(s, i0);
          });
        }
      };
      if constexpr (STAGED) {
        // the frame was fetched into the slot's buffer during the previous iteration (or by the prologue)
        simt::mbar_wait(bar, stage_phase);
        stage_phase ^= 1u;
        if (p.sample_dtype == DT_F32)
          load_frame(reinterpret_cast<const float*>(sm), std::false_type{});
        else
          load_frame(reinterpret_cast<const double*>(sm), std::false_type{});
        frame_sync<TF>(slot, SLOTS);  // every lane holds its samples: the exchanges may overwrite the buffer
      } else {
        const long long base = (valid ? f : p.batch - 1) * p.hop;
        if (p.sample_dtype == DT_F32)
          load_frame(static_cast<const float*>(p.samples) + base, std::true_type{});
        else
          load_frame(static_cast<const double*>(p.samples) + base, std::true_type{});
#if PDSP_NEXT_PREFETCH
        // experiment: hint the slot's NEXT frame into L1 (1) / L2 (2), one 128-byte line per lane and instruction,
        // behind this frame's own loads - the wait on those is 30 % of the north-star kernel's stall samples
        if (f + SLOTS < p.batch) {
          const size_t ssz = p.sample_dtype == DT_F32 ? 4 : 8;
          const char* nx = static_cast<const char*>(p.samples) + (size_t)((f + SLOTS) * p.hop) * ssz;
          const int lines = (int)((N * ssz) >> 7);
          for (int l = tl; l < lines; l += TF) {
            if (PDSP_NEXT_PREFETCH == 1)
              simt::prefetch_l1(nx + ((size_t)l << 7));
            else
              simt::prefetch_l2(nx + ((size_t)l << 7));
          }
        }
#endif
      }
      if constexpr (FUSE_WIN) {
        // applyWindow fused into stage 0 of the first radix-P butterfly (pairs q, q + P/2): x = v[q]*w[q], then
        // sum = fma(v[q+P/2], w[q+P/2], x) and diff = fma(-v[q+P/2], w[q+P/2], x) - 6 instead of 8 DP instructions per pair
        // of complex points (the kernel is bound by the energy of its double-precision work: profiles/r2/README.md).  A
        // rectangular window runs the same code with w = 1 (exact).
        constexpr int H = P / 2;
        static_for<0, H>([&](auto qi) {
          constexpr int q = decltype(qi)::value;
          cx<T> w0{(T)1, (T)1}, w1{(T)1, (T)1};
          if (win != nullptr) {
            w0 = ldg_cx(reinterpret_cast<const cx<T>*>(win) + (t + TF * q));
            w1 = ldg_cx(reinterpret_cast<const cx<T>*>(win) + (t + TF * (q + H)));
          }
          const cx<T> x = ew_mul(v[q], w0);
          const cx<T> y = v[q + H];
          const cx<T> sum{y.x * w1.x + x.x, y.y * w1.y + x.y};
          const cx<T> dif{x.x - y.x * w1.x, x.y - y.y * w1.y};
          v[q] = sum;
          v[q + H] = mul_w32<(q * 16) / H>(dif);
        });
      } else if (win != nullptr) {
        static_for<0, P>([&](auto qi) {
          constexpr int q = decltype(qi)::value;
          const cx<T> w = ldg_cx(reinterpret_cast<const cx<T>*>(win) + (t + TF * q));
          v[q] = ew_mul(v[q], w);
        });
      }
    } else {
      const long long base = valid ? f * p.hop : 0;
      const int lim_f = valid ? lim : 0;
      // sample type and alignment are decided once per frame, outside the unrolled loop (a warp-uniform test
      // inside it keeps the loads from being issued as one batch)
      auto load_gen = [&](auto* s, auto vec_c) {
        constexpr bool VEC = decltype(vec_c)::value;
        static_for<0, P>([&](auto qi) {
          constexpr int q = decltype(qi)::value;
          const int i0 = 2 * (t + TF * q);
          T x0 = (T)0, x1 = (T)0;
          if (VEC && i0 + 1 < lim_f) {
            const cx<T> pr = load_pair<T>(s, i0);
            x0 = pr.x;
            x1 = pr.y;
          } else {
            if (i0 < lim_f) x0 = (T)s[i0];
            if (i0 + 1 < lim_f) x1 = (T)s[i0 + 1];
          }
          v[q] = cx<T>{x0, x1};
        });
      };
      if (p.sample_dtype == DT_F32) {
        const float* s = static_cast<const float*>(p.samples) + base;
        if (p.vec_ok)
          load_gen(s, std::true_type{});
        else
          load_gen(s, std::false_type{});
      } else {
        const double* s = static_cast<const double*>(p.samples) + base;
        if (p.vec_ok)
          load_gen(s, std::true_type{});
        else
          load_gen(s, std::false_type{});
      }
      if (win != nullptr) {
        static_for<0, P>([&](auto qi) {
          constexpr int q = decltype(qi)::value;
          const cx<T> w = ldg_cx(reinterpret_cast<const cx<T>*>(win) + (t + TF * q));
          v[q].x *= w.x;
          v[q].y *= w.y;
        });
      }
    }

    // ---- M-point complex FFT of the packed frame
    if constexpr (STAGED) {
      // after the last exchange through the buffer: send for the next frame of this slot
      E::fft(v, t, sm, tw, slot, SLOTS, [&]() {
        if (tl == 0 && it + 1 < n_iter) stage_issue(f + SLOTS);
      });
    } else if constexpr (FUSE_WIN) {
      E::template fft<false, typename E::NoHook, false, true>(v, t, sm, tw, slot, SLOTS);
    } else {
      E::fft(v, t, sm, tw, slot, SLOTS);
    }

    // ---- Hermitian split: X[k], X[M-k] for the thread's pairs (and X[M/2] in thread 0), kept in registers
    // The bins of one thread form two streams: stream 0 walks k = t + TF*q upward (XA[q]), stream 1 walks M - k
    // downward (XB[q]).  The transform's registers die here; everything below - outputs, the hypot() rerun, the
    // peak record - works from XA / XB / XH.
    constexpr int NQ = M == 1 ? 1 : P / 2;
    cx<T> XA[NQ], XB[NQ], XH{(T)0, (T)0};
    if constexpr (M == 1) {
      // N = 2: X[0] = x0 + x1, X[1] = x0 - x1
      const T fs = FOLD ? s_mid : (T)1;
      XA[0] = cx<T>{(v[0].x + v[0].y) * fs, (T)0};
      XB[0] = cx<T>{(v[0].x - v[0].y) * fs, (T)0};
    } else {
      // W_N^k * (-i/2) for k = t + TF*q: one table entry per thread (k = t) times the constant
      // W_N^{TF*q} = exp(-2*pi*i*q/(2P)) when that is a 32nd root of unity, else a load per pair
      constexpr bool DERIVE = PDSP_DERIVE_POST && sizeof(T) == 8 && (16 % P) == 0 && P >= 2;  // see FftEngine::fft
      const T hs = FOLD ? (T)0.5 * s_mid : (T)0.5;                         // folded amplitude scale (a power of two)
      cx<T> post0{(T)0, (T)0};
      if constexpr (DERIVE) {
        post0 = ldg_cx(post + t);
        if constexpr (FOLD) post0 = ew_mul(post0, cx<T>{s_mid, s_mid});
      }
      static_for<0, P / 2>([&](auto qi) {
        constexpr int q = decltype(qi)::value;
        cx<T> zp;  // Z[(M - k) % M], k = t + TF*q < M/2
        if constexpr (PAIRED) {
          const cx<T> mine = v[P - 1 - q];
          cx<T> got;
          got.x = simt::shfl_xor(mine.x, 16, 32);
          got.y = simt::shfl_xor(mine.y, 16, 32);
          zp = (t == 0) ? v[(P - q) % P] : (t == TF / 2 ? mine : got);
        } else if constexpr (TF == 1) {
          zp = v[(P - q) % P];
        } else {
          const cx<T> mine = v[P - 1 - q];
          cx<T> got;
          got.x = simt::shfl(mine.x, (TF - t) & (TF - 1), TF);
          got.y = simt::shfl(mine.y, (TF - t) & (TF - 1), TF);
          zp = (t == 0) ? v[(P - q) % P] : got;
        }
        const cx<T> a = v[q];
        const cx<T> sum = ew_fma(zp, cx<T>{(T)1, (T)-1}, a);  // A + conj(Zp)
        const cx<T> dif = ew_fma(zp, cx<T>{(T)-1, (T)1}, a);  // A - conj(Zp)
        cx<T> w;                                               // (wi/2, -wr/2): W_N^k * (-i/2)
        if constexpr (DERIVE) {
          w = mul_w32<(q * 16 / P) % 16>(post0);
        } else {
          w = ldg_cx(post + t + TF * q);
          if constexpr (FOLD) w = ew_mul(w, cx<T>{s_mid, s_mid});
        }
        const cx<T> tt = cmul(dif, w);
        cx<T> xa = ew_fma(sum, cx<T>{hs, hs}, tt);                      // X[k]
        cx<T> xb = ew_fma(sum, cx<T>{hs, -hs}, cx<T>{-tt.x, tt.y});     // X[M-k] = conj(E - W*O)
        if constexpr (q == 0) {
          // DC and Nyquist of a real frame are real; the reference's imaginary parts there are +0
          // (sums of +0), so atan2 gives 0 / +pi rather than -0 / -pi.
          if (t == 0) {
            xa.y = (T)0;
            xb.y = (T)0;
          }
        }
        XA[q] = xa;
        XB[q] = xb;
      });
      // self-paired bin M/2 = conj(Z[M/2]) (thread 0; above every stream-0 bin of that thread)
      if (t == 0) XH = FOLD ? cx<T>{v[P / 2].x * s_mid, -(v[P / 2].y * s_mid)} : cx<T>{v[P / 2].x, -v[P / 2].y};
    }

    // ---- fused epilogue.  Every address is a per-thread base plus a compile-time offset.
    T* o_re = valid && want_cplx ? static_cast<T*>(p.out_re) + f * cbins : nullptr;
    T* o_im = valid && want_cplx ? static_cast<T*>(p.out_im) + f * cbins : nullptr;
    T* o_amp = valid && want_amp ? static_cast<T*>(p.amp) + f * bins : nullptr;
    T* o_ph = valid && PHASE && want_phase ? static_cast<T*>(p.phase) + f * bins : nullptr;
    T best_v = (T)0;  // scaled amplitude (squared key in peak-only mode) of the thread's best non-DC bin, 0 = none
    int best_k = 0;
    T dc_amp = (T)0;

    auto epilogue = [&](auto careful_c) {
      constexpr bool CAREFUL = decltype(careful_c)::value;  // second run with hypot() for out-of-range sums
      // findPeak keeps the first of equal values (strict '>' scanning upward): stream 0 is visited in
      // ascending k so '>' suffices; stream 1 is visited in descending k so '>=' lets the lower bin
      // win a tie (its threshold starts at the smallest positive number: a candidate needs v > 0).
      // Only (value, bin) are tracked per bin; the winner's re / im are picked out of XA / XB once per frame.
      T c0v = (T)0, c1v = sizeof(T) == 8 ? (T)4.9406564584124654e-324 : (T)1.401298464324817e-45;
      int c0k = 0, c1k = 0;
      unsigned hi_max = 0u, lo_min = 0xffffffffu;  // exponent range of re^2+im^2 seen by this thread (fp64)

      auto emit = [&](auto stream_c, auto off_c, auto edge_c, int k, cx<T> X, T ph) {
        constexpr int STREAM = decltype(stream_c)::value;
        constexpr int OFF = decltype(off_c)::value;    // element offset from the stream's base pointer
        constexpr bool EDGE = decltype(edge_c)::value;  // k may be 0 or M (DC / Nyquist)
        const int b0 = STREAM == 0 ? t : M - t;          // base bin of the stream
        const int m0 = STREAM == 0 ? N - t : M + t;      // base of the mirrored bin N - k
        if (want_cplx && o_re != nullptr) {
          (o_re + b0)[OFF] = X.x;
          (o_im + b0)[OFF] = X.y;
          if (cfull && (!EDGE || (k != 0 && k != M))) {
            (o_re + m0)[-OFF] = X.x;
            (o_im + m0)[-OFF] = -X.y;
          }
        }
        if (need_mag) {
          T a;
          if constexpr (KEYSQ && !CAREFUL) {
            // ranking key |X|^2 relative to the interior scale: Nyquist (and DC) carry (s_edge/s_mid)^2
            const T ss = X.x * X.x + X.y * X.y;
            if constexpr (sizeof(T) == 8) {
              const unsigned hi = (unsigned)__double2hiint((double)ss);
              hi_max = hi > hi_max ? hi : hi_max;
              lo_min = (hi - 1u) < lo_min ? (hi - 1u) : lo_min;
            }
            a = ss;
            if constexpr (EDGE) a = (k == 0 || k == M) ? ss * edge_key : ss;
          } else {
            T mag;
            if constexpr (CAREFUL) {
              mag = t_hypot_slow(X.x, X.y);
            } else {
              const T ss = X.x * X.x + X.y * X.y;
              if constexpr (sizeof(T) == 8) {
                const unsigned hi = (unsigned)__double2hiint((double)ss);
                hi_max = hi > hi_max ? hi : hi_max;
                lo_min = (hi - 1u) < lo_min ? (hi - 1u) : lo_min;
              }
              mag = fast_sqrt(ss);
            }
            if constexpr (FOLD) {
              a = mag;  // X carries s_mid already
              if constexpr (EDGE) a = (k == 0 || k == M) ? mag * edge_ratio : mag;
            } else {
              T scale = s_mid;
              if constexpr (EDGE) scale = (k == 0 || k == M) ? s_edge : s_mid;
              a = mag * scale;
            }
          }
          if (want_amp && o_amp != nullptr) {
            store_stream(o_amp + b0 + ((EDGE && k == M) ? -sh : sh) + OFF, a);  // shifted: k < M -> k + M, the Nyquist bin M -> 0
            if (two_sided && (!EDGE || (k != 0 && k != M))) store_stream(o_amp + m0 - sh - OFF, a);
          }
          if (want_peak) {
            bool is_dc = false;
            if constexpr (EDGE) is_dc = k == 0;
            if (is_dc) {
              dc_amp = a;
            } else if constexpr (STREAM == 0) {
              if (a > c0v) c0v = a, c0k = k;
            } else {
              if (a >= c1v) c1v = a, c1k = k;
            }
          }
        }
        if constexpr (PHASE) {
          if (want_phase && o_ph != nullptr) {
            (o_ph + b0 + ((EDGE && k == M) ? -sh : sh))[OFF] = ph;
            if (two_sided && (!EDGE || (k != 0 && k != M))) (o_ph + m0 - sh)[-OFF] = -ph;
          }
        }
      };
      using I0 = std::integral_constant<int, 0>;
      using I1 = std::integral_constant<int, 1>;

      if constexpr (M == 1) {
        Pair2<T> ph{(T)0, (T)0};
        if constexpr (PHASE) {
          if (want_phase) ph = t_atan2_x2(XA[0].y, XA[0].x, XB[0].y, XB[0].x);
        }
        emit(I0{}, I0{}, std::true_type{}, 0, XA[0], ph.a);
        emit(I1{}, I0{}, std::true_type{}, 1, XB[0], ph.b);
      } else {
        static_for<0, P / 2>([&](auto qi) {
          constexpr int q = decltype(qi)::value;
          const int k = t + TF * q;  // 0 <= k < M/2
          const cx<T> xa = XA[q], xb = XB[q];
          Pair2<T> ph{(T)0, (T)0};
          if constexpr (PHASE) {
            if (want_phase) {
              if constexpr (GEN || sizeof(T) == 4)
                ph = t_atan2_x2(xa.y, xa.x, xb.y, xb.x);
              else
                ph = atan2_x2_inline(xa.y, xa.x, xb.y, xb.x);
            }
          }
          emit(I0{}, std::integral_constant<int, TF * q>{}, std::bool_constant<q == 0>{}, k, xa, ph.a);
          emit(I1{}, std::integral_constant<int, -TF * q>{}, std::bool_constant<q == 0>{}, M - k, xb, ph.b);
        });
        if (t == 0) {
          T ph = (T)0;
          if constexpr (PHASE) {
            if (want_phase) ph = t_atan2(XH.y, XH.x);
          }
          emit(I0{}, std::integral_constant<int, M / 2>{}, std::false_type{}, M / 2, XH, ph);
        }
      }
      int verdict = 0;  // bit 0: a sum of squares left the safe range; bit 1: every sum of this thread was 0
      if constexpr (sizeof(T) == 8 && !CAREFUL) {
        // sums of squares outside [2^-1019, 2^+1022) rerun with hypot(); an exact 0 is fine (sqrt(0) = 0)
        // unless it is the underflow of a non-zero bin, which the caller settles for all-zero threads
        if (need_mag) verdict = ((hi_max >= 0x7fd00000u || lo_min < 0x003fffffu) ? 1 : 0) | (hi_max == 0u ? 2 : 0);
      }
      // merge the two streams: higher value wins, equal values go to the lower bin
      best_v = c0v, best_k = c0k;
      if (c1k != 0 && (best_k == 0 || peak_better(c1v, c1k, best_v, best_k))) best_v = c1v, best_k = c1k;
      return verdict;
    };

    {
      int verdict = epilogue(std::false_type{});
      if (!valid) verdict = 0;  // tail slots hold no frame
      if constexpr (sizeof(T) == 8) {
        // rare: some |X|^2 over/underflowed - redo the epilogue with hypot(), like Math.hypot (a warp-uniform decision)
        bool redo = simt::any((verdict & 1) != 0);
        if (!redo && simt::any((verdict & 2) != 0)) {
          // some thread saw only zeros: genuine silence, or bins below 2^-521 whose squares underflowed?
          bool nonzero = XH.x != (T)0 || XH.y != (T)0;
          static_for<0, NQ>([&](auto q) {
            nonzero = nonzero || XA[decltype(q)::value].x != (T)0 || XA[decltype(q)::value].y != (T)0 ||
                      XB[decltype(q)::value].x != (T)0 || XB[decltype(q)::value].y != (T)0;
          });
          redo = simt::any(nonzero);
        }
        if (redo) epilogue(std::true_type{});
      }
    }

    // ---- findPeak: (value desc, index asc) reduction over the frame's threads
    if (want_peak) {
      T bv = best_v;
      int bk = best_k;
      constexpr int W = TF < 32 ? TF : 32;
      PDSP_UNROLL
      for (int m = W / 2; m >= 1; m >>= 1) {
        const T ov = simt::shfl_xor(bv, m, W);
        const int ok = simt::shfl_xor(bk, m, W);
        if (peak_better(ov, ok, bv, bk)) {
          bv = ov;
          bk = ok;
        }
      }
      if constexpr (TF > 32) {
        // cross-warp stage through shared memory (reuses the exchange buffer)
        constexpr int NW = TF / 32;
        T* rv = reinterpret_cast<T*>(sm);
        int* rk = reinterpret_cast<int*>(rv + NW);
        if ((tl & 31) == 0) {
          rv[tl >> 5] = bv;
          rk[tl >> 5] = bk;
        }
        frame_sync<TF>(slot, SLOTS);
        bv = rv[0];
        bk = rk[0];
        for (int w = 1; w < NW; ++w) {
          const T ov = rv[w];
          const int ok = rk[w];
          if (peak_better(ov, ok, bv, bk)) {
            bv = ov;
            bk = ok;
          }
        }
        frame_sync<TF>(slot, SLOTS);
      }
      // The thread that owns the winning bin parks (index, re, im[, amplitude]) in the record; no non-DC
      // bin > 0 -> bin 0.  frequency / phase (and the amplitude in peak-only mode) are filled in by the
      // CTA-wide finishing loop below, where every lane has a record - an atan2 issued here would occupy
      // the whole warp for one lane.
      const bool owner = bk != 0 ? (best_k == bk) : (t == 0);
      if (valid && owner) {
        // which register holds bin bk: stream 0 (k = t + TF*q < M/2), stream 1 (k = M - t - TF*q > M/2), or XH (k = M/2)
        cx<T> Xw;
        if (M > 1 && 2 * bk == M)
          Xw = XH;
        else if (2 * bk < M || M == 1)
          Xw = pick_reg(XA, M == 1 ? 0 : (bk - t) / TF);
        else
          Xw = pick_reg(XB, (M - bk - t) / TF);
        if (M == 1 && bk == 1) Xw = XB[0];
        PeakRec<T> rec;
        rec.index = bk;
        rec.frequency = Xw.x;  // bk == 0: XA[0] of thread 0 = (X[0], +0)
        rec.phase = Xw.y;
        rec.amplitude = bk != 0 ? bv : dc_amp;  // a squared key in peak-only mode; recomputed below
        if constexpr (sizeof(T) == 8) rec.pad = 0;
        static_cast<PeakRec<T>*>(p.peaks)[f] = rec;
      }
    }
  }

  if (want_peak) {
    // finish the records of the frames this CTA processed: amplitude (peak-only mode), frequency, phase
    simt::sync_block();
    PeakRec<T>* recs = static_cast<PeakRec<T>*>(p.peaks);
    const long long f_first = g_begin * SLOTS;
    const long long f_last = g_end * SLOTS < p.batch ? g_end * SLOTS : p.batch;
    for (long long f = f_first + tid; f < f_last; f += THREADS) {
      PeakRec<T> rec = recs[f];
      const T re = rec.frequency, im = rec.phase;
      if (KEYSQ)
        rec.amplitude = t_mag_checked(re, im) * (FOLD ? ((rec.index == 0 || rec.index == M) ? edge_ratio : (T)1)
                                                      : ((rec.index == 0 || rec.index == M) ? s_edge : s_mid));
      rec.frequency = (T)((double)rec.index * p.bin_hz);
      rec.phase = t_atan2(im, re);
      recs[f] = rec;
      for (int g = 0; g < p.n_peers; ++g) static_cast<PeakRec<T>*>(p.peer[g])[p.peer_offset + f] = rec;
    }
  }
  ring_doorbell(p.door);
}

// Kernel entry points: the occupancy target is either __launch_bounds__(THREADS, MINB) or, for the
// tuning variants that want an exact register budget, __launch_bounds__(THREADS) __maxnreg__(MAXREG).
template <typename T, int LOG2M, int LOG2P, int MAXRB, int THREADS, int MINB, int MODE>
PDSP_GLOBAL void PDSP_LAUNCH_BOUNDS(THREADS, MINB) r2c_kernel(const R2CParams p) {
  r2c_body<T, LOG2M, LOG2P, MAXRB, THREADS, MODE>(p);
}
template <typename T, int LOG2M, int LOG2P, int MAXRB, int THREADS, int MAXREG, int MODE>
PDSP_GLOBAL void PDSP_KERNEL_LIMITS(THREADS, MAXREG) r2c_kernel_mr(const R2CParams p) {
  r2c_body<T, LOG2M, LOG2P, MAXRB, THREADS, MODE>(p);
}

template <typename T, int LOG2M, int LOG2P, int MAXRB, int THREADS, int MINB>
PDSP_GLOBAL void PDSP_LAUNCH_BOUNDS(THREADS, MINB) c2c_kernel(const C2CParams p) {
  using E = FftEngine<T, LOG2M, LOG2P, MAXRB>;
  constexpr int M = E::M, P = E::P, TF = E::TF;
  constexpr int SLOTS = THREADS / TF;
  static_assert(THREADS % TF == 0 && SLOTS >= 1, "CTA must hold whole frames");
  constexpr int SLOT_ELEMS = E::NEEDS_SMEM ? E::SMEM_ELEMS : 0;
  const int tid = simt::tid();
  const int slot = tid / TF;
  const int t = tid % TF;
  cx<T>* sm = reinterpret_cast<cx<T>*>(simt::smem()) + (size_t)slot * SLOT_ELEMS;
  const cx<T>* PDSP_RESTRICT tw = static_cast<const cx<T>*>(p.tw);
  const T scale = (T)(1.0 / (double)M);

  for (long long f0 = (long long)simt::bid() * SLOTS; f0 < p.batch; f0 += (long long)simt::nblocks() * SLOTS) {
    const long long f = f0 + slot;
    const bool valid = f < p.batch;
    cx<T> v[P];
    const T* ire = static_cast<const T*>(p.in_re) + (valid ? f * M : 0);
    const T* iim = p.in_im != nullptr ? static_cast<const T*>(p.in_im) + (valid ? f * M : 0) : nullptr;
    static_for<0, P>([&](auto qi) {
      constexpr int q = decltype(qi)::value;
      const int e = t + TF * q;
      const T re = valid ? ire[e] : (T)0;
      const T im = (valid && iim != nullptr) ? iim[e] : (T)0;
      // inverse = swap(FFT(swap(x))) / N
      v[q] = p.inverse ? cx<T>{im, re} : cx<T>{re, im};
    });
    E::fft(v, t, sm, tw, slot, SLOTS);
    if (valid) {
      T* ore = static_cast<T*>(p.out_re) + f * M;
      T* oim = static_cast<T*>(p.out_im) + f * M;
      static_for<0, P>([&](auto qi) {
        constexpr int q = decltype(qi)::value;
        const int e = t + TF * q;
        if (p.inverse) {
          ore[e] = v[q].y * scale;
          oim[e] = v[q].x * scale;
        } else {
          ore[e] = v[q].x;
          oim[e] = v[q].y;
        }
      });
    }
  }
  ring_doorbell(p.door);
}

}  // namespace pdsp
