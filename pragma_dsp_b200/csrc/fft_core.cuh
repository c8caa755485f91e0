// fft_core.cuh - register/shared-memory Stockham FFT engine (sm_100a), shared by all kernels.
//
// Replaces the reference's in-place radix-2 sweep `Radix2Fft.transform`
// (/root/reference/src/core/fft.ts:89-151) with an autosort (Stockham) mixed-radix FFT:
//   * one frame of M = 2^LOG2M complex points is held by TF = M/P threads, P = 2^LOG2P
//     points per thread in registers (thread t holds elements t + TF*q, q = 0..P-1, on entry
//     and on exit, natural order - no bit-reversal pass, no `rev[]` table);
//   * each pass runs P/R radix-R butterflies per thread (R = 2..32, constant twiddles folded
//     at compile time), multiplies by table twiddles, and exchanges through padded shared
//     memory (conflict-free for the radix-8/16 first pass);
//   * sign/normalisation follow the reference: forward e^{-2*pi*i*k*n/N}, unnormalised.
#pragma once
#include <type_traits>

#include "simt.h"

namespace pdsp {

template <typename T>
struct alignas(2 * sizeof(T)) cx {
  T x, y;
};

template <typename T>
PDSP_DEVICE cx<T> cadd(cx<T> a, cx<T> b) {
  return cx<T>{a.x + b.x, a.y + b.y};
}
template <typename T>
PDSP_DEVICE cx<T> csub(cx<T> a, cx<T> b) {
  return cx<T>{a.x - b.x, a.y - b.y};
}
template <typename T>
PDSP_DEVICE cx<T> cmul(cx<T> a, cx<T> w) {
  return cx<T>{a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x};
}

// elementwise a * s + b and a * s (per component, not complex products)
template <typename T>
PDSP_DEVICE cx<T> ew_fma(cx<T> a, cx<T> s, cx<T> b) {
  return cx<T>{a.x * s.x + b.x, a.y * s.y + b.y};
}
template <typename T>
PDSP_DEVICE cx<T> ew_mul(cx<T> a, cx<T> s) {
  return cx<T>{a.x * s.x, a.y * s.y};
}

#if defined(__CUDACC__) && !defined(PDSP_EMU)
// fp32 on sm_100a: a complex number is one 64-bit register pair, and Blackwell's packed fp32 pipe
// instructions (add/mul/fma.f32x2 -> SASS FADD2 / FMUL2 / FFMA2, with scalar-broadcast and half-swap
// operand modes) do both components at once: a complex add is 1 instruction instead of 2, a complex
// multiply 2 instead of 4.  The fp32 kernels are issue-bound (profiles/r1), so this is their main lever.
PDSP_DEVICE float2 as_f2(cx<float> a) { return make_float2(a.x, a.y); }
PDSP_DEVICE cx<float> as_cx(float2 a) { return cx<float>{a.x, a.y}; }
PDSP_DEVICE cx<float> cadd(cx<float> a, cx<float> b) { return as_cx(__fadd2_rn(as_f2(a), as_f2(b))); }
PDSP_DEVICE cx<float> csub(cx<float> a, cx<float> b) { return as_cx(__fadd2_rn(as_f2(a), make_float2(-b.x, -b.y))); }
PDSP_DEVICE cx<float> cmul(cx<float> a, cx<float> w) {
  // (a.x*w.x - a.y*w.y, a.x*w.y + a.y*w.x) = a.x * (w.x, w.y) + a.y * (-w.y, w.x)
  return as_cx(__ffma2_rn(make_float2(a.x, a.x), as_f2(w), __fmul2_rn(make_float2(a.y, a.y), make_float2(-w.y, w.x))));
}
PDSP_DEVICE cx<float> ew_fma(cx<float> a, cx<float> s, cx<float> b) { return as_cx(__ffma2_rn(as_f2(a), as_f2(s), as_f2(b))); }
PDSP_DEVICE cx<float> ew_mul(cx<float> a, cx<float> s) { return as_cx(__fmul2_rn(as_f2(a), as_f2(s))); }
#endif

// read-only (non-coherent) load of a table entry
#if defined(__CUDACC__) && !defined(PDSP_EMU)
#ifndef PDSP_TABLE_LD
#define PDSP_TABLE_LD 0  // experiment: 1 = table loads ask the L1 to evict their lines last (ld.global.nc.L1::evict_last)
#endif
PDSP_DEVICE cx<double> ldg_cx(const cx<double>* p) {
#if PDSP_TABLE_LD
  cx<double> v;
  asm volatile("ld.global.nc.L1::evict_last.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
#else
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return cx<double>{v.x, v.y};
#endif
}
PDSP_DEVICE cx<float> ldg_cx(const cx<float>* p) {
#if PDSP_TABLE_LD
  cx<float> v;
  asm volatile("ld.global.nc.L1::evict_last.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
#else
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return cx<float>{v.x, v.y};
#endif
}
// L2-coherent load (ld.global.cg): data another CTA of the same launch has just published
PDSP_DEVICE cx<double> ldcg_cx(const cx<double>* p) {
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return cx<double>{v.x, v.y};
}
PDSP_DEVICE cx<float> ldcg_cx(const cx<float>* p) {
  const float2 v = __ldcg(reinterpret_cast<const float2*>(p));
  return cx<float>{v.x, v.y};
}
#else
template <typename T>
PDSP_DEVICE cx<T> ldg_cx(const cx<T>* p) {
  return *p;
}
template <typename T>
PDSP_DEVICE cx<T> ldcg_cx(const cx<T>* p) {
  return *p;
}
#endif

// compile-time loop: f(std::integral_constant<int, I>) for I in [B, E)
template <int B, int E, class F>
PDSP_DEVICE void static_for(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(f);
  }
}

constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }
constexpr int bitrev(int k, int bits) {
  int r = 0;
  for (int b = 0; b < bits; ++b) r |= ((k >> b) & 1) << (bits - 1 - b);
  return r;
}

// Compile-time sine/cosine (constant evaluation only: Taylor series in long double after reduction to
// [-pi, pi]); used for per-register rotation constants such as the window recurrence in r2c_body.
constexpr long double kPiL = 3.141592653589793238462643383279502884L;
constexpr long double cx_reduce(long double x) {
  while (x > kPiL) x -= 2 * kPiL;
  while (x < -kPiL) x += 2 * kPiL;
  return x;
}
constexpr long double cx_sin(long double x) {
  x = cx_reduce(x);
  long double term = x, sum = x;
  for (int k = 1; k < 20; ++k) {
    term *= -x * x / ((2 * k) * (2 * k + 1));
    sum += term;
  }
  return sum;
}
constexpr long double cx_cos(long double x) {
  x = cx_reduce(x);
  long double term = 1, sum = 1;
  for (int k = 1; k < 20; ++k) {
    term *= -x * x / ((2 * k - 1) * (2 * k));
    sum += term;
  }
  return sum;
}

// cos(pi*k/16), k = 0..8 (round-to-nearest binary64 literals)
constexpr double cos16(int k) {
  switch (k) {
    case 0: return 1.0;
    case 1: return 0.98078528040323044912618223613424;
    case 2: return 0.92387953251128675612818318939679;
    case 3: return 0.83146961230254523707878837761791;
    case 4: return 0.70710678118654752440084436210485;
    case 5: return 0.55557023301960222474283081394853;
    case 6: return 0.38268343236508977172845998403040;
    case 7: return 0.19509032201612826784828486847702;
    default: return 0.0;
  }
}
// W32^K = exp(-2*pi*i*K/32) = (wr, wi), K in [0, 16)
constexpr double w32_re(int K) { return K <= 8 ? cos16(K) : -cos16(16 - K); }
constexpr double w32_im(int K) { return K <= 8 ? -cos16(8 - K) : -cos16(K - 8); }

#if defined(__CUDACC__) && !defined(PDSP_EMU)
// fp32: the 32nd roots of unity as operand pairs in the constant bank - (wr, wi) and (-wi, wr) - so that a packed complex
// multiply by a compile-time root is FMUL2 + FFMA2 with c[][] operands.  As immediates every use had to assemble both
// 64-bit pairs in registers first (MOV pairs: 45 of ~700 warp instructions per frame in the fp32 N = 1024 kernel).
#define PDSP_W32F(K) {{(float)w32_re(K), (float)w32_im(K)}, {-(float)w32_im(K), (float)w32_re(K)}}
static __constant__ float2 kW32f[16][2] = {PDSP_W32F(0),  PDSP_W32F(1),  PDSP_W32F(2),  PDSP_W32F(3), PDSP_W32F(4),  PDSP_W32F(5),
                                           PDSP_W32F(6),  PDSP_W32F(7),  PDSP_W32F(8),  PDSP_W32F(9), PDSP_W32F(10), PDSP_W32F(11),
                                           PDSP_W32F(12), PDSP_W32F(13), PDSP_W32F(14), PDSP_W32F(15)};
#undef PDSP_W32F
#endif
#ifndef PDSP_F32_CONST_TWIDDLES
#define PDSP_F32_CONST_TWIDDLES 1
#endif
// fp64: the same roots as c[][] operands of DMUL / DFMA instead of 64-bit literals (each literal costs two moves into a
// register pair wherever the compiler does not keep it live).  Experiment switch.
#ifndef PDSP_F64_CONST_TWIDDLES
#define PDSP_F64_CONST_TWIDDLES 0
#endif
#if defined(__CUDACC__) && !defined(PDSP_EMU) && PDSP_F64_CONST_TWIDDLES
#define PDSP_W32D(K) {w32_re(K), w32_im(K)}
static __constant__ double kW32d[16][2] = {PDSP_W32D(0),  PDSP_W32D(1),  PDSP_W32D(2),  PDSP_W32D(3), PDSP_W32D(4),  PDSP_W32D(5),
                                           PDSP_W32D(6),  PDSP_W32D(7),  PDSP_W32D(8),  PDSP_W32D(9), PDSP_W32D(10), PDSP_W32D(11),
                                           PDSP_W32D(12), PDSP_W32D(13), PDSP_W32D(14), PDSP_W32D(15)};
#undef PDSP_W32D
#endif

// d * W32^K with the trivial cases folded
template <int K, typename T>
PDSP_DEVICE cx<T> mul_w32(cx<T> d) {
  static_assert(K >= 0 && K < 16, "twiddle index");
  if constexpr (K == 0) {
    return d;
  } else if constexpr (K == 8) {  // -i
    return cx<T>{d.y, -d.x};
  } else if constexpr (sizeof(T) == 4) {
    // fp32: every non-trivial constant twiddle is one packed multiply + one packed fma
#if defined(__CUDACC__) && !defined(PDSP_EMU) && PDSP_F32_CONST_TWIDDLES
    return as_cx(__ffma2_rn(make_float2(d.x, d.x), kW32f[K][0], __fmul2_rn(make_float2(d.y, d.y), kW32f[K][1])));
#else
    constexpr T wr = (T)w32_re(K);
    constexpr T wi = (T)w32_im(K);
    return cmul(d, cx<T>{wr, wi});
#endif
#if defined(__CUDACC__) && !defined(PDSP_EMU) && PDSP_F64_CONST_TWIDDLES
  } else if constexpr (K == 4) {  // (1-i)/sqrt2
    const T c = kW32d[4][0];
    return cx<T>{(d.x + d.y) * c, (d.y - d.x) * c};
  } else if constexpr (K == 12) {  // (-1-i)/sqrt2
    const T c = kW32d[4][0];
    return cx<T>{(d.y - d.x) * c, -(d.x + d.y) * c};
  } else {
    const T wr = kW32d[K][0];
    const T wi = kW32d[K][1];
    return cx<T>{d.x * wr - d.y * wi, d.x * wi + d.y * wr};
  }
#else
  } else if constexpr (K == 4) {  // (1-i)/sqrt2
    constexpr T c = (T)cos16(4);
    return cx<T>{(d.x + d.y) * c, (d.y - d.x) * c};
  } else if constexpr (K == 12) {  // (-1-i)/sqrt2
    constexpr T c = (T)cos16(4);
    return cx<T>{(d.y - d.x) * c, -(d.x + d.y) * c};
  } else {
    constexpr T wr = (T)w32_re(K);
    constexpr T wi = (T)w32_im(K);
    return cx<T>{d.x * wr - d.y * wi, d.x * wi + d.y * wr};
  }
#endif
}

// d * W32^K for any K in [0, 32)
template <int K, typename T>
PDSP_DEVICE cx<T> mul_w32_full(cx<T> d) {
  if constexpr (K < 16) {
    return mul_w32<K>(d);
  } else {
    const cx<T> r = mul_w32<K - 16>(d);
    return cx<T>{-r.x, -r.y};
  }
}

// In-register radix-R DFT, decimation in frequency: natural order in, X[k] left in
// a[bitrev(k, log2 R)].  R in {1, 2, 4, 8, 16, 32}.
// FIRST: first stage to run (1: the caller has already done stage 0 - the r2c kernels fuse the window multiply into it).
template <typename T, int R, int FIRST = 0>
PDSP_DEVICE void dif_butterfly(cx<T> (&a)[R]) {
  constexpr int LR = ilog2(R);
  static_for<FIRST, LR>([&](auto st) {
    constexpr int h = R >> (decltype(st)::value + 1);
    static_for<0, R / 2>([&](auto bi) {
      constexpr int g = (decltype(bi)::value / h) * 2 * h;
      constexpr int i = decltype(bi)::value % h;
      constexpr int K = i * 16 / h;
      const cx<T> x = a[g + i];
      const cx<T> y = a[g + i + h];
      a[g + i] = cadd(x, y);
      a[g + i + h] = mul_w32<K>(csub(x, y));
    });
  });
}

// How the threads of one frame synchronise around a shared-memory exchange.
template <int TF>
PDSP_DEVICE void frame_sync(int slot, int slots_per_cta) {
  if constexpr (TF <= 32) {
    simt::sync_warp();  // a frame never spans warps
  } else {
    if (slots_per_cta == 1)
      simt::sync_block();
    else
      simt::sync_named(1 + slot, TF);
  }
}

#ifndef PDSP_SHUFFLE_EXCHANGE
#define PDSP_SHUFFLE_EXCHANGE 1
#endif
// fp64: derive the last pass' / the post-pass' twiddles from one table entry per thread (DP instructions) instead of
// loading one per butterfly (L1 wavefronts); experiment switches, see profiles/r2/README.md
#ifndef PDSP_DERIVE_LAST
#define PDSP_DERIVE_LAST 1
#endif
#ifndef PDSP_DERIVE_POST
#define PDSP_DERIVE_POST 1
#endif
// fp64, radix-16 middle passes: load W^r and W^{4r} only and build the other 13 twiddles W^{s*r} by complex products
// (52 DP instructions for 13 L1 loads; depth <= 3 products, ~1e-15 relative).  Experiment switch, default off.
#ifndef PDSP_DERIVE_MID
#define PDSP_DERIVE_MID 0
#endif

template <typename T, int LOG2M, int LOG2P, int MAXRB>
struct FftEngine {
  static_assert(LOG2P <= LOG2M, "points per thread cannot exceed the frame");
  static_assert(MAXRB >= 1 && MAXRB <= 5, "radix 2..32");
  static constexpr int M = 1 << LOG2M;
  static constexpr int P = 1 << LOG2P;
  static constexpr int TF = M / P;
  static constexpr int RB = LOG2P == 0 ? 1 : (MAXRB < LOG2P ? MAXRB : LOG2P);  // log2 radix of a full pass
  static constexpr int NPASS = LOG2M == 0 ? 0 : (LOG2M + RB - 1) / RB;
  static constexpr int pass_bits(int i) { return i < NPASS - 1 ? RB : LOG2M - RB * (NPASS - 1); }
  // one padding element per PAD_UNIT elements: stride of the first-pass scatter becomes odd
  static constexpr int ROW = 128 / (int)sizeof(cx<T>);
  static constexpr int PAD_UNIT = (1 << RB) > ROW ? (1 << RB) : ROW;
  static constexpr int PAD_SHIFT = ilog2(PAD_UNIT);
  static constexpr int SMEM_ELEMS = M + (M >> PAD_SHIFT) + 1;  // per frame slot
  static constexpr bool NEEDS_SMEM = NPASS > 1;
  // Per-schedule twiddle table: pass i >= 1 owns a block of (R_i - 1) * Ns_i entries laid out
  // [(s - 1) * Ns_i + r] = W_{Ns_i*R_i}^{s*r}, so that for a fixed leg s the lanes of a warp
  // (consecutive r) read consecutive entries - 4 wavefronts per warp load instead of up to 32 with
  // a natural-order exp(-2*pi*i*k/M) table (profiles/r1: that cost more L1 cycles than both exchanges).
  static constexpr int tw_offset(int pass) {
    int off = 0;
    for (int i = 1; i < pass; ++i) off += ((1 << pass_bits(i)) - 1) << (RB * i);
    return off;
  }
  static constexpr int TW_ELEMS = tw_offset(NPASS);
  PDSP_DEVICE static int pad(int i) { return i + (i >> PAD_SHIFT); }
  // Index of element t + TF*q in the padded buffer.  When TF is a multiple of the padding unit, TF*q contributes no carry
  // into (t + TF*q) >> PAD_SHIFT, so pad(t + TF*q) = pad(t) + q*(TF + TF/PAD_UNIT): one base register and an immediate
  // offset per access.  (ptxas does not see this: the fp64 N = 1024 kernel recomputed LOP3 + LEA.HI + IMAD in front of
  // each of its 16 LDS.128 - 45 of 1,285 warp instructions per frame.)
  static constexpr bool LINEAR_GATHER = (TF % PAD_UNIT) == 0;
  static constexpr int GATHER_STEP = TF + (TF >> PAD_SHIFT);
  PDSP_DEVICE static int gather_base(int t) { return pad(t); }
  template <int Q>
  PDSP_DEVICE static int gather_index(int t, int base) {
    if constexpr (LINEAR_GATHER)
      return base + Q * GATHER_STEP;
    else
      return pad(t + TF * Q);
  }

  // Exchange by shuffle instead of shared memory (see fft()): fp64, M = 512 = 16 x 16 x 2, one warp per frame, pass 1.
  template <bool BLOCKSYNC>
  static constexpr bool shuffle_pass(int pass) {
    return PDSP_SHUFFLE_EXCHANGE && !BLOCKSYNC && sizeof(T) == 8 && LOG2M == 9 && LOG2P == 4 && RB == 4 && NPASS == 3 &&
           pass == NPASS - 2;
  }
  // the last pass whose output goes through the shared-memory buffer (-1: none)
  template <bool BLOCKSYNC>
  static constexpr int last_smem_pass() {
    int r = -1;
    for (int i = 0; i + 1 < NPASS; ++i)
      if (!shuffle_pass<BLOCKSYNC>(i)) r = i;
    return r;
  }
  struct NoHook {
    PDSP_DEVICE void operator()() const {}
  };

  // v[q] holds element t + TF*q (natural order) on entry and the transform on exit.
  // tw: the per-schedule table described at tw_offset() (TW_ELEMS entries).
  // BLOCKSYNC: the frame's threads are spread over the CTA's warps (column-tiled passes of the
  // large-N path), so every exchange is a __syncthreads.
  // after_smem(): called once, as soon as the frame's threads are done with the shared-memory buffer (after the last
  // exchange that goes through it, or straight away when there is none) - the staged kernels refill it from there.
  // SPLITX: the exchange buffer holds ONE scalar plane (T elements, same padded indexing): the real parts go through it
  // first, then the imaginary parts - half the shared memory for two more barriers per exchange.  The large-FFT
  // pipeline kernel (bigfft3_kernels.cuh) uses it to keep a whole second tile landing while the current one is
  // transformed.  `sm` then points at the sequence's slot of T elements (passed as cx<T>* for one signature).
  // PRE0: the caller has run stage 0 of the first pass' butterflies itself (only for P == R, one butterfly per thread).
  static constexpr bool CAN_PRE0 = NPASS >= 1 && P == (1 << RB) && RB >= 1;
  template <bool BLOCKSYNC = false, class Hook = NoHook, bool SPLITX = false, bool PRE0 = false>
  PDSP_DEVICE static void fft(cx<T> (&v)[P], int t, cx<T>* sm, const cx<T>* PDSP_RESTRICT tw, int slot,
                              int slots_per_cta, Hook after_smem = Hook{}) {
    static_assert(!PRE0 || CAN_PRE0, "a pre-run first stage needs one full-radix butterfly per thread in pass 0");
    if constexpr (last_smem_pass<BLOCKSYNC>() < 0) after_smem();
    auto sync = [&]() {
      if constexpr (BLOCKSYNC)
        simt::sync_block();
      else
        frame_sync<TF>(slot, slots_per_cta);
    };
    static_for<0, NPASS>([&](auto pi) {
      constexpr int pass = decltype(pi)::value;
      constexpr int b = pass_bits(pass);
      constexpr int R = 1 << b;
      constexpr int BPT = P / R;  // butterflies per thread
      constexpr int NSL = RB * pass;  // log2 Ns
      constexpr int NS = 1 << NSL;
      constexpr bool last = pass == NPASS - 1;
      // Last pass with several butterflies per thread: Ns = TF*BPT, so r = t + TF*u and
      // W_M^{s*(t + TF*u)} = W_M^{s*t} * exp(-2*pi*i*s*u/P): one table entry per leg (u = 0) and a
      // compile-time 32nd root of unity for the others - P/R - 1 fewer L1 loads per leg.
      // (doubles only: the fp64 kernels are L1-bound with DP slack, the fp32 kernels are issue-bound with
      // L1 slack - there a load is cheaper than the four FP instructions of the derivation)
      constexpr bool DERIVE = PDSP_DERIVE_LAST && sizeof(T) == 8 && last && NSL > 0 && BPT > 1 && (32 % P) == 0;
      // Exchange by shuffle (fp64, M = 512 = 16 x 16 x 2, one warp per frame): the radix-2 last pass pairs element
      // j with j + 256, and after this radix-16 pass thread (h, l) = (t >> 4, t & 15) holds the 16 outputs
      // k = 0..15 of position 256h + l + 16k.  The natural layout t' + 32q the last pass reads wants, in thread
      // (h', l), the outputs of parity h' of both (0, l) and (1, l): each thread keeps 8 of its outputs and swaps
      // the other 8 with lane t ^ 16.  32 SHFLs replace 16 STS.128 + 16 LDS.128 (128 L1 wavefronts) and two
      // warp barriers; the kernel is L1-bound.
      constexpr bool SHUF = shuffle_pass<BLOCKSYNC>(pass);
      [[maybe_unused]] cx<T> w0[R];
      static_for<0, BPT>([&](auto ui) {
        constexpr int u = decltype(ui)::value;
        cx<T> a[R];
        static_for<0, R>([&](auto s) { a[decltype(s)::value] = v[u + decltype(s)::value * BPT]; });
        [[maybe_unused]] const int j = t + TF * u;
        if constexpr (NSL > 0) {
          if constexpr (DERIVE) {
            static_for<1, R>([&](auto si) {
              constexpr int s = decltype(si)::value;
              if constexpr (u == 0) w0[s] = ldg_cx(tw + tw_offset(pass) + (s - 1) * NS + t);
              const cx<T> w = mul_w32_full<(s * u * (32 / P)) % 32>(w0[s]);
              a[s] = cmul(a[s], w);
            });
          } else if constexpr (PDSP_DERIVE_MID && sizeof(T) == 8 && R == 16 && !last) {
            const cx<T>* PDSP_RESTRICT twp = tw + tw_offset(pass) + (j & (NS - 1));
            cx<T> w[16];
            w[1] = ldg_cx(twp);
            w[4] = ldg_cx(twp + 3 * NS);
            w[2] = cmul(w[1], w[1]);
            w[3] = cmul(w[2], w[1]);
            w[8] = cmul(w[4], w[4]);
            w[5] = cmul(w[4], w[1]);
            w[6] = cmul(w[4], w[2]);
            w[7] = cmul(w[4], w[3]);
            w[12] = cmul(w[8], w[4]);
            w[9] = cmul(w[8], w[1]);
            w[10] = cmul(w[8], w[2]);
            w[11] = cmul(w[8], w[3]);
            w[13] = cmul(w[12], w[1]);
            w[14] = cmul(w[12], w[2]);
            w[15] = cmul(w[12], w[3]);
            static_for<1, R>([&](auto s) { a[decltype(s)::value] = cmul(a[decltype(s)::value], w[decltype(s)::value]); });
          } else {
            const cx<T>* PDSP_RESTRICT twp = tw + tw_offset(pass) + (j & (NS - 1));
            static_for<1, R>([&](auto s) {
              const cx<T> w = ldg_cx(twp + (decltype(s)::value - 1) * NS);  // W_{Ns*R}^{s*r}
              a[decltype(s)::value] = cmul(a[decltype(s)::value], w);
            });
          }
        }
        dif_butterfly<T, R, (PRE0 && pass == 0) ? 1 : 0>(a);
        if constexpr (last) {
          static_for<0, R>([&](auto k) { v[u + decltype(k)::value * BPT] = a[bitrev(decltype(k)::value, b)]; });
        } else if constexpr (SHUF) {
          static_assert(BPT == 1 && R == 16 && P == 16 && TF == 32, "shuffle exchange layout");
          const bool h = ((t >> 4) & 1) != 0;
          static_for<0, 8>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            const cx<T> e = a[bitrev(2 * i, b)], o = a[bitrev(2 * i + 1, b)];
            const cx<T> send = h ? e : o;
            cx<T> got;
            got.x = simt::shfl_xor(send.x, 16, 32);
            got.y = simt::shfl_xor(send.y, 16, 32);
            v[i] = h ? got : e;
            v[8 + i] = h ? o : got;
          });
        } else if constexpr (SPLITX) {
          // outputs stay in the registers their inputs came from until both planes have been exchanged below
          static_for<0, R>([&](auto k) { v[u + decltype(k)::value * BPT] = a[bitrev(decltype(k)::value, b)]; });
        } else {
          const int base = ((j >> NSL) << (NSL + b)) + (j & (NS - 1));
          static_for<0, R>(
              [&](auto k) { sm[pad(base + (decltype(k)::value << NSL))] = a[bitrev(decltype(k)::value, b)]; });
        }
      });
      if constexpr (!last && !SHUF && SPLITX) {
        T* PDSP_RESTRICT smt = reinterpret_cast<T*>(sm);
        static_for<0, 2>([&](auto pl) {
          constexpr bool IM = decltype(pl)::value == 1;
          static_for<0, BPT>([&](auto ui) {
            constexpr int u = decltype(ui)::value;
            const int j = t + TF * u;
            const int base = ((j >> NSL) << (NSL + b)) + (j & (NS - 1));
            static_for<0, R>([&](auto k) {
              const cx<T> o = v[u + decltype(k)::value * BPT];
              smt[pad(base + (decltype(k)::value << NSL))] = IM ? o.y : o.x;
            });
          });
          sync();
          const int gb = gather_base(t);
          static_for<0, P>([&](auto q) {
            const T g = smt[gather_index<decltype(q)::value>(t, gb)];
            if constexpr (IM)
              v[decltype(q)::value].y = g;
            else
              v[decltype(q)::value].x = g;
          });
          sync();
        });
        if constexpr (pass == last_smem_pass<BLOCKSYNC>()) after_smem();
      } else if constexpr (!last && !SHUF) {
        sync();
        const int gb = gather_base(t);
        static_for<0, P>([&](auto q) { v[decltype(q)::value] = sm[gather_index<decltype(q)::value>(t, gb)]; });
        sync();
        if constexpr (pass == last_smem_pass<BLOCKSYNC>()) after_smem();
      }
    });
  }
};

}  // namespace pdsp
