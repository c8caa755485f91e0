// pragma_b200.cu - C ABI (include/pragma_b200.h): contexts, plans, staging pipeline, dispatch.
//
// Host-side responsibilities only: validate, build tables once per plan (K3: twiddles
// /root/reference/src/core/fft.ts:45-61 and windows src/xform/fourier.ts:14-52 are per-plan
// constants, computed in extended precision on the host and uploaded), move bytes between
// caller memory and HBM through pinned staging on two streams, and launch the kernels in
// fft_kernels.cuh.  There is no CPU implementation of the transform in this library.
#ifdef PDSP_EMU
#include "cuda_stub.h"
#else
#include <cuda_runtime.h>
#endif
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/pragma_b200.h"
#include "bigfft2_kernels.cuh"
#include "fft_launch.cuh"
#include "inst_groups.h"

#define PDSP_EXPORT extern "C" __attribute__((visibility("default")))

namespace pdsp {
// launchers defined in the inst.cu translation units
#define X(kind, tag, ctype, lo, hi) PDSP_DECL_##kind(tag, lo, hi)
#define PDSP_DECL_0(tag, lo, hi) \
  cudaError_t launch_r2c_##tag##_##lo##_##hi(int, int, const R2CParams&, const LaunchCtx&);
#define PDSP_DECL_1(tag, lo, hi) cudaError_t launch_c2c_##tag##_##lo##_##hi(int, const C2CParams&, const LaunchCtx&);
PDSP_GROUPS(X)
#undef X

// multi-pass large-N path (bigfft.cu)
int big_pass_c(int log2l);
void big_pass_tma_box(int log2l, int* cols, int* rows);
cudaError_t launch_big_pass(bool f64, int log2l, const BigPassParams& p, const LaunchCtx& lc);
cudaError_t launch_big_pass_tma(bool f64, int log2l, const BigPassParams& p, const simt::TensorMap2D& tm_re,
                                const simt::TensorMap2D& tm_im, const LaunchCtx& lc);
// second generation (bigfft2.cu): TMA tile loads and stores
int big2_pass_c(int log2l);
bool big_pipe_supported(bool f64, int log2l);
bool big_fused_supported(bool f64, int la, int lb);
cudaError_t launch_big_fused(int la, int lb, const BigTileParams& pa, const BigTileParams& pb, const BigFusedSync& fs,
                             const simt::TensorMap* ma, const simt::TensorMap* mb, const LaunchCtx& lc);
cudaError_t launch_big_pipe(bool f64, int log2l, bool last, const BigPassParams& p, const simt::TensorMap2D& tm_re,
                            const simt::TensorMap2D& tm_im, const LaunchCtx& lc);
cudaError_t launch_big_tile(bool f64, int log2l, int io, const BigTileParams& p, const simt::TensorMap* maps, const LaunchCtx& lc);

// tuning variants (inst_var.cu), one symbol per (type, variant): only in -DPDSP_TUNING builds
// (python -m pragma_dsp_b200.build --tuning); the shipped library carries the default mapping alone
#if defined(PDSP_TUNING) && !defined(PDSP_EMU)
#define PDSP_VARS(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) X(17)
#else
#define PDSP_VARS(X)
#endif
#define X(v)                                                                           \
  cudaError_t launch_r2c_var_f64_##v(int, const R2CParams&, const LaunchCtx&);         \
  cudaError_t launch_r2c_var_f32_##v(int, const R2CParams&, const LaunchCtx&);
PDSP_VARS(X)
#undef X

// Tunables of a context.  Read from the environment ONCE, in pdsp_ctx_create (PDSP_* variables), and changed
// afterwards only through pdsp_ctx_tune: nothing on the launch path calls getenv.
struct Tune {
  int variant = 0;          // PDSP_VARIANT: kernel mapping of N = 1024 (PDSP_TUNING builds only)
  int big_tma = -1;         // PDSP_BIG_TMA: -1 auto, 0 per-thread tile loads, 1 TMA tile loads (multi-pass path)
  int big_interleave = 1;   // PDSP_BIG_INTERLEAVE: interleaved (re, im) work buffer between passes
  int big_prefetch = 1;     // PDSP_BIG_PREFETCH: L2 prefetch of a CTA's next tile
  long long big_chunk = 0;  // PDSP_BIG_CHUNK: transforms per group of passes (0 = default)
  int big_factors[3] = {0, 0, 0};  // PDSP_BIG_FACTORS "a,b[,c]": log2 of forced pass lengths (test hook)
  int n_big_factors = 0;
  long long chunk_bytes = 0;  // PDSP_CHUNK_BYTES: staging chunk size of the host pipeline (0 = default)
  int staged = -1;            // PDSP_STAGED: 1 = bulk-staged sample loads where a staged kernel exists (default: direct loads)
  int big_resident = -1;      // PDSP_BIG_RESIDENT: three-pass transforms, passes 1+2 in L2-sized k1 groups: -1 / 0 off (measured slower), 1 on (automatic group size), n > 1 blocks per group
  int big_v2 = -1;            // PDSP_BIG_V2: large-FFT pass generation: -1 per pass (second where its box rows are >= 64 bytes), 0 first, 1 second
  int big_pipe = -1;          // PDSP_BIG_PIPE: third-generation (pipeline) large-FFT passes: -1 automatic (1024-point passes), 0 off, 1 every pass length it exists for (256 / 512 / 1024, fp64), 2 + m: the passes whose bit is set in m
  int big_fused = -1;         // PDSP_BIG_FUSED: three-pass transforms, middle + last pass in one persistent launch with tile-level hand-over through the L2: -1 automatic, 0 off, 1 on
  int fast = 1;               // PDSP_FAST: 0 disables the single-call fast lane (small host jobs then use the staging pipeline)
  int doorbell = 1;           // PDSP_DOORBELL: 0 = the fast lane waits with cudaStreamSynchronize instead of the in-kernel doorbell
  int copy_threads = 3;       // PDSP_COPY_THREADS: helper threads of an ingestion ring's host copies (0 = the caller alone)
};
static int tune_set(Tune& t, const char* key, const char* val) {
  if (!key) return 1;
  const char* v = val ? val : "";
  const bool unset = v[0] == 0;
  if (!strcmp(key, "variant")) {
    const int x = atoi(v);
    t.variant = (x < 0 || x >= kNumVariants) ? 0 : x;
  } else if (!strcmp(key, "big_tma")) {
    t.big_tma = unset ? -1 : (v[0] != '0');
  } else if (!strcmp(key, "big_interleave")) {
    t.big_interleave = unset ? 1 : (v[0] != '0');
  } else if (!strcmp(key, "big_prefetch")) {
    t.big_prefetch = unset ? 1 : (v[0] != '0');
  } else if (!strcmp(key, "big_chunk")) {
    t.big_chunk = atoll(v) > 0 ? atoll(v) : 0;
  } else if (!strcmp(key, "big_factors")) {
    int a = 0, b = 0, c3 = 0;
    const int got = sscanf(v, "%d,%d,%d", &a, &b, &c3);
    t.n_big_factors = got >= 2 ? got : 0;
    t.big_factors[0] = a, t.big_factors[1] = b, t.big_factors[2] = got == 3 ? c3 : 0;
  } else if (!strcmp(key, "chunk_bytes")) {
    t.chunk_bytes = atoll(v) > 0 ? atoll(v) : 0;
  } else if (!strcmp(key, "staged")) {
    t.staged = unset ? -1 : (v[0] != '0');
  } else if (!strcmp(key, "big_resident")) {
    t.big_resident = unset ? -1 : atoi(v);  // 0 off, 1 on (automatic group size), > 1: k1 blocks per group
  } else if (!strcmp(key, "big_v2")) {
    t.big_v2 = unset ? -1 : (v[0] != '0');
  } else if (!strcmp(key, "big_pipe")) {
    t.big_pipe = unset ? -1 : atoi(v);  // 0 off, 1 every supported length, 2 + mask: passes whose bit is set in mask (bit j = pass j)
  } else if (!strcmp(key, "big_fused")) {
    t.big_fused = unset ? -1 : atoi(v);  // 0 off, n >= 1: on, with the last-pass tiles n groups behind the middle tiles
  } else if (!strcmp(key, "fast")) {
    t.fast = unset ? 1 : (v[0] != '0');
  } else if (!strcmp(key, "doorbell")) {
    t.doorbell = unset ? 1 : (v[0] != '0');
  } else if (!strcmp(key, "copy_threads")) {
    t.copy_threads = unset ? 3 : (atoi(v) < 0 ? 0 : atoi(v));
  } else {
    return 1;
  }
  return 0;
}
static void tune_from_env(Tune& t) {
  static const char* const keys[][2] = {{"variant", "PDSP_VARIANT"},           {"big_tma", "PDSP_BIG_TMA"},
                                        {"big_interleave", "PDSP_BIG_INTERLEAVE"}, {"big_prefetch", "PDSP_BIG_PREFETCH"},
                                        {"big_chunk", "PDSP_BIG_CHUNK"},       {"big_factors", "PDSP_BIG_FACTORS"},
                                        {"chunk_bytes", "PDSP_CHUNK_BYTES"},   {"staged", "PDSP_STAGED"},
                                        {"big_resident", "PDSP_BIG_RESIDENT"}, {"fast", "PDSP_FAST"}, {"copy_threads", "PDSP_COPY_THREADS"}, {"big_v2", "PDSP_BIG_V2"}, {"doorbell", "PDSP_DOORBELL"}, {"big_pipe", "PDSP_BIG_PIPE"}, {"big_fused", "PDSP_BIG_FUSED"}};
  for (auto& k : keys)
    if (const char* e = getenv(k[1])) tune_set(t, k[0], e);
}

static cudaError_t dispatch_r2c(bool f64, int log2m, int mode, int variant, const R2CParams& p, const LaunchCtx& lc) {
  (void)variant;
  if (log2m == kVariantLog2M && (mode == MD_AMP || mode == (MD_AMP | MD_PEAK) || mode == MD_PEAK)) {
    switch (variant) {
#define X(v) \
  case v:    \
    return f64 ? launch_r2c_var_f64_##v(mode, p, lc) : launch_r2c_var_f32_##v(mode, p, lc);
      PDSP_VARS(X)
#undef X
      default:
        break;
    }
  }
#define X(kind, tag, ctype, lo, hi) PDSP_TRY_R_##kind(tag, ctype, lo, hi)
#define PDSP_TRY_R_0(tag, ctype, lo, hi)                                   \
  if (f64 == (sizeof(ctype) == 8) && log2m >= lo && log2m <= hi)           \
    return launch_r2c_##tag##_##lo##_##hi(log2m, mode, p, lc);
#define PDSP_TRY_R_1(tag, ctype, lo, hi)
  PDSP_GROUPS(X)
#undef X
  return cudaErrorInvalidValue;
}
static cudaError_t dispatch_c2c(bool f64, int log2m, const C2CParams& p, const LaunchCtx& lc) {
#define X(kind, tag, ctype, lo, hi) PDSP_TRY_C_##kind(tag, ctype, lo, hi)
#define PDSP_TRY_C_0(tag, ctype, lo, hi)
#define PDSP_TRY_C_1(tag, ctype, lo, hi)                                   \
  if (f64 == (sizeof(ctype) == 8) && log2m >= lo && log2m <= hi)           \
    return launch_c2c_##tag##_##lo##_##hi(log2m, p, lc);
  PDSP_GROUPS(X)
#undef X
  return cudaErrorInvalidValue;
}

// ---- small elementwise kernels: magnitude()/phase() on caller arrays, N = 1 frames ----------
PDSP_GLOBAL void k_magnitude(const double* PDSP_RESTRICT re, const double* PDSP_RESTRICT im, long long n,
                             double* PDSP_RESTRICT out) {
  const long long stride = (long long)simt::nblocks() * simt::nthreads();
  for (long long i = simt::bid() * (long long)simt::nthreads() + simt::tid(); i < n; i += stride)
    out[i] = hypot(re[i], im[i]);
}
PDSP_GLOBAL void k_phase(const double* PDSP_RESTRICT re, const double* PDSP_RESTRICT im, long long n,
                         double* PDSP_RESTRICT out) {
  const long long stride = (long long)simt::nblocks() * simt::nthreads();
  for (long long i = simt::bid() * (long long)simt::nthreads() + simt::tid(); i < n; i += stride)
    out[i] = fast_atan2(im[i], re[i]);
}
// applyWindow (src/xform/fourier.ts:54-67) and fftShift (:122-134) on caller arrays
PDSP_GLOBAL void k_apply_window(const double* PDSP_RESTRICT in, const double* PDSP_RESTRICT w, long long n,
                                double* PDSP_RESTRICT out) {
  const long long stride = (long long)simt::nblocks() * simt::nthreads();
  for (long long i = simt::bid() * (long long)simt::nthreads() + simt::tid(); i < n; i += stride) out[i] = in[i] * w[i];
}
PDSP_GLOBAL void k_fft_shift(const double* PDSP_RESTRICT in, long long n, double* PDSP_RESTRICT out) {
  const long long stride = (long long)simt::nblocks() * simt::nthreads();
  const long long mid = n / 2;
  for (long long i = simt::bid() * (long long)simt::nthreads() + simt::tid(); i < n; i += stride) {
    long long j = i + mid;
    if (j >= n) j -= n;
    out[i] = in[j];
  }
}
// out = a * (conj?) b * scale on planar arrays (src/math/complex.ts:87-105 mul / conj / scale, fused)
template <typename T>
PDSP_GLOBAL void k_complex_mul(const T* PDSP_RESTRICT are, const T* PDSP_RESTRICT aim, const T* PDSP_RESTRICT bre,
                               const T* PDSP_RESTRICT bim, int conj_b, T scale, long long n, T* PDSP_RESTRICT ore,
                               T* PDSP_RESTRICT oim) {
  const long long stride = (long long)simt::nblocks() * simt::nthreads();
  for (long long i = simt::bid() * (long long)simt::nthreads() + simt::tid(); i < n; i += stride) {
    const T ar = are[i], ai = aim[i], br = bre[i];
    const T bi = conj_b ? -bim[i] : bim[i];
    ore[i] = (ar * br - ai * bi) * scale;
    oim[i] = (ar * bi + ai * br) * scale;
  }
}
// N = 1: X[0] = x[0] * w[0] (createWindow(size 1) = [1]); one thread per frame
template <typename T>
PDSP_GLOBAL void k_r2c_n1(const R2CParams p) {
  const long long stride = (long long)simt::nblocks() * simt::nthreads();
  for (long long f = simt::bid() * (long long)simt::nthreads() + simt::tid(); f < p.batch; f += stride) {
    T x = (T)0;
    if (p.frame_len >= 1)
      x = p.sample_dtype == DT_F32 ? (T) static_cast<const float*>(p.samples)[f * p.hop]
                                   : (T) static_cast<const double*>(p.samples)[f * p.hop];
    if (p.out_re) {
      static_cast<T*>(p.out_re)[f] = x;
      static_cast<T*>(p.out_im)[f] = (T)0;
    }
    const T a = (T)fabs((double)x) * (T)p.scale_edge;
    const T ph = (T)atan2(0.0, (double)x);
    if (p.amp) static_cast<T*>(p.amp)[f] = a;
    if (p.phase) static_cast<T*>(p.phase)[f] = ph;
    if (p.peaks) {
      PeakRec<T> r;
      memset(&r, 0, sizeof(r));
      r.index = 0;
      r.frequency = (T)0;
      r.amplitude = a;
      r.phase = ph;
      static_cast<PeakRec<T>*>(p.peaks)[f] = r;
    }
  }
}

// ---- spectrum() for FFT sizes beyond one CTA (N > 16384): buildFrame + applyWindow into a dense plane, the
// multi-pass transform, then this epilogue over the full complex spectrum (src/public/spectrum.ts:107-142).
template <typename T, typename S>
PDSP_GLOBAL void k_build_frames(const S* PDSP_RESTRICT s, long long hop, int frame_len, const T* PDSP_RESTRICT win, int n,
                                long long batch, T* PDSP_RESTRICT out) {
  const long long stride = (long long)simt::nblocks() * simt::nthreads();
  const long long total = batch * (long long)n;
  for (long long i = simt::bid() * (long long)simt::nthreads() + simt::tid(); i < total; i += stride) {
    const long long f = i / n;
    const int e = (int)(i - f * n);
    T x = e < frame_len ? (T)s[f * hop + e] : (T)0;  // buildFrame: truncate / zero-pad (spectrum.ts:36-43)
    if (win != nullptr) x *= win[e];                 // applyWindow (fourier.ts:54-67)
    out[i] = x;
  }
}

// One CTA per frame: amplitude scaling (spectrum.ts:45-72), phase, findPeak (:74-105) with its first-of-equals rule.
template <typename T>
PDSP_GLOBAL void k_big_epilogue(const T* PDSP_RESTRICT re, const T* PDSP_RESTRICT im, int n, int bins, T s_edge, T s_mid,
                                int one_sided, int shift, double bin_hz, long long batch, T* PDSP_RESTRICT amp,
                                T* PDSP_RESTRICT phase, PeakRec<T>* PDSP_RESTRICT peaks) {
  T* sv = reinterpret_cast<T*>(simt::smem());                         // per-warp best value
  int* sk = reinterpret_cast<int*>(simt::smem() + 32 * sizeof(T));    // per-warp best bin
  const int tid = simt::tid(), nth = simt::nthreads();
  const int half = n / 2;
  for (long long f = simt::bid(); f < batch; f += simt::nblocks()) {
    const T* fre = re + f * (long long)n;
    const T* fim = im + f * (long long)n;
    T bv = (T)0;
    int bk = 0;
    for (int k = tid; k < bins; k += nth) {
      const T xr = fre[k];
      // DC and Nyquist of a real frame are real; the reference's imaginary parts there are sums of +0
      const T xi = (k == 0 || k == half) ? (T)0 : fim[k];
      const T scale = (one_sided && (k == 0 || k == half)) ? s_edge : s_mid;
      const T a = (T)hypot((double)xr, (double)xi) * scale;
      const int ks = shift ? ((k + half) & (n - 1)) : k;  // fftShift fused into the store (two-sided rows)
      if (amp != nullptr) amp[f * (long long)bins + ks] = a;
      if (phase != nullptr) phase[f * (long long)bins + ks] = fast_atan2(xi, xr);
      if (k >= 1 && a > bv) {  // k ascends within a thread: strict '>' keeps the first of equal values
        bv = a;
        bk = k;
      }
    }
    if (peaks != nullptr) {
      for (int m = 16; m >= 1; m >>= 1) {
        const T ov = simt::shfl_xor(bv, m, 32);
        const int ok = simt::shfl_xor(bk, m, 32);
        if (ok != 0 && (bk == 0 || peak_better(ov, ok, bv, bk))) {
          bv = ov;
          bk = ok;
        }
      }
      if ((tid & 31) == 0) {
        sv[tid >> 5] = bv;
        sk[tid >> 5] = bk;
      }
      simt::sync_block();
      if (tid == 0) {
        for (int w = 1; w < (nth + 31) / 32; ++w) {
          if (sk[w] != 0 && (bk == 0 || peak_better(sv[w], sk[w], bv, bk))) {
            bv = sv[w];
            bk = sk[w];
          }
        }
        // no non-DC bin above zero: findPeak falls back to the DC bin
        const T xr = fre[bk];
        const T xi = (bk == 0 || bk == half) ? (T)0 : fim[bk];
        const T scale = (one_sided && (bk == 0 || bk == half)) ? s_edge : s_mid;
        PeakRec<T> r;
        memset(&r, 0, sizeof(r));
        r.index = bk;
        r.frequency = (T)((double)bk * bin_hz);
        r.amplitude = (T)hypot((double)xr, (double)xi) * scale;
        r.phase = fast_atan2(xi, xr);
        peaks[f] = r;
      }
      simt::sync_block();
    }
  }
}
}  // namespace pdsp

using namespace pdsp;

// ------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = std::string("pragma-dsp/b200: ") + buf;
  return 1;
}
#define CU(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) return fail("%s: %s", #call, cudaGetErrorString(e_));         \
  } while (0)

// ------------------------------------------------------------------------------ objects
static const int kMaxBigLog2N = 28;  // largest multi-pass transform (2^28 complex points)
static const int kSlots = 3;  // staging pipeline depth (chunks in flight)

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  void* d_in = nullptr;
  size_t d_in_cap = 0;
  void* d_out = nullptr;
  size_t d_out_cap = 0;
  void* h_in = nullptr;
  size_t h_in_cap = 0;
  void* h_out = nullptr;
  size_t h_out_cap = 0;
  // pending copy-out of a pageable destination: (dst, src in h_out, bytes)
  struct Flush {
    void* dst;
    const void* src;
    size_t bytes;
  };
  std::vector<Flush> flush;
  bool busy = false;
};

// Single-call fast lane (Radix2Fft.forward / spectrum() on one or a few frames: src/core/fft.ts:77-79,
// src/public/spectrum.ts:107-142 are synchronous one-frame calls).  The 3-slot pipeline costs four driver calls,
// two copy-engine hops and an event wait per call - more than the transform.  Small jobs instead go through ONE
// kernel launch: samples are memcpy'd into a host-mapped pinned buffer that the kernel reads and writes directly
// over PCIe, and completion is a doorbell word the last CTA stores into that buffer (the host spins on it).
struct FastLane {
  unsigned char* h = nullptr;  // pinned + mapped: [in | outputs], kFastBytes, then the doorbell word
  unsigned char* d = nullptr;  // device alias of h
  unsigned* d_count = nullptr;  // CTA counter of the doorbell (device memory, zero between launches)
  unsigned seq = 0;
  cudaStream_t stream = nullptr;
};
static const size_t kFastBytes = 256u << 10;

struct pdsp_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  Slot slots[kSlots];
  std::mutex mu;       // serialises the host-entry staging pipeline (slots)
  std::recursive_mutex plan_mu;  // guards the plan cache and lazily built tables (taken inside `mu`)
  std::map<std::pair<int, int>, pdsp_plan*> plans;
  std::map<std::tuple<int, int, int>, void*> pass_tw;  // (f64, log2m, rb) -> per-pass twiddle table
  std::atomic<long long> launches{0};
  Tune tune;  // read from the environment at creation, changed by pdsp_ctx_tune
  FastLane fast;
  std::atomic<long long> fast_calls{0};
};

// multi-pass plan of a transform too long for one CTA: N = 2^lg[0] * 2^lg[1] (* 2^lg[2])
struct BigPlan {
  int npass = 0;
  int lg[3] = {0, 0, 0};
  void* tw_hi[2] = {nullptr, nullptr};               // two-level inter-pass twiddles of passes 0 and 1
  void* tw_lo[2] = {nullptr, nullptr};
  int log_b[2] = {0, 0};
  // Intermediate planes between the passes, ONE SET PER STREAM: the host pipeline runs neighbouring chunks
  // on different streams, and passes of two transforms sharing a work buffer would overwrite each other.
  struct Work {
    void* re = nullptr;  // frames * N elements (2x when interleaved: cx<T> elements, `im` unused)
    void* im = nullptr;
    long long frames = 0;
    bool interleaved = false;
    unsigned* sync = nullptr;  // fused middle + last pass: 1024 group counters + 1 error word (device memory)
    // large-N spectrum(): windowed frames and their full complex spectra (spec_frames transforms each)
    void* spec_x = nullptr;
    void* spec_re = nullptr;
    void* spec_im = nullptr;
    long long spec_frames = 0;
  };
  std::map<cudaStream_t, Work> work;
};

struct pdsp_plan {
  pdsp_ctx* ctx;
  int n;
  int log2n;
  int precision;
  void* d_post = nullptr;  // cx<T>[n/4 + 1]
  void* d_win[4] = {nullptr, nullptr, nullptr, nullptr};
  BigPlan* big = nullptr;      // built lazily for n > 8192
};

static int set_device(const pdsp_ctx* c) {
  CU(cudaSetDevice(c->device));
  return 0;
}

static int ensure(void** p, size_t* cap, size_t need, bool host) {
  if (*cap >= need) return 0;
  if (*p) {
    if (host)
      CU(cudaFreeHost(*p));
    else
      CU(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
  }
  size_t sz = need + need / 4 + 256;
  if (host)
    CU(cudaHostAlloc(p, sz, cudaHostAllocDefault));
  else
    CU(cudaMalloc(p, sz));
  *cap = sz;
  return 0;
}

static bool is_device_visible(const void* p) {
  // true when `p` can be the source/target of an async DMA without staging (pinned or registered host memory)
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// exp(-2*pi*i*k/n) with exact values on the axes and octant symmetry, evaluated in long double
static void twiddle(int k, int n, long double* re, long double* im) {
  // reduce k/n to the first octant
  const long double two_pi = 6.283185307179586476925286766559005768L;
  int q = (int)(((long long)k * 8) / n);  // octant 0..7
  long long kk = k;
  auto cs = [&](long long j, long double* c, long double* s) {  // angle 2*pi*j/n, 0 <= j <= n/8
    if (j == 0) {
      *c = 1.0L;
      *s = 0.0L;
    } else if (8 * j == n) {
      *c = *s = 0.707106781186547524400844362104849039L;
    } else {
      long double a = two_pi * (long double)j / (long double)n;
      *c = cosl(a);
      *s = sinl(a);
    }
  };
  long double c, s;  // cos/sin of the positive angle theta = 2*pi*k/n
  switch (q) {
    case 0: cs(kk, &c, &s); break;
    case 1: { long double a, b; cs(n / 4 - kk, &a, &b); c = b; s = a; } break;
    case 2: { long double a, b; cs(kk - n / 4, &a, &b); c = -b; s = a; } break;
    case 3: { long double a, b; cs(n / 2 - kk, &a, &b); c = -a; s = b; } break;
    case 4: { long double a, b; cs(kk - n / 2, &a, &b); c = -a; s = -b; } break;
    case 5: { long double a, b; cs(3 * (long long)n / 4 - kk, &a, &b); c = -b; s = -a; } break;
    case 6: { long double a, b; cs(kk - 3 * (long long)n / 4, &a, &b); c = b; s = -a; } break;
    default: { long double a, b; cs(n - kk, &a, &b); c = a; s = -b; } break;
  }
  *re = c;
  *im = -s;
}

// Per-pass twiddles of an M = 2^log2m point Stockham schedule whose full passes have radix 2^rb
// (same schedule arithmetic as FftEngine: all passes rb bits, the last one the remainder).  Block of
// pass i >= 1: [(s - 1) * Ns + r] = exp(-2*pi*i * s*r / (Ns * R)), s = 1..R-1, r = 0..Ns-1.
template <typename T>
static int upload_pass_twiddles(int log2m, int rb, void** d_out) {
  std::vector<cx<T>> tab;
  const int npass = log2m == 0 ? 0 : (log2m + rb - 1) / rb;
  for (int i = 1; i < npass; ++i) {
    const int bits = i < npass - 1 ? rb : log2m - rb * (npass - 1);
    const int R = 1 << bits, Ns = 1 << (rb * i), n = Ns * R;
    for (int sidx = 1; sidx < R; ++sidx)
      for (int r = 0; r < Ns; ++r) {
        long double re, im;
        const int k = (int)(((long long)sidx * r) % n);
        if (n >= 8) {
          twiddle(k, n, &re, &im);
        } else {  // n = 4
          static const long double c4[4] = {1, 0, -1, 0}, s4[4] = {0, -1, 0, 1};
          re = c4[k * (4 / n)];
          im = s4[k * (4 / n)];
        }
        tab.push_back(cx<T>{(T)re, (T)im});
      }
  }
  if (tab.empty()) tab.push_back(cx<T>{(T)1, (T)0});
  CU(cudaMalloc(d_out, sizeof(cx<T>) * tab.size()));
  CU(cudaMemcpy(*d_out, tab.data(), sizeof(cx<T>) * tab.size(), cudaMemcpyHostToDevice));
  return 0;
}

static const void* pass_twiddles_cb(void* owner, bool f64, int log2m, int rb) {
  pdsp_ctx* c = static_cast<pdsp_ctx*>(owner);
  std::lock_guard<std::recursive_mutex> lk(c->plan_mu);
  auto key = std::make_tuple((int)f64, log2m, rb);
  auto it = c->pass_tw.find(key);
  if (it != c->pass_tw.end()) return it->second;
  void* d = nullptr;
  const int rc = f64 ? upload_pass_twiddles<double>(log2m, rb, &d) : upload_pass_twiddles<float>(log2m, rb, &d);
  if (rc) return nullptr;
  c->pass_tw[key] = d;
  return d;
}

// Hermitian post-pass table of a real plan: post[k] = W_N^k * (-i/2) = (wi/2, -wr/2), k = 0..N/4
template <typename T>
static int upload_tables(pdsp_plan* pl) {
  const int n = pl->n;
  const int m = n / 2;
  const int np = m / 2 + 1;
  std::vector<cx<T>> post((size_t)np);
  for (int k = 0; k < np; ++k) {
    long double re = 1, im = 0;
    if (n >= 8) {
      twiddle(k, n, &re, &im);
    } else if (n == 4) {
      static const long double c4[4] = {1, 0, -1, 0}, s4[4] = {0, -1, 0, 1};
      re = c4[k];
      im = s4[k];
    }
    post[k] = cx<T>{(T)(im / 2), (T)(-re / 2)};
  }
  CU(cudaMalloc(&pl->d_post, sizeof(cx<T>) * (size_t)np));
  CU(cudaMemcpy(pl->d_post, post.data(), sizeof(cx<T>) * (size_t)np, cudaMemcpyHostToDevice));
  return 0;
}

static int window_host(int window, int size, double* out) {
  // src/xform/fourier.ts:14-52, same formulas and operation order, binary64
  if (size <= 0) return fail("Window size must be positive, got %d", size);
  if (window < 0 || window > 3) return fail("Unsupported window type: %d", window);
  if (size == 1) {
    out[0] = 1.0;
    return 0;
  }
  const double pi = 3.141592653589793;
  for (int i = 0; i < size; ++i) {
    const double f = (2 * pi * i) / (size - 1);
    switch (window) {
      case PDSP_WIN_RECT: out[i] = 1.0; break;
      case PDSP_WIN_HANN: out[i] = 0.5 * (1 - cos(f)); break;
      case PDSP_WIN_HAMMING: out[i] = 0.54 - 0.46 * cos(f); break;
      default: out[i] = 0.42 - 0.5 * cos(f) + 0.08 * cos(2 * f); break;
    }
  }
  return 0;
}

static int plan_window(pdsp_plan* pl, int window, const void** d_win) {
  if (window < 0 || window > 3) return fail("Unsupported window type: %d", window);
  if (window == PDSP_WIN_RECT) {  // multiply by 1 is skipped in the kernel
    *d_win = nullptr;
    return 0;
  }
  std::lock_guard<std::recursive_mutex> lk(pl->ctx->plan_mu);
  if (!pl->d_win[window]) {
    std::vector<double> w((size_t)pl->n);
    if (window_host(window, pl->n, w.data())) return 1;
    void* d = nullptr;
    if (pl->precision == PDSP_F64) {
      CU(cudaMalloc(&d, sizeof(double) * w.size()));
      CU(cudaMemcpy(d, w.data(), sizeof(double) * w.size(), cudaMemcpyHostToDevice));
    } else {
      std::vector<float> wf(w.begin(), w.end());
      CU(cudaMalloc(&d, sizeof(float) * wf.size()));
      CU(cudaMemcpy(d, wf.data(), sizeof(float) * wf.size(), cudaMemcpyHostToDevice));
    }
    pl->d_win[window] = d;
  }
  *d_win = pl->d_win[window];
  return 0;
}

// ------------------------------------------------------------------------------ kernel launches
static size_t esize(int dtype) { return dtype == PDSP_F64 ? 8 : 4; }

struct SpecGeom {
  int n, bins;
  size_t in_elems_per_frame_stride;  // hop
  size_t out_esize, peak_size;
};

struct PeerSpec {
  void* const* ptrs;
  int n;
  long long offset;
};

struct BigPlan;
static int big_plan(pdsp_plan* pl, BigPlan** out);
static int launch_c2c(pdsp_plan* pl, const void* d_re, const void* d_im, long long batch, void* d_ore, void* d_oim,
                      int inverse, cudaStream_t st, const Doorbell* door = nullptr, bool* door_used = nullptr);
static int launch_spectrum_big(pdsp_plan* pl, const R2CParams& p, long long batch, cudaStream_t st);

static int launch_spectrum(pdsp_plan* pl, const pdsp_spectrum_desc* d, const void* d_samples, long long batch,
                           void* d_amp, void* d_phase, void* d_peaks, void* d_cre, void* d_cim, int cfull,
                           cudaStream_t st, const PeerSpec* peers = nullptr, const Doorbell* door = nullptr,
                           bool* door_used = nullptr) {
  pdsp_ctx* c = pl->ctx;
  const int n = pl->n;
  R2CParams p;
  memset(&p, 0, sizeof p);
  p.samples = d_samples;
  p.sample_dtype = d->sample_dtype == PDSP_F64 ? DT_F64 : DT_F32;
  const size_t es = esize(d->sample_dtype);
  p.vec_ok = ((reinterpret_cast<uintptr_t>(d_samples) % (2 * es)) == 0 && (d->hop % 2) == 0) ? 1 : 0;
  p.frame_len = d->frame_len;
  p.hop = d->hop;
  p.batch = batch;
  const void* win = nullptr;
  if (plan_window(pl, d->window, &win)) return 1;
  p.window = win;
  p.post = pl->d_post;
  p.out_re = d_cre;
  p.out_im = d_cim;
  p.cfull = cfull;
  p.amp = d_amp;
  p.phase = d_phase;
  p.peaks = d_peaks;
  if (peers) {
    for (int g = 0; g < peers->n; ++g) p.peer[g] = peers->ptrs[g];
    p.n_peers = peers->n;
    p.peer_offset = peers->offset;
  }
  p.two_sided = d->sides == PDSP_SIDES_TWO;
  p.shift = (p.two_sided && d->fft_shift) ? 1 : 0;
  if (d->raw_magnitude) {
    p.scale_edge = p.scale_mid = 1.0;
  } else if (p.two_sided) {
    p.scale_edge = p.scale_mid = 1.0 / (double)n;  // mag / size (spectrum.ts:63-72), exact for 2^k
  } else {
    p.scale_edge = 1.0 / (double)n;  // mag / size for DC and Nyquist (spectrum.ts:52-55)
    p.scale_mid = 2.0 / (double)n;   // (2 * mag) / size, identical bits for power-of-two size
  }
  p.bin_hz = d->sample_rate / (double)n;  // binFrequencies: quotient first (fourier.ts:160)
  if (door_used) *door_used = false;
  {
    // beyond one CTA (or forced by the big_factors test hook): window -> multi-pass transform -> epilogue
    bool big = pl->log2n - 1 > kMaxLog2M;
    if (!big && n > 1 && c->tune.n_big_factors) {
      BigPlan* bp = nullptr;
      if (big_plan(pl, &bp)) return 1;
      big = bp != nullptr;
    }
    if (big) {
      if (d_cre) return fail("internal: complex output of a large real transform goes through launch_c2c");
      if (peers && peers->n > 0) return fail("pdsp_spectrum_dev_gather is limited to FFT sizes up to %d", 2 << kMaxLog2M);
      return launch_spectrum_big(pl, p, batch, st);
    }
  }
  LaunchCtx lc{c->device, c->sm_count, st, pass_twiddles_cb, c};
  cudaError_t e;
  if (n == 1) {
    const int threads = 128;
    long long blocks = (batch + threads - 1) / threads;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    if (pl->precision == PDSP_F64)
      PDSP_LAUNCH(k_r2c_n1<double>, (int)blocks, threads, 0, st, p);
    else
      PDSP_LAUNCH(k_r2c_n1<float>, (int)blocks, threads, 0, st, p);
    e = cudaGetLastError();
  } else {
    // compile-time specialised kernel when the call is the regular batched shape, generic otherwise
    int mode = (d_amp ? MD_AMP : 0) | (d_phase ? MD_PHASE : 0) | (d_peaks ? MD_PEAK : 0) | (d_cre ? MD_CPLX : 0) |
               (p.two_sided ? MD_TWO : 0) | (d->frame_len < n ? MD_PAD : 0);
    if ((mode & MD_PAD) && (d->frame_len & 1)) mode = -1;  // MD_PAD loads whole sample pairs: even frame lengths only
    const bool regular = p.vec_ok && (d_cre == nullptr || (cfull && !p.two_sided)) &&
                         mode_is_specialised(mode);
    if (regular && c->tune.staged > 0 && !(mode & (MD_PAD | MD_TWO)) && d->frame_len >= n) {
      // bulk-staged sample loads (MD_STAGED), opt-in: measured 3-9 % SLOWER than the direct loads on B200
      // (profiles/r2/README.md), kept as a switchable experiment for N = 1024.  16-byte aligned frames, sample type
      // no wider than the plan's; dispatch falls back to the direct-load kernel for sizes that have no staged form
      const bool aligned = (reinterpret_cast<uintptr_t>(d_samples) % 16) == 0 && ((size_t)d->hop * es) % 16 == 0 && ((size_t)n * es) % 16 == 0;
      if (aligned && es <= esize(pl->precision)) mode |= MD_STAGED;
    }
    if (door) {
      p.door = *door;
      if (door_used) *door_used = true;
    }
    e = dispatch_r2c(pl->precision == PDSP_F64, pl->log2n - 1, regular ? mode : MD_GENERIC, c->tune.variant, p, lc);
  }
  if (e != cudaSuccess) return fail("r2c launch (n=%d): %s", n, cudaGetErrorString(e));
  c->launches++;
  return 0;
}

// ------------------------------------------------------------------------------ large-N (multi-pass) path
template <typename T>
static int upload_big_twiddles(int log_nt, int log_b, void** d_hi, void** d_lo) {
  const long long nt = 1LL << log_nt, nb = 1LL << log_b, nh = nt >> log_b;
  std::vector<cx<T>> hi((size_t)nh), lo((size_t)nb);
  for (long long a = 0; a < nh; ++a) {
    long double re, im;
    twiddle((int)(a << log_b), (int)nt, &re, &im);
    hi[(size_t)a] = cx<T>{(T)re, (T)im};
  }
  for (long long b = 0; b < nb; ++b) {
    long double re, im;
    twiddle((int)b, (int)nt, &re, &im);
    lo[(size_t)b] = cx<T>{(T)re, (T)im};
  }
  CU(cudaMalloc(d_hi, sizeof(cx<T>) * hi.size()));
  CU(cudaMalloc(d_lo, sizeof(cx<T>) * lo.size()));
  CU(cudaMemcpy(*d_hi, hi.data(), sizeof(cx<T>) * hi.size(), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(*d_lo, lo.data(), sizeof(cx<T>) * lo.size(), cudaMemcpyHostToDevice));
  return 0;
}

// 2-D tensor map over one plane viewed as [rows][cols] (cols contiguous) with a {box_cols, box_rows} box.
static int make_tensor_map(simt::TensorMap2D* tm, const void* base, bool f64, long long cols, long long rows,
                           int box_cols, int box_rows) {
#ifdef PDSP_EMU
  tm->base = base;
  tm->dim0 = cols;
  tm->dim1 = rows;
  tm->esize = f64 ? 8 : 4;
  tm->stride1_bytes = cols * tm->esize;
  tm->box0 = box_cols;
  tm->box1 = box_rows;
  return 0;
#else
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled is not available from this driver");
    encode = reinterpret_cast<encode_fn>(fn);
  }
  const cuuint64_t es = f64 ? 8 : 4;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * es};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(tm, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                            const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  return 0;
#endif
}

// Builds (once) the pass structure of a large transform.  *out stays null when PDSP_BIG_FACTORS is
// set but does not describe this size (the in-CTA kernel is used instead).
static int big_plan(pdsp_plan* pl, BigPlan** out) {
  *out = nullptr;
  std::lock_guard<std::recursive_mutex> lk(pl->ctx->plan_mu);
  if (pl->big) {
    *out = pl->big;
    return 0;
  }
  const int n = pl->log2n;
  int lg[3] = {0, 0, 0}, np = 0;
  const Tune& tn = pl->ctx->tune;
  if (tn.n_big_factors) {  // test hook: forced pass lengths (log2)
    const int got = tn.n_big_factors;
    if (tn.big_factors[0] + tn.big_factors[1] + (got == 3 ? tn.big_factors[2] : 0) == n) {
      lg[0] = tn.big_factors[0], lg[1] = tn.big_factors[1], lg[2] = got == 3 ? tn.big_factors[2] : 0;
      np = got;
    } else if (n <= kMaxLog2M) {
      return 0;
    }
  }
  if (np == 0) {
    np = n <= 2 * kBigMaxLog2L ? 2 : 3;
    for (int j = 0; j < np; ++j) lg[j] = n / np + (j < n % np ? 1 : 0);
  }
  for (int j = 0; j < np; ++j)
    if (lg[j] < kBigMinLog2L || lg[j] > kBigMaxLog2L)
      return fail("FFT size 2^%d cannot be split into %d passes of 2^%d..2^%d points", n, np, kBigMinLog2L, kBigMaxLog2L);
  BigPlan* bp = new BigPlan();
  bp->npass = np;
  const size_t es = esize(pl->precision);
  for (int j = 0; j < np; ++j) bp->lg[j] = lg[j];
  int rest = n;  // log2(L_j * I_j) of pass j
  for (int j = 0; j + 1 < np; ++j) {
    bp->log_b[j] = (rest + 1) / 2;
    const int rc = pl->precision == PDSP_F64 ? upload_big_twiddles<double>(rest, bp->log_b[j], &bp->tw_hi[j], &bp->tw_lo[j])
                                             : upload_big_twiddles<float>(rest, bp->log_b[j], &bp->tw_hi[j], &bp->tw_lo[j]);
    if (rc) return 1;
    rest -= lg[j];
  }
  (void)es;
  pl->big = bp;
  *out = bp;
  return 0;
}

// Intermediate passes exchange interleaved (re, im) elements unless tune.big_interleave = 0: 512-byte instead of
// 2 x 256-byte tile rows (2^24: 0.375 -> 0.353 ms, profiles/r1/README.md)

// One pass of the first-generation path (bigfft_kernels.cuh: per-thread tile loads and stores, or TMA tile loads) over the
// `nf` transforms of a group.  fre/fim, gre/gim: the group's caller-facing input / output planes; wk: its work buffer.
static int launch_pass_v1(pdsp_plan* pl, BigPlan* bp, BigPlan::Work* wk, int j, long long O, long long I, long long nf,
                          const char* fre, const char* fim, char* gre, char* gim, int inverse, const LaunchCtx& lc,
                          bool pipe = false) {
  pdsp_ctx* c = pl->ctx;
  const size_t es = esize(pl->precision);
  const long long N = 1LL << pl->log2n;
  const int np = bp->npass;
  long long Ls[3] = {1, 1, 1};
  for (int k = 0; k < np; ++k) Ls[k] = 1LL << bp->lg[k];
  const long long L = Ls[j];
  const int C = big_pass_c(bp->lg[j]);
  const bool last = j == np - 1;
  BigPassParams p;
  memset(&p, 0, sizeof p);
  const bool interleave = wk->interleaved;
  p.in_re = j == 0 ? (const void*)fre : wk->re;
  p.in_im = j == 0 ? (const void*)fim : (interleave ? nullptr : wk->im);
  p.out_re = last ? (void*)gre : wk->re;
  p.out_im = last ? (void*)gim : (interleave ? nullptr : wk->im);
  p.in_cplx = (interleave && j != 0) ? 1 : 0;
  p.out_cplx = (interleave && !last) ? 1 : 0;
  p.n_frames = nf;
  p.in_frame = p.out_frame = N;
  p.swap_in = (j == 0 && inverse) ? 1 : 0;
  p.swap_out = (last && inverse) ? 1 : 0;
  p.scale = (last && inverse) ? 1.0 / (double)N : 1.0;
  p.l2_prefetch = c->tune.big_prefetch;  // default on: 2^24 0.43 -> 0.37 ms (profiles/r1/README.md)
  if (!last) {
    // view [O][L][I]: C adjacent inner indices per CTA, transform along the stride-I axis in place
    p.n_lo = I / C;
    p.n_groups = O * p.n_lo;
    p.in_hi = p.out_hi = L * I;
    p.in_lo = p.out_lo = C;
    p.in_c = p.out_c = 1;
    p.in_e = p.out_e = I;
    p.tw_hi = bp->tw_hi[j];
    p.tw_lo = bp->tw_lo[j];
    p.log_b = bp->log_b[j];
    p.stage_in = 0;
  } else if (np == 2) {
    // rows k1 (contiguous, L2 long); C adjacent k1 per CTA; X[k1 + L1*k2]
    p.n_lo = 1;
    p.n_groups = Ls[0] / C;
    p.in_hi = C * L;
    p.in_c = L;
    p.in_e = 1;
    p.out_hi = C;
    p.out_c = 1;
    p.out_e = Ls[0];
    p.stage_in = 1;
  } else {
    // rows (k1, k2); groups (k1 tile, k2); X[k1 + L1*k2 + L1*L2*k3]
    p.n_lo = Ls[1];
    p.n_groups = (Ls[0] / C) * Ls[1];
    p.in_hi = C * Ls[1] * L;
    p.in_lo = L;
    p.in_c = Ls[1] * L;
    p.in_e = 1;
    p.out_hi = C;
    p.out_lo = Ls[0];
    p.out_c = 1;
    p.out_e = Ls[0] * Ls[1];
    p.stage_in = 1;
  }
  cudaError_t e;
  // TMA tile loads for the strided passes while one group's planes fit the L2 (2^20: 0.158 vs 0.169 ms per 8
  // transforms); per-thread loads beyond that (2^24: 0.353 vs 0.373 ms).  Tunable big_tma forces either.
  const bool tma = c->tune.big_tma >= 0 ? c->tune.big_tma != 0 : (2 * es * (size_t)N * (size_t)nf <= ((size_t)128 << 20));
  if (pipe && last) {
    // third generation (bigfft3_kernels.cuh): the rows of the next tile land in shared memory by bulk copies while the
    // current one is transformed; same parameter block, the tensor maps are unused
    simt::TensorMap2D none;
    memset(&none, 0, sizeof none);
    e = launch_big_pipe(pl->precision == PDSP_F64, bp->lg[j], true, p, none, none, lc);
  } else if (!last && (tma || pipe)) {
    // TMA-staged tile gather: planes viewed as [frames*O*L rows][I cols], box {C, min(L, 256)}
    simt::TensorMap2D tm_re, tm_im;
    int bc = 0, br = 0;
    big_pass_tma_box(bp->lg[j], &bc, &br);
    const bool f64p = pl->precision == PDSP_F64;
    if (p.in_cplx) {
      // interleaved work buffer viewed as [rows][2*I] scalars; {2*C, rows/2} boxes keep a box at 64 KB
      if (make_tensor_map(&tm_re, p.in_re, f64p, 2 * I, nf * O * L, 2 * bc, br / 2)) return 1;
      tm_im = tm_re;
    } else {
      if (make_tensor_map(&tm_re, p.in_re, f64p, I, nf * O * L, bc, br)) return 1;
      if (make_tensor_map(&tm_im, p.in_im ? p.in_im : p.in_re, f64p, I, nf * O * L, bc, br)) return 1;
    }
    e = pipe ? launch_big_pipe(f64p, bp->lg[j], false, p, tm_re, tm_im, lc) : launch_big_pass_tma(f64p, bp->lg[j], p, tm_re, tm_im, lc);
  } else {
    e = launch_big_pass(pl->precision == PDSP_F64, bp->lg[j], p, lc);
  }
  if (e != cudaSuccess) return fail("big FFT pass %d (n=2^%d): %s", j, pl->log2n, cudaGetErrorString(e));
  c->launches++;
  return 0;
}

// Rank-3 tensor map: element (x, y, z) at base + (x + y*stride1 + z*stride2) elements, box {b0, b1, b2}.
static int make_tensor_map3(simt::TensorMap* tm, const void* base, bool f64, const long long dims[3], long long stride1,
                            long long stride2, const int box[3]) {
  const long long es = f64 ? 8 : 4;
#ifdef PDSP_EMU
  tm->base = const_cast<void*>(base);
  for (int d = 0; d < 3; ++d) tm->dim[d] = dims[d], tm->box[d] = box[d];
  tm->stride_bytes[0] = stride1 * es;
  tm->stride_bytes[1] = stride2 * es;
  tm->esize = (int)es;
  return 0;
#else
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled is not available from this driver");
    encode = reinterpret_cast<encode_fn>(fn);
  }
  const cuuint64_t gdims[3] = {(cuuint64_t)dims[0], (cuuint64_t)dims[1], (cuuint64_t)dims[2]};
  const cuuint64_t gstr[2] = {(cuuint64_t)(stride1 * es), (cuuint64_t)(stride2 * es)};
  const cuuint32_t gbox[3] = {(cuuint32_t)box[0], (cuuint32_t)box[1], (cuuint32_t)box[2]};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode(tm, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base),
                            gdims, gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeTiled (rank 3, dims %lld x %lld x %lld, box %d x %d x %d) failed: CUresult %d", dims[0], dims[1],
                dims[2], box[0], box[1], box[2], (int)r);
  return 0;
#endif
}

// One pass of the second-generation path (bigfft2_kernels.cuh) over `nf` transforms whose planes start at the given
// pointers.  k_cnt > 0 (three-pass plans, passes 1 and 2, nf == 1) restricts the pass to the k1 range [k0, k0 + k_cnt):
// the caller then runs passes 1 and 2 group by group so that a group's work-buffer slice never leaves the L2.
static int launch_pass_v2(pdsp_plan* pl, BigPlan* bp, int j, long long O, long long I, long long nf, const void* in_re,
                          const void* in_im, bool in_cplx, void* out_re, void* out_im, bool out_cplx, int inverse,
                          const LaunchCtx& lc, long long k0, long long k_cnt, BigTileParams* cap_p = nullptr,
                          simt::TensorMap* cap_maps = nullptr) {
  pdsp_ctx* c = pl->ctx;
  const bool f64 = pl->precision == PDSP_F64;
  const size_t es = esize(pl->precision);
  const long long N = 1LL << pl->log2n;
  const int np = bp->npass;
  long long Ls[3] = {1, 1, 1};
  for (int k = 0; k < np; ++k) Ls[k] = 1LL << bp->lg[k];
  const long long L = Ls[j];
  const int C = big2_pass_c(bp->lg[j]);
  const bool last = j == np - 1;
  const int BR = (int)(L < 256 ? L : 256);
  BigTileParams p;
  memset(&p, 0, sizeof p);
  simt::TensorMap maps[4];
  p.n_frames = nf;
  p.swap_in = (j == 0 && inverse) ? 1 : 0;
  p.swap_out = (last && inverse) ? 1 : 0;
  p.scale = (last && inverse) ? 1.0 / (double)N : 1.0;
  p.has_im = in_im != nullptr ? 1 : 0;
  p.l2_prefetch = c->tune.big_prefetch;
  const size_t ies = es * (in_cplx ? 2 : 1), oes = es * (out_cplx ? 2 : 1);  // bytes per element of the in / out buffers
  int io;
  if (!last) {
    // view [nf*O][L][I]: tile = C adjacent columns (box {C, BR, 1}, L / BR boxes down the transformed axis)
    if (I % C) return fail("internal: pass %d of 2^%d has %lld columns, not a multiple of the tile width %d", j, pl->log2n, I, C);
    p.n_lo = I / C;
    const long long o_cnt = k_cnt > 0 ? k_cnt : O;
    p.n_groups = o_cnt * p.n_lo;
    p.tw_hi = bp->tw_hi[j];
    p.tw_lo = bp->tw_lo[j];
    p.log_b = bp->log_b[j];
    const int wi = in_cplx ? 2 : 1, wo = out_cplx ? 2 : 1;  // scalars per element in the mapped buffer
    const size_t skip = (size_t)k0 * (size_t)(L * I);      // elements before the first k1 block of the range
    const char* bin_re = static_cast<const char*>(in_re) + skip * ies;
    const char* bin_im = in_im ? static_cast<const char*>(in_im) + skip * ies : bin_re;
    char* bout_re = static_cast<char*>(out_re) + skip * oes;
    char* bout_im = out_im ? static_cast<char*>(out_im) + skip * oes : bout_re;
    const long long din[3] = {wi * I, L, nf * o_cnt}, dout[3] = {wo * I, L, nf * o_cnt};
    const int bin[3] = {wi * C, BR, 1}, bout[3] = {wo * C, BR, 1};
    if (make_tensor_map3(&maps[0], bin_re, f64, din, wi * I, wi * I * L, bin)) return 1;
    if (make_tensor_map3(&maps[1], bin_im, f64, din, wi * I, wi * I * L, bin)) return 1;
    if (make_tensor_map3(&maps[2], bout_re, f64, dout, wo * I, wo * I * L, bout)) return 1;
    if (make_tensor_map3(&maps[3], bout_im, f64, dout, wo * I, wo * I * L, bout)) return 1;
    p.in_lo[0] = wi * C, p.in_hi[2] = 1, p.in_fr[2] = (int)o_cnt;
    p.in_box_dim = 1, p.in_box_step = BR, p.in_boxes = (int)(L / BR);
    p.in_box_bytes = (unsigned)(wi * C * BR * es);
    p.out_lo[0] = wo * C, p.out_hi[2] = 1, p.out_fr[2] = (int)o_cnt;
    p.out_box_dim = 1, p.out_box_step = BR, p.out_boxes = (int)(L / BR);
    p.out_box_bytes = (unsigned)(wo * C * BR * es);
    io = (in_cplx ? 1 : 0) | (out_cplx ? 2 : 0);
  } else {
    // rows (k1[, k2]), L contiguous elements each; output X[k1 + L1*k2 (+ L1*L2*k3)]: view {L1, L2', nf*L}, box {C, 1, BR}
    const long long L1 = Ls[0], L2 = np == 3 ? Ls[1] : 1;
    const long long k1_cnt = k_cnt > 0 ? k_cnt : L1;
    if (k1_cnt % C || k0 % C) return fail("internal: last pass of 2^%d: k1 range [%lld, +%lld) is not a multiple of the tile height %d", pl->log2n, k0, k1_cnt, C);
    p.n_lo = L2;
    p.n_groups = (k1_cnt / C) * L2;
    p.in_re = static_cast<const char*>(in_re) + (size_t)k0 * (size_t)(L2 * L) * ies;
    p.in_im = in_im ? static_cast<const char*>(in_im) + (size_t)k0 * (size_t)(L2 * L) * ies : nullptr;
    p.in_frame = N;
    p.in_g_hi = C * L2 * L;
    p.in_g_lo = np == 3 ? L : 0;
    p.in_c = L2 * L;
    const long long dout[3] = {L1 - k0, L2, nf * L};
    const int bout[3] = {C, 1, BR};
    if (make_tensor_map3(&maps[2], static_cast<char*>(out_re) + (size_t)k0 * es, f64, dout, L1, L1 * L2, bout)) return 1;
    if (make_tensor_map3(&maps[3], static_cast<char*>(out_im) + (size_t)k0 * es, f64, dout, L1, L1 * L2, bout)) return 1;
    maps[0] = maps[2], maps[1] = maps[3];
    p.out_lo[1] = 1, p.out_hi[0] = C, p.out_fr[2] = (int)L;
    p.out_box_dim = 2, p.out_box_step = BR, p.out_boxes = (int)(L / BR);
    p.out_box_bytes = (unsigned)(C * BR * es);
    io = 4 | (in_cplx ? 1 : 0);
  }
  if (cap_p != nullptr) {  // the fused launch takes the parameter block and the maps instead
    *cap_p = p;
    for (int k = 0; k < 4; ++k) cap_maps[k] = maps[k];
    return 0;
  }
  const cudaError_t e = launch_big_tile(f64, bp->lg[j], io, p, maps, lc);
  if (e != cudaSuccess) return fail("big FFT pass %d (n=2^%d, TMA tiles): %s", j, pl->log2n, cudaGetErrorString(e));
  c->launches++;
  return 0;
}

// Multi-pass transform.  Each pass runs on the generation that is faster for its length (both exchange the same work
// buffer): the second (bigfft2_kernels.cuh: TMA box loads and box stores, two CTAs per SM, transposed box stores in the
// last pass) wherever its tiles have rows of at least 64 bytes - passes of up to 256 points; measured 3-10 % ahead at
// 2^16, 2^22 and 2^24 - and the first (bigfft_kernels.cuh) for the 512- and 1024-point passes, whose 4- and 8-column
// tiles make 32-byte box rows (the TMA unit then spends its time on row requests: 2^20 = 1024 x 1024 ran 0.19 of the
// roofline against 0.26).  `chunk` transforms go through all passes together.
#ifndef PDSP_BIG_FUSED_DEFAULT
#define PDSP_BIG_FUSED_DEFAULT false
#endif
static int launch_big(pdsp_plan* pl, BigPlan* bp, const void* d_re, const void* d_im, long long batch, void* d_ore,
                      void* d_oim, int inverse, cudaStream_t st, bool v2_allowed) {
  pdsp_ctx* c = pl->ctx;
  const bool f64 = pl->precision == PDSP_F64;
  const size_t es = esize(pl->precision);
  const long long N = 1LL << pl->log2n;
  LaunchCtx lc{c->device, c->sm_count, st, pass_twiddles_cb, c};
  const int np = bp->npass;
  long long Ls[3] = {1, 1, 1};
  for (int j = 0; j < np; ++j) Ls[j] = 1LL << bp->lg[j];
  long long chunk = (1LL << 25) / N;  // <= 2^25 elements (256 MB of doubles) per work plane
  if (c->tune.big_chunk > 0) chunk = c->tune.big_chunk;  // experiment / test hook: transforms per group of passes
  if (chunk < 1) chunk = 1;
  if (chunk > batch) chunk = batch;
  const bool il = c->tune.big_interleave != 0;
  BigPlan::Work* wk = nullptr;
  {
    std::lock_guard<std::recursive_mutex> lk(c->plan_mu);
    wk = &bp->work[st];
    if (wk->frames < chunk || wk->interleaved != il) {
      CU(cudaStreamSynchronize(st));
      CU(cudaFree(wk->re));
      CU(cudaFree(wk->im));
      wk->re = wk->im = nullptr;
      CU(cudaMalloc(&wk->re, (il ? 2 : 1) * es * (size_t)N * (size_t)chunk));
      if (!il) CU(cudaMalloc(&wk->im, es * (size_t)N * (size_t)chunk));
      wk->interleaved = il;
      wk->frames = chunk;
    }
  }
  for (long long f0 = 0; f0 < batch; f0 += chunk) {
    const long long nf = batch - f0 < chunk ? batch - f0 : chunk;
    const char* fre = static_cast<const char*>(d_re) + (size_t)f0 * N * es;
    const char* fim = d_im ? static_cast<const char*>(d_im) + (size_t)f0 * N * es : nullptr;
    char* gre = static_cast<char*>(d_ore) + (size_t)f0 * N * es;
    char* gim = static_cast<char*>(d_oim) + (size_t)f0 * N * es;
    long long O = 1, I = N;
    // generation of each pass: tunable big_v2 = 0 / 1 forces the first / second (where TMA can take the shape at all)
    bool pass_v2[3] = {false, false, false};
    for (int j = 0; j < np; ++j) {
      const int Cj = big2_pass_c(bp->lg[j]);
      pass_v2[j] = v2_allowed && (size_t)Cj * es >= 16 && (c->tune.big_v2 < 0 ? (size_t)Cj * es >= 64 : c->tune.big_v2 != 0);
    }
    // third generation (pipeline passes, bigfft3_kernels.cuh): fp64 passes of 256 / 512 / 1024 points over the
    // interleaved work buffer.  Automatic for 1024-point passes, whose tiles are too large for two CTAs per SM, and for
    // the first pass of three-pass transforms (planar caller input: the full-width tile reads 256-byte rows where the
    // second generation's half tile reads 128 - 2^24 330 -> 328 us, 2^26 1.40 -> 1.36 ms, profiles/r2/sweep_big_pipe_mask.txt);
    // tunable big_pipe = 0 / 1 turns it off / on for every length it exists for, 2 + m selects passes by bit mask.
    bool pass_v3[3] = {false, false, false};
    for (int j = 0; j < np; ++j) {
      pass_v3[j] = il && v2_allowed && big_pipe_supported(f64, bp->lg[j]) &&
                   (c->tune.big_pipe < 0 ? (bp->lg[j] == 10 || (j == 0 && np == 3)) : (c->tune.big_pipe >= 2 ? (((c->tune.big_pipe - 2) >> j) & 1) != 0 : c->tune.big_pipe != 0));
      if (pass_v3[j]) pass_v2[j] = false;
    }
    // Three-pass transforms whose work buffer exceeds the L2 (2^24: 256 MB): after pass 0, passes 1 and 2 run over GROUPS
    // of k1 blocks - pass 1 writes a group's slice of the work buffer (<= 32 MB), pass 2 reads it back while it is still
    // in the L2 and stores the final (transposed) output.  DRAM then sees the work buffer once between passes 0 and 1
    // only: 4 x N*16 bytes instead of 6.  MEASURED SLOWER and therefore off unless the tunable big_resident asks for it
    // (2^24: 0.331 ms pass by pass, 0.404-0.510 ms in groups of 64-16 blocks; 2^26: 1.40 vs 1.55-1.96 ms -
    // profiles/r2/sweep_big_resident.txt): a group launch moves only 16-64 MB, so every CTA handles 1-2 tiles and the ramp
    // and tail of 2 x (L1 / group) launches cost more than the saved DRAM round trip - the passes are bound by the
    // per-tile latency chain, not by DRAM.
    long long group = 0;
    if (np == 3 && il && pass_v2[0] && pass_v2[1] && pass_v2[2] && c->tune.big_resident > 0) {
      const size_t block = (size_t)Ls[1] * (size_t)Ls[2] * 2 * es;  // one k1 block of the interleaved work buffer
      const long long c_last = big2_pass_c(bp->lg[2]);
      long long g = c_last;
      while (g * 2 <= Ls[0] && (size_t)(g * 2) * block <= ((size_t)32 << 20)) g *= 2;
      if ((size_t)Ls[0] * block > ((size_t)64 << 20) && g < Ls[0] && Ls[0] % g == 0) group = g;
      if (c->tune.big_resident > 1 && Ls[0] % c->tune.big_resident == 0 && c->tune.big_resident % c_last == 0)
        group = c->tune.big_resident;  // explicit group size (k1 blocks)
    }
    // Fused form of the same idea: ONE persistent launch runs the middle and the last pass, a last-pass tile waiting
    // (acquire on a per-group counter) until the middle tiles of its C2 k1 blocks have been stored - no launch per group,
    // no ramp and tail (bigfft_fused_kernel).
    const bool fused = np == 3 && il && pass_v2[1] && pass_v2[2] && big_fused_supported(f64, bp->lg[1], bp->lg[2]) &&
                       (c->tune.big_fused < 0 ? PDSP_BIG_FUSED_DEFAULT : c->tune.big_fused != 0) && Ls[0] / big2_pass_c(bp->lg[2]) <= 1024;
    if (fused) {
      if (!wk->sync) {
        std::lock_guard<std::recursive_mutex> lk(c->plan_mu);
        if (!wk->sync) {
          void* sp = nullptr;
          CU(cudaMalloc(&sp, 1025 * sizeof(unsigned)));
          wk->sync = static_cast<unsigned*>(sp);
        }
      }
      for (long long f = 0; f < nf; ++f) {
        const char* xre = fre + (size_t)f * N * es;
        const char* xim = fim ? fim + (size_t)f * N * es : nullptr;
        char* yre = gre + (size_t)f * N * es;
        char* yim = gim + (size_t)f * N * es;
        char* wbuf = static_cast<char*>(wk->re) + (size_t)f * N * 2 * es;
        const long long I0 = N / Ls[0], I1 = I0 / Ls[1];
        if (pass_v2[0]) {
          if (launch_pass_v2(pl, bp, 0, 1, I0, 1, xre, xim, false, wbuf, nullptr, true, inverse, lc, 0, 0)) return 1;
        } else {
          // first or third generation: they address the group's work buffer through wk, so point a one-frame view at it
          BigPlan::Work one = *wk;
          one.re = wbuf;
          if (launch_pass_v1(pl, bp, &one, 0, 1, I0, 1, xre, xim, yre, yim, inverse, lc, pass_v3[0])) return 1;
        }
        BigTileParams pa, pb;
        simt::TensorMap ma[4], mb[4];
        if (launch_pass_v2(pl, bp, 1, Ls[0], I1, 1, wbuf, nullptr, true, wbuf, nullptr, true, inverse, lc, 0, 0, &pa, ma)) return 1;
        if (launch_pass_v2(pl, bp, 2, Ls[0] * Ls[1], 1, 1, wbuf, nullptr, true, yre, yim, false, inverse, lc, 0, 0, &pb, mb)) return 1;
        BigFusedSync fs;
        fs.n_groups = Ls[0] / big2_pass_c(bp->lg[2]);
        fs.n1g = pa.n_groups / fs.n_groups;
        fs.n2g = pb.n_groups / fs.n_groups;
        fs.skew = c->tune.big_fused > 1 ? c->tune.big_fused : 1;
        fs.counters = wk->sync;
        fs.error = wk->sync + 1024;
        if (pa.n_groups % fs.n_groups || pb.n_groups % fs.n_groups) return fail("internal: fused passes of 2^%d do not tile into %lld groups", pl->log2n, fs.n_groups);
        CU(cudaMemsetAsync(wk->sync, 0, 1025 * sizeof(unsigned), st));
        const cudaError_t e = launch_big_fused(bp->lg[1], bp->lg[2], pa, pb, fs, ma, mb, lc);
        if (e != cudaSuccess) return fail("big FFT fused passes (n=2^%d): %s", pl->log2n, cudaGetErrorString(e));
        c->launches++;
      }
      continue;
    }
    if (group > 0) {
      for (long long f = 0; f < nf; ++f) {
        const char* xre = fre + (size_t)f * N * es;
        const char* xim = fim ? fim + (size_t)f * N * es : nullptr;
        char* yre = gre + (size_t)f * N * es;
        char* yim = gim + (size_t)f * N * es;
        char* wbuf = static_cast<char*>(wk->re) + (size_t)f * N * 2 * es;
        const long long I0 = N / Ls[0], I1 = I0 / Ls[1];
        if (launch_pass_v2(pl, bp, 0, 1, I0, 1, xre, xim, false, wbuf, nullptr, true, inverse, lc, 0, 0)) return 1;
        for (long long k0 = 0; k0 < Ls[0]; k0 += group) {
          if (launch_pass_v2(pl, bp, 1, Ls[0], I1, 1, wbuf, nullptr, true, wbuf, nullptr, true, inverse, lc, k0, group)) return 1;
          if (launch_pass_v2(pl, bp, 2, Ls[0] * Ls[1], 1, 1, wbuf, nullptr, true, yre, yim, false, inverse, lc, k0, group)) return 1;
        }
      }
      continue;
    }
    for (int j = 0; j < np; ++j) {
      const long long L = Ls[j];
      I /= L;
      const bool last = j == np - 1;
      const bool v2 = pass_v2[j];
      if (!v2) {
        if (launch_pass_v1(pl, bp, wk, j, O, I, nf, fre, fim, gre, gim, inverse, lc, pass_v3[j])) return 1;
        O *= L;
        continue;
      }
      const bool in_cplx = il && j != 0, out_cplx = il && !last;
      const void* in_re = j == 0 ? (const void*)fre : wk->re;
      const void* in_im = j == 0 ? (const void*)fim : (il ? nullptr : wk->im);
      void* out_re = last ? (void*)gre : wk->re;
      void* out_im = last ? (void*)gim : (il ? nullptr : wk->im);
      if (launch_pass_v2(pl, bp, j, O, I, nf, in_re, in_im, in_cplx, out_re, out_im, out_cplx, inverse, lc, 0, 0)) return 1;
      O *= L;
    }
  }
  return 0;
}

static int launch_c2c(pdsp_plan* pl, const void* d_re, const void* d_im, long long batch, void* d_ore, void* d_oim,
                      int inverse, cudaStream_t st, const Doorbell* door, bool* door_used) {
  pdsp_ctx* c = pl->ctx;
  if (door_used) *door_used = false;
  if (pl->log2n > kMaxLog2M || c->tune.n_big_factors) {
    BigPlan* bp = nullptr;
    if (big_plan(pl, &bp)) return 1;
    if (bp) {
      // TMA box stores need 16-byte aligned planes; anything else stays on the first generation
      const uintptr_t al = reinterpret_cast<uintptr_t>(d_re) | reinterpret_cast<uintptr_t>(d_im) | reinterpret_cast<uintptr_t>(d_ore) |
                           reinterpret_cast<uintptr_t>(d_oim);
      return launch_big(pl, bp, d_re, d_im, batch, d_ore, d_oim, inverse, st, (al & 15u) == 0);
    }
  }
  C2CParams p;
  memset(&p, 0, sizeof p);
  p.in_re = d_re;
  p.in_im = d_im;
  p.out_re = d_ore;
  p.out_im = d_oim;
  p.batch = batch;
  p.inverse = inverse;
  if (door) {
    p.door = *door;
    if (door_used) *door_used = true;
  }
  LaunchCtx lc{c->device, c->sm_count, st, pass_twiddles_cb, c};
  cudaError_t e = dispatch_c2c(pl->precision == PDSP_F64, pl->log2n, p, lc);
  if (e != cudaSuccess) return fail("c2c launch (n=%d): %s", pl->n, cudaGetErrorString(e));
  c->launches++;
  return 0;
}

// spectrum() of frames longer than one CTA can hold.  Frames are processed in sub-batches through per-stream
// scratch planes (windowed frames, full complex spectra).
static int launch_spectrum_big(pdsp_plan* pl, const R2CParams& p, long long batch, cudaStream_t st) {
  pdsp_ctx* c = pl->ctx;
  BigPlan* bp = nullptr;
  if (big_plan(pl, &bp)) return 1;
  if (!bp) return fail("internal: no multi-pass plan for n=%d", pl->n);
  const int n = pl->n;
  const size_t es = esize(pl->precision);
  long long sub = (1LL << 24) / n;  // <= 2^24 elements per scratch plane
  if (sub < 1) sub = 1;
  if (sub > batch) sub = batch;
  BigPlan::Work* wk = nullptr;
  {
    std::lock_guard<std::recursive_mutex> lk(c->plan_mu);
    wk = &bp->work[st];
    if (wk->spec_frames < sub) {
      CU(cudaStreamSynchronize(st));
      CU(cudaFree(wk->spec_x));
      CU(cudaFree(wk->spec_re));
      CU(cudaFree(wk->spec_im));
      wk->spec_x = wk->spec_re = wk->spec_im = nullptr;
      wk->spec_frames = 0;
      CU(cudaMalloc(&wk->spec_x, es * (size_t)n * (size_t)sub));
      CU(cudaMalloc(&wk->spec_re, es * (size_t)n * (size_t)sub));
      CU(cudaMalloc(&wk->spec_im, es * (size_t)n * (size_t)sub));
      wk->spec_frames = sub;
    }
  }
  const int bins = p.two_sided ? n : n / 2 + 1;
  const size_t ses = p.sample_dtype == DT_F64 ? 8 : 4;
  const size_t pk = pl->precision == PDSP_F64 ? sizeof(PeakRec<double>) : sizeof(PeakRec<float>);
  const int threads = 256;
  for (long long f0 = 0; f0 < batch; f0 += sub) {
    const long long nf = batch - f0 < sub ? batch - f0 : sub;
    const char* src = static_cast<const char*>(p.samples) + (size_t)f0 * (size_t)p.hop * ses;
    long long blocks = (nf * (long long)n + threads - 1) / threads;
    if (blocks > (long long)c->sm_count * 16) blocks = (long long)c->sm_count * 16;
    if (pl->precision == PDSP_F64) {
      if (p.sample_dtype == DT_F64)
        PDSP_LAUNCH((k_build_frames<double, double>), (int)blocks, threads, 0, st, reinterpret_cast<const double*>(src), p.hop,
                    p.frame_len, static_cast<const double*>(p.window), n, nf, static_cast<double*>(wk->spec_x));
      else
        PDSP_LAUNCH((k_build_frames<double, float>), (int)blocks, threads, 0, st, reinterpret_cast<const float*>(src), p.hop,
                    p.frame_len, static_cast<const double*>(p.window), n, nf, static_cast<double*>(wk->spec_x));
    } else {
      if (p.sample_dtype == DT_F64)
        PDSP_LAUNCH((k_build_frames<float, double>), (int)blocks, threads, 0, st, reinterpret_cast<const double*>(src), p.hop,
                    p.frame_len, static_cast<const float*>(p.window), n, nf, static_cast<float*>(wk->spec_x));
      else
        PDSP_LAUNCH((k_build_frames<float, float>), (int)blocks, threads, 0, st, reinterpret_cast<const float*>(src), p.hop,
                    p.frame_len, static_cast<const float*>(p.window), n, nf, static_cast<float*>(wk->spec_x));
    }
    CU(cudaGetLastError());
    c->launches++;
    if (launch_c2c(pl, wk->spec_x, nullptr, nf, wk->spec_re, wk->spec_im, 0, st)) return 1;
    long long eb = nf < (long long)c->sm_count * 8 ? nf : (long long)c->sm_count * 8;
    const size_t smem = 32 * 8 + 32 * sizeof(int);
    char* amp = p.amp ? static_cast<char*>(p.amp) + (size_t)f0 * bins * es : nullptr;
    char* ph = p.phase ? static_cast<char*>(p.phase) + (size_t)f0 * bins * es : nullptr;
    char* pkp = p.peaks ? static_cast<char*>(p.peaks) + (size_t)f0 * pk : nullptr;
    if (pl->precision == PDSP_F64)
      PDSP_LAUNCH(k_big_epilogue<double>, (int)eb, threads, smem, st, static_cast<const double*>(wk->spec_re),
                  static_cast<const double*>(wk->spec_im), n, bins, p.scale_edge, p.scale_mid, p.two_sided ? 0 : 1, p.shift, p.bin_hz, nf,
                  reinterpret_cast<double*>(amp), reinterpret_cast<double*>(ph), reinterpret_cast<PeakRec<double>*>(pkp));
    else
      PDSP_LAUNCH(k_big_epilogue<float>, (int)eb, threads, smem, st, static_cast<const float*>(wk->spec_re),
                  static_cast<const float*>(wk->spec_im), n, bins, (float)p.scale_edge, (float)p.scale_mid,
                  p.two_sided ? 0 : 1, p.shift, p.bin_hz, nf, reinterpret_cast<float*>(amp), reinterpret_cast<float*>(ph),
                  reinterpret_cast<PeakRec<float>*>(pkp));
    CU(cudaGetLastError());
    c->launches++;
  }
  return 0;
}

// ------------------------------------------------------------------------------ staging pipeline
// Moves `chunks` pieces of a host job through HBM: H2D (direct DMA when the caller's memory is
// pinned, through the slot's pinned buffer otherwise), kernel, D2H, on kSlots streams so the
// copies of neighbouring chunks overlap the kernel.
struct HostIO {
  const void* src;  // host input for this chunk
  size_t src_bytes;
  struct Out {
    void* dst;
    size_t bytes;
    size_t d_off;  // offset inside slot.d_out
  };
  Out outs[4];
  int n_outs;
};

static int slot_wait(Slot& s) {
  if (!s.busy) return 0;
  CU(cudaEventSynchronize(s.done));
  for (auto& f : s.flush) memcpy(f.dst, f.src, f.bytes);
  s.flush.clear();
  s.busy = false;
  return 0;
}

static int slot_in(Slot& s, const void* src, size_t bytes, bool src_pinned, size_t d_extra = 0) {
  if (ensure(&s.d_in, &s.d_in_cap, bytes + d_extra, false)) return 1;
  if (bytes == 0) return 0;
  if (src_pinned) {
    CU(cudaMemcpyAsync(s.d_in, src, bytes, cudaMemcpyHostToDevice, s.stream));
  } else {
    if (ensure(&s.h_in, &s.h_in_cap, bytes, true)) return 1;
    memcpy(s.h_in, src, bytes);
    CU(cudaMemcpyAsync(s.d_in, s.h_in, bytes, cudaMemcpyHostToDevice, s.stream));
  }
  return 0;
}

static int slot_out(Slot& s, const HostIO::Out* outs, int n, const bool* pinned) {
  size_t stage = 0;
  for (int i = 0; i < n; ++i)
    if (!pinned[i]) stage += (outs[i].bytes + 255) & ~(size_t)255;
  if (stage && ensure(&s.h_out, &s.h_out_cap, stage, true)) return 1;
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    if (outs[i].bytes == 0) continue;
    const char* dsrc = static_cast<const char*>(s.d_out) + outs[i].d_off;
    if (pinned[i]) {
      CU(cudaMemcpyAsync(outs[i].dst, dsrc, outs[i].bytes, cudaMemcpyDeviceToHost, s.stream));
    } else {
      char* h = static_cast<char*>(s.h_out) + off;
      CU(cudaMemcpyAsync(h, dsrc, outs[i].bytes, cudaMemcpyDeviceToHost, s.stream));
      s.flush.push_back({outs[i].dst, h, outs[i].bytes});
      off += (outs[i].bytes + 255) & ~(size_t)255;
    }
  }
  CU(cudaEventRecord(s.done, s.stream));
  s.busy = true;
  return 0;
}

static int drain(pdsp_ctx* c) {
  for (int i = 0; i < kSlots; ++i)
    if (slot_wait(c->slots[i])) return 1;
  return 0;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static long long pick_chunk(const pdsp_ctx* ctx, long long batch, size_t bytes_per_frame) {
  // ~24 MB of traffic per chunk keeps three chunks in flight without hoarding HBM or pinned memory
  size_t target = 24u << 20;
  if (ctx->tune.chunk_bytes > 0) target = (size_t)ctx->tune.chunk_bytes;  // test hook: small jobs still exercise the multi-chunk pipeline
  long long c = (long long)(target / (bytes_per_frame ? bytes_per_frame : 1));
  if (c < 1) c = 1;
  if (c > batch) c = batch;
  // small jobs: still split in two so H2D of the second half overlaps the first kernel
  return c;
}

static int fast_init(pdsp_ctx* c);

// ------------------------------------------------------------------------------ C ABI
PDSP_EXPORT int pdsp_abi_version(void) { return PDSP_ABI_VERSION; }
PDSP_EXPORT const char* pdsp_last_error(void) { return g_err.c_str(); }

PDSP_EXPORT int pdsp_device_count(int* count) {
  if (!count) return fail("null argument");
  CU(cudaGetDeviceCount(count));
  return 0;
}

PDSP_EXPORT int pdsp_ctx_create(int device, pdsp_ctx** out) {
  if (!out) return fail("null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0)
    return fail("no CUDA device available (%s); this library has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev || device >= kMaxDevices) return fail("device %d out of range (%d devices)", device, ndev);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail("device %d is sm_%d%d; this build targets sm_100a (B200)", device, prop.major, prop.minor);
  pdsp_ctx* c = new pdsp_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  tune_from_env(c->tune);
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (int i = 0; i < kSlots; ++i) {
    CU(cudaStreamCreateWithFlags(&c->slots[i].stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->slots[i].done, cudaEventDisableTiming));
  }
  if (fast_init(c)) return 1;
  *out = c;
  return 0;
}

PDSP_EXPORT int pdsp_ctx_destroy(pdsp_ctx* c) {
  if (!c) return 0;
  if (set_device(c)) return 1;
  cudaDeviceSynchronize();
  for (auto& kv : c->plans) {
    pdsp_plan* pl = kv.second;
    cudaFree(pl->d_post);
    for (int i = 0; i < 4; ++i) cudaFree(pl->d_win[i]);
    if (pl->big) {
      for (int i = 0; i < 2; ++i) {
        cudaFree(pl->big->tw_hi[i]);
        cudaFree(pl->big->tw_lo[i]);
      }
      for (auto& kv : pl->big->work) {
        cudaFree(kv.second.re);
        cudaFree(kv.second.im);
        cudaFree(kv.second.sync);
        cudaFree(kv.second.spec_x);
        cudaFree(kv.second.spec_re);
        cudaFree(kv.second.spec_im);
      }
      delete pl->big;
    }
    delete pl;
  }
  for (auto& kv : c->pass_tw) cudaFree(kv.second);
  for (int i = 0; i < kSlots; ++i) {
    Slot& s = c->slots[i];
    cudaFree(s.d_in);
    cudaFree(s.d_out);
    cudaFreeHost(s.h_in);
    cudaFreeHost(s.h_out);
    cudaEventDestroy(s.done);
    cudaStreamDestroy(s.stream);
  }
  cudaFreeHost(c->fast.h);
  cudaFree(c->fast.d_count);
  if (c->fast.stream) cudaStreamDestroy(c->fast.stream);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

PDSP_EXPORT int pdsp_ctx_sync(pdsp_ctx* c) {
  if (!c) return fail("null context");
  if (set_device(c)) return 1;
  if (drain(c)) return 1;
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
PDSP_EXPORT int pdsp_ctx_tune(pdsp_ctx* c, const char* key, const char* value) {
  if (!c || !key) return fail("null argument");
  std::lock_guard<std::mutex> lk(c->mu);
  if (tune_set(c->tune, key, value)) return fail("unknown tunable '%s'", key);
  return 0;
}
PDSP_EXPORT int pdsp_ctx_device(const pdsp_ctx* c) { return c ? c->device : -1; }
PDSP_EXPORT int pdsp_ctx_sm_count(const pdsp_ctx* c) { return c ? c->sm_count : 0; }
PDSP_EXPORT int64_t pdsp_ctx_launch_count(const pdsp_ctx* c) { return c ? (int64_t)c->launches.load() : 0; }
PDSP_EXPORT int64_t pdsp_ctx_fast_call_count(const pdsp_ctx* c) { return c ? (int64_t)c->fast_calls.load() : 0; }

PDSP_EXPORT int pdsp_is_power_of_two(int32_t n) { return n > 0 && (n & (n - 1)) == 0; }
PDSP_EXPORT int32_t pdsp_next_power_of_two(int32_t n) {
  if (n <= 1) return 1;
  int32_t p = 1;
  while (p < n) p <<= 1;
  return p;
}
PDSP_EXPORT int pdsp_create_window(int window, int32_t size, double* out) {
  if (!out && size > 0) return fail("null argument");
  return window_host(window, size, out);
}
PDSP_EXPORT int pdsp_bin_frequencies(int32_t size, double sample_rate, int sides, double* out, int32_t* bins) {
  if (size <= 0) return fail("FFT size must be positive, got %d", size);
  if (!(sample_rate > 0)) return fail("Sample rate must be positive, got %g", sample_rate);
  const int32_t b = sides == PDSP_SIDES_ONE ? size / 2 + 1 : size;
  if (bins) *bins = b;
  if (out) {
    const double scale = sample_rate / size;
    for (int32_t i = 0; i < b; ++i) out[i] = i * scale;
  }
  return 0;
}

PDSP_EXPORT int pdsp_plan_get(pdsp_ctx* c, int32_t size, int precision, pdsp_plan** out) {
  if (!c || !out) return fail("null argument");
  *out = nullptr;
  if (!pdsp_is_power_of_two(size)) return fail("FFT size must be power of two, got %d", size);
  if (precision != PDSP_F32 && precision != PDSP_F64) return fail("unknown precision %d", precision);
  int log2n = 0;
  while ((1 << log2n) < size) ++log2n;
  if (log2n > kMaxBigLog2N) return fail("FFT size %d exceeds the supported maximum 2^%d", size, kMaxBigLog2N);
  if (set_device(c)) return 1;
  std::lock_guard<std::recursive_mutex> lk(c->plan_mu);
  auto key = std::make_pair((int)size, precision);
  auto it = c->plans.find(key);
  if (it != c->plans.end()) {
    *out = it->second;
    return 0;
  }
  pdsp_plan* pl = new pdsp_plan();
  pl->ctx = c;
  pl->n = size;
  pl->log2n = log2n;
  pl->precision = precision;
  // single-CTA tables exist for every size the r2c (N <= 16384) or c2c (N <= 8192) kernels can take
  int rc = 0;
  if (log2n - 1 <= kMaxLog2M) rc = precision == PDSP_F64 ? upload_tables<double>(pl) : upload_tables<float>(pl);
  if (rc) {
    delete pl;
    return 1;
  }
  c->plans[key] = pl;
  *out = pl;
  return 0;
}
// Frees the scratch planes the multi-pass path keeps for (plan, stream).  The map is keyed by the stream handle, so a
// stream must be released before it is destroyed (a recycled handle would inherit a stale entry).
static int release_stream_work(pdsp_plan* pl, cudaStream_t st) {
  if (!pl || !pl->big) return 0;
  std::lock_guard<std::recursive_mutex> lk(pl->ctx->plan_mu);
  auto it = pl->big->work.find(st);
  if (it == pl->big->work.end()) return 0;
  CU(cudaStreamSynchronize(st));
  BigPlan::Work& w = it->second;
  cudaFree(w.re);
  cudaFree(w.im);
  cudaFree(w.sync);
  cudaFree(w.spec_x);
  cudaFree(w.spec_re);
  cudaFree(w.spec_im);
  pl->big->work.erase(it);
  return 0;
}
PDSP_EXPORT int pdsp_plan_release_stream(pdsp_plan* pl, void* stream) {
  if (!pl) return fail("null argument");
  if (set_device(pl->ctx)) return 1;
  return release_stream_work(pl, stream ? static_cast<cudaStream_t>(stream) : pl->ctx->stream);
}
PDSP_EXPORT int32_t pdsp_plan_size(const pdsp_plan* p) { return p ? p->n : 0; }
PDSP_EXPORT int pdsp_plan_precision(const pdsp_plan* p) { return p ? p->precision : -1; }

static int check_desc(const pdsp_plan* pl, const pdsp_spectrum_desc* d) {
  if (!pl || !d) return fail("null argument");
  if (d->batch < 0) return fail("negative batch");
  if (d->frame_len < 0) return fail("negative frame length");
  if (d->hop < 0) return fail("negative hop");
  if (!(d->sample_rate > 0)) return fail("Sample rate must be positive, got %g", d->sample_rate);
  if (d->window < 0 || d->window > 3) return fail("Unsupported window type: %d", d->window);
  if (d->sides != PDSP_SIDES_ONE && d->sides != PDSP_SIDES_TWO) return fail("unknown sides %d", d->sides);
  if (d->sample_dtype != PDSP_F32 && d->sample_dtype != PDSP_F64) return fail("unknown sample dtype %d", d->sample_dtype);
  if (d->fft_shift != 0 && d->fft_shift != 1) return fail("fft_shift must be 0 or 1, got %d", d->fft_shift);
  if (d->fft_shift && d->sides != PDSP_SIDES_TWO) return fail("fft_shift applies to two-sided spectra only");
  return 0;
}

PDSP_EXPORT int pdsp_spectrum_dev(pdsp_plan* pl, const pdsp_spectrum_desc* d, const void* d_samples, void* d_amp,
                                  void* d_phase, void* d_peaks, void* stream) {
  if (check_desc(pl, d)) return 1;
  if (set_device(pl->ctx)) return 1;
  if (d->batch == 0) return 0;
  if (!d_samples && d->frame_len > 0) return fail("null samples");
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : pl->ctx->stream;
  return launch_spectrum(pl, d, d_samples, d->batch, d_amp, d_phase, d_peaks, nullptr, nullptr, 0, st);
}

PDSP_EXPORT int pdsp_spectrum_dev_gather(pdsp_plan* pl, const pdsp_spectrum_desc* d, const void* d_samples, void* d_amp,
                                         void* d_phase, void* d_peaks, void* const* peer_peaks, int n_peers,
                                         int64_t record_offset, void* stream) {
  if (check_desc(pl, d)) return 1;
  if (!d_peaks) return fail("pdsp_spectrum_dev_gather needs the local peaks workspace");
  if (n_peers < 0 || n_peers > 8 || (n_peers > 0 && !peer_peaks)) return fail("n_peers must be 0..8 with a pointer table");
  if (record_offset < 0) return fail("negative record offset");
  if (pl->n == 1) return fail("pdsp_spectrum_dev_gather needs an FFT size of at least 2");
  if (set_device(pl->ctx)) return 1;
  if (d->batch == 0) return 0;
  if (!d_samples && d->frame_len > 0) return fail("null samples");
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : pl->ctx->stream;
  PeerSpec ps{peer_peaks, n_peers, record_offset};
  return launch_spectrum(pl, d, d_samples, d->batch, d_amp, d_phase, d_peaks, nullptr, nullptr, 0, st, &ps);
}

PDSP_EXPORT int pdsp_ipc_export(pdsp_ctx* c, void* d_ptr, unsigned char handle[64]) {
  if (!c || !d_ptr || !handle) return fail("null argument");
  if (set_device(c)) return 1;
#ifdef PDSP_EMU
  return fail("CUDA IPC is not available in the emulated test build");
#else
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, d_ptr));
  memcpy(handle, &h, 64);
  return 0;
#endif
}
PDSP_EXPORT int pdsp_ipc_open(pdsp_ctx* c, const unsigned char handle[64], void** d_ptr) {
  if (!c || !d_ptr || !handle) return fail("null argument");
  if (set_device(c)) return 1;
#ifdef PDSP_EMU
  return fail("CUDA IPC is not available in the emulated test build");
#else
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
#endif
}
PDSP_EXPORT int pdsp_ipc_close(pdsp_ctx* c, void* d_ptr) {
  if (!c || !d_ptr) return fail("null argument");
  if (set_device(c)) return 1;
#ifdef PDSP_EMU
  return fail("CUDA IPC is not available in the emulated test build");
#else
  CU(cudaIpcCloseMemHandle(d_ptr));
  return 0;
#endif
}

PDSP_EXPORT int pdsp_fft_forward_real_dev(pdsp_plan* pl, const void* d_in, int in_dtype, int64_t batch, void* d_ore,
                                          void* d_oim, int full, void* stream) {
  if (!pl || !d_in || !d_ore || !d_oim) return fail("null argument");
  if (set_device(pl->ctx)) return 1;
  if (batch <= 0) return batch == 0 ? 0 : fail("negative batch");
  if (pl->log2n - 1 > kMaxLog2M) {
    if (in_dtype != pl->precision) return fail("large real transforms need input in the plan's precision");
    if (!full) return fail("large real transforms write all N bins");
    cudaStream_t st2 = stream ? static_cast<cudaStream_t>(stream) : pl->ctx->stream;
    return launch_c2c(pl, d_in, nullptr, batch, d_ore, d_oim, 0, st2);
  }
  pdsp_spectrum_desc d;
  memset(&d, 0, sizeof d);
  d.sample_dtype = in_dtype;
  d.frame_len = pl->n;
  d.hop = pl->n;
  d.batch = batch;
  d.window = PDSP_WIN_RECT;
  d.sides = PDSP_SIDES_ONE;
  d.sample_rate = 1.0;
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : pl->ctx->stream;
  return launch_spectrum(pl, &d, d_in, batch, nullptr, nullptr, nullptr, d_ore, d_oim, full ? 1 : 0, st);
}

PDSP_EXPORT int pdsp_fft_complex_dev(pdsp_plan* pl, const void* d_re, const void* d_im, int64_t batch, void* d_ore,
                                     void* d_oim, int inverse, void* stream) {
  if (!pl || !d_re || !d_ore || !d_oim) return fail("null argument");
  if (set_device(pl->ctx)) return 1;
  if (batch <= 0) return batch == 0 ? 0 : fail("negative batch");
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : pl->ctx->stream;
  return launch_c2c(pl, d_re, d_im, batch, d_ore, d_oim, inverse, st);
}

// ---- single-call fast lane
static int fast_init(pdsp_ctx* c) {
  FastLane& f = c->fast;
  CU(cudaStreamCreateWithFlags(&f.stream, cudaStreamNonBlocking));
  void* h = nullptr;
  CU(cudaHostAlloc(&h, kFastBytes + 256, cudaHostAllocMapped));
  f.h = static_cast<unsigned char*>(h);
  void* d = nullptr;
  CU(cudaHostGetDevicePointer(&d, h, 0));
  f.d = static_cast<unsigned char*>(d);
  memset(f.h + kFastBytes, 0, 256);
  void* cnt = nullptr;
  CU(cudaMalloc(&cnt, 256));
  CU(cudaMemset(cnt, 0, 256));
  f.d_count = static_cast<unsigned*>(cnt);
  return 0;
}
static Doorbell fast_door(pdsp_ctx* c) {
  FastLane& f = c->fast;
  if (!c->tune.doorbell) return Doorbell{nullptr, nullptr, 0};
  ++f.seq;
  if (f.seq == 0) f.seq = 1;
  return Doorbell{reinterpret_cast<unsigned*>(f.d + kFastBytes), f.d_count, f.seq};
}
// Waits for the launch that carries doorbell `seq` (door_used), or for the lane's stream otherwise.
static int fast_wait(pdsp_ctx* c, bool door_used) {
  FastLane& f = c->fast;
  if (!door_used || !c->tune.doorbell) {
    CU(cudaStreamSynchronize(f.stream));
    return 0;
  }
  volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(f.h + kFastBytes);
  for (unsigned long long spins = 1;; ++spins) {
    if (*flag == f.seq) break;
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
    if ((spins & 0xffffull) == 0) {
      // a faulted kernel never rings: let the runtime report it instead of spinning for ever
      const cudaError_t q = cudaStreamQuery(f.stream);
      if (q != cudaSuccess && q != cudaErrorNotReady) return fail("fast lane: %s", cudaGetErrorString(q));
      if (q == cudaSuccess && *flag != f.seq) {
        CU(cudaStreamSynchronize(f.stream));  // finished without ringing (should not happen): results are complete anyway
        break;
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return 0;
}

// An empty launch through the fast lane: what a one-frame call costs before any transform work (launch latency +
// doorbell over PCIe).  bench.py's c1 workload reports it next to the call latencies.
PDSP_GLOBAL void k_ping(const Doorbell door) { ring_doorbell(door); }
PDSP_EXPORT int pdsp_ctx_ping(pdsp_ctx* c) {
  if (!c) return fail("null context");
  if (set_device(c)) return 1;
  std::lock_guard<std::mutex> lk(c->mu);
  const Doorbell door = fast_door(c);
  PDSP_LAUNCH(k_ping, 1, 32, 0, c->fast.stream, door);
  CU(cudaGetLastError());
  c->launches++;
  return fast_wait(c, true);
}

// ---- host-buffer spectrum: chunked, pipelined over kSlots streams
PDSP_EXPORT int pdsp_spectrum(pdsp_plan* pl, const pdsp_spectrum_desc* d, const void* samples, void* amplitude,
                              void* phase, void* peaks) {
  if (check_desc(pl, d)) return 1;
  pdsp_ctx* c = pl->ctx;
  if (set_device(c)) return 1;
  if (d->batch == 0) return 0;
  if (!samples && d->frame_len > 0) return fail("null samples");
  std::lock_guard<std::mutex> lk(c->mu);
  const int n = pl->n;
  const int bins = d->sides == PDSP_SIDES_TWO ? n : n / 2 + 1;
  const size_t es = esize(d->sample_dtype), os = esize(pl->precision);
  const size_t pk = pl->precision == PDSP_F64 ? sizeof(pdsp_peak_f64) : sizeof(pdsp_peak_f32);
  const size_t out_per_frame = (amplitude ? bins * os : 0) + (phase ? bins * os : 0) + (peaks ? pk : 0);
  const size_t in_per_frame = (size_t)(d->hop > 0 ? (d->hop < d->frame_len ? d->hop : d->frame_len) : 0) * es;
  {
    // single-call fast lane: the whole job (input span + outputs) fits the mapped buffer and runs as one fused kernel
    const size_t span = d->frame_len > 0 ? ((size_t)(d->batch - 1) * (size_t)d->hop + (size_t)d->frame_len) * es : 0;
    const size_t a_bytes = amplitude ? (size_t)d->batch * bins * os : 0, p_bytes = phase ? (size_t)d->batch * bins * os : 0;
    const size_t k_bytes = peaks ? (size_t)d->batch * pk : 0;
    const size_t a_off = align256(span), p_off = a_off + align256(a_bytes), k_off = p_off + align256(p_bytes);
    if (c->tune.fast && n > 1 && pl->log2n - 1 <= kMaxLog2M && !c->tune.n_big_factors && k_off + align256(k_bytes) <= kFastBytes) {
      FastLane& f = c->fast;
      if (span) memcpy(f.h, samples, span);
      const Doorbell door = fast_door(c);
      bool used = false;
      if (launch_spectrum(pl, d, f.d, d->batch, amplitude ? f.d + a_off : nullptr, phase ? f.d + p_off : nullptr,
                          peaks ? f.d + k_off : nullptr, nullptr, nullptr, 0, f.stream, nullptr, &door, &used))
        return 1;
      if (fast_wait(c, used)) return 1;
      if (amplitude) memcpy(amplitude, f.h + a_off, a_bytes);
      if (phase) memcpy(phase, f.h + p_off, p_bytes);
      if (peaks) memcpy(peaks, f.h + k_off, k_bytes);
      c->fast_calls++;
      return 0;
    }
  }
  const long long chunk = pick_chunk(c, d->batch, in_per_frame + out_per_frame);
  const bool src_pinned = samples ? is_device_visible(samples) : true;
  const bool pin[3] = {amplitude ? is_device_visible(amplitude) : true, phase ? is_device_visible(phase) : true,
                       peaks ? is_device_visible(peaks) : true};
  int si = 0;
  int rc = 0;
  for (long long f0 = 0; f0 < d->batch && !rc; f0 += chunk, si = (si + 1) % kSlots) {
    Slot& s = c->slots[si];
    const long long nb = d->batch - f0 < chunk ? d->batch - f0 : chunk;
    if ((rc = slot_wait(s))) break;
    // input span of this chunk: frames f0 .. f0+nb-1
    const size_t span = d->frame_len > 0 ? ((size_t)(nb - 1) * (size_t)d->hop + (size_t)d->frame_len) * es : 0;
    const char* src = static_cast<const char*>(samples) + (size_t)f0 * (size_t)d->hop * es;
    if ((rc = slot_in(s, src, span, src_pinned))) break;
    const size_t a_bytes = amplitude ? (size_t)nb * bins * os : 0;
    const size_t p_bytes = phase ? (size_t)nb * bins * os : 0;
    const size_t k_bytes = peaks ? (size_t)nb * pk : 0;
    const size_t a_off = 0, p_off = align256(a_bytes), k_off = p_off + align256(p_bytes);
    if ((rc = ensure(&s.d_out, &s.d_out_cap, k_off + align256(k_bytes) + 256, false))) break;
    char* dout = static_cast<char*>(s.d_out);
    pdsp_spectrum_desc dd = *d;
    dd.batch = nb;
    if ((rc = launch_spectrum(pl, &dd, s.d_in, nb, amplitude ? dout + a_off : nullptr, phase ? dout + p_off : nullptr,
                              peaks ? dout + k_off : nullptr, nullptr, nullptr, 0, s.stream)))
      break;
    HostIO::Out outs[3] = {
        {amplitude ? static_cast<char*>(amplitude) + (size_t)f0 * bins * os : nullptr, a_bytes, a_off},
        {phase ? static_cast<char*>(phase) + (size_t)f0 * bins * os : nullptr, p_bytes, p_off},
        {peaks ? static_cast<char*>(peaks) + (size_t)f0 * pk : nullptr, k_bytes, k_off}};
    rc = slot_out(s, outs, 3, pin);
  }
  if (drain(c)) rc = 1;
  return rc;
}

// ---- host-buffer transforms (fp64 plans: the reference's ComplexArray is Float64Array planes)
static int host_transform(pdsp_plan* pl, const void* in_re, const void* in_im, int in_dtype, int64_t batch,
                          double* out_re, double* out_im, int mode /*0 real fwd, 1 complex fwd, 2 inverse*/) {
  if (!pl || !in_re || !out_re || !out_im) return fail("null argument");
  if (mode != 0 && !in_im) return fail("null argument");
  if (pl->precision != PDSP_F64) return fail("host transform entry points need a PDSP_F64 plan (Float64Array planes)");
  if (mode == 0 && pl->log2n - 1 > kMaxLog2M && in_dtype != PDSP_F64)
    return fail("real transforms above %d points take Float64Array input", 2 << kMaxLog2M);
  if (batch < 0) return fail("negative batch");
  pdsp_ctx* c = pl->ctx;
  if (set_device(c)) return 1;
  if (batch == 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const size_t n = (size_t)pl->n;
  const size_t es = mode == 0 ? esize(in_dtype) : 8;
  const size_t in_per_frame = n * es * (mode == 0 ? 1 : 2);
  {
    // single-call fast lane (the reference's primary call shape: one frame in, one ComplexArray out)
    const size_t plane = (size_t)batch * n * es, oplane = (size_t)batch * n * 8;
    const size_t im_off = align256(plane), or_off = im_off + (mode != 0 ? align256(plane) : 0), oi_off = or_off + align256(oplane);
    const bool single_cta_kernel = mode == 0 ? (pl->log2n - 1 <= kMaxLog2M && n > 1) : pl->log2n <= kMaxLog2M;
    if (c->tune.fast && single_cta_kernel && !c->tune.n_big_factors && oi_off + align256(oplane) <= kFastBytes) {
      FastLane& f = c->fast;
      memcpy(f.h, in_re, plane);
      if (mode != 0) memcpy(f.h + im_off, in_im, plane);
      const Doorbell door = fast_door(c);
      bool used = false;
      int rc;
      if (mode == 0) {
        pdsp_spectrum_desc dd;
        memset(&dd, 0, sizeof dd);
        dd.sample_dtype = in_dtype;
        dd.frame_len = pl->n;
        dd.hop = pl->n;
        dd.batch = batch;
        dd.window = PDSP_WIN_RECT;
        dd.sample_rate = 1.0;
        rc = launch_spectrum(pl, &dd, f.d, batch, nullptr, nullptr, nullptr, f.d + or_off, f.d + oi_off, 1, f.stream, nullptr, &door, &used);
      } else {
        rc = launch_c2c(pl, f.d, f.d + im_off, batch, f.d + or_off, f.d + oi_off, mode == 2, f.stream, &door, &used);
      }
      if (rc) return 1;
      if (fast_wait(c, used)) return 1;
      memcpy(out_re, f.h + or_off, oplane);
      memcpy(out_im, f.h + oi_off, oplane);
      c->fast_calls++;
      return 0;
    }
  }
  const long long chunk = pick_chunk(c, batch, in_per_frame + 16 * n);
  const bool pin_re = is_device_visible(in_re), pin_im = in_im ? is_device_visible(in_im) : true;
  const bool pin[2] = {is_device_visible(out_re), is_device_visible(out_im)};
  int si = 0, rc = 0;
  for (long long f0 = 0; f0 < batch && !rc; f0 += chunk, si = (si + 1) % kSlots) {
    Slot& s = c->slots[si];
    const long long nb = batch - f0 < chunk ? batch - f0 : chunk;
    if ((rc = slot_wait(s))) break;
    const size_t plane = (size_t)nb * n * es;
    const size_t plane_al = align256(plane);
    // d_in holds [re plane | im plane]; pageable sources are staged in h_in the same way
    if ((rc = ensure(&s.d_in, &s.d_in_cap, 2 * plane_al, false))) break;
    if ((!pin_re || !pin_im) && (rc = ensure(&s.h_in, &s.h_in_cap, 2 * plane_al, true))) break;
    if ((rc = slot_in(s, static_cast<const char*>(in_re) + (size_t)f0 * n * es, plane, pin_re, plane_al))) break;
    if (mode != 0) {
      const char* im_src = static_cast<const char*>(in_im) + (size_t)f0 * n * 8;
      char* d_im = static_cast<char*>(s.d_in) + plane_al;
      const void* dma_src = im_src;
      if (!pin_im) {
        char* h = static_cast<char*>(s.h_in) + plane_al;
        memcpy(h, im_src, plane);
        dma_src = h;
      }
      const cudaError_t ce = cudaMemcpyAsync(d_im, dma_src, plane, cudaMemcpyHostToDevice, s.stream);
      if (ce != cudaSuccess) {  // leave through drain(): slots of earlier chunks still point at caller buffers
        rc = fail("cudaMemcpyAsync (imaginary plane): %s", cudaGetErrorString(ce));
        break;
      }
    }
    const size_t oplane = (size_t)nb * n * 8, oplane_al = align256(oplane);
    if ((rc = ensure(&s.d_out, &s.d_out_cap, 2 * oplane_al, false))) break;
    char* dout = static_cast<char*>(s.d_out);
    if (mode == 0 && pl->log2n - 1 > kMaxLog2M) {
      // real frame longer than the in-CTA limit: multi-pass complex transform of (x, 0)
      rc = launch_c2c(pl, s.d_in, nullptr, nb, dout, dout + oplane_al, 0, s.stream);
    } else if (mode == 0) {
      pdsp_spectrum_desc dd;
      memset(&dd, 0, sizeof dd);
      dd.sample_dtype = in_dtype;
      dd.frame_len = pl->n;
      dd.hop = pl->n;
      dd.batch = nb;
      dd.window = PDSP_WIN_RECT;
      dd.sample_rate = 1.0;
      rc = launch_spectrum(pl, &dd, s.d_in, nb, nullptr, nullptr, nullptr, dout, dout + oplane_al, 1, s.stream);
    } else {
      rc = launch_c2c(pl, s.d_in, static_cast<char*>(s.d_in) + plane_al, nb, dout, dout + oplane_al, mode == 2, s.stream);
    }
    if (rc) break;
    HostIO::Out outs[2] = {{reinterpret_cast<char*>(out_re) + (size_t)f0 * n * 8, oplane, 0},
                           {reinterpret_cast<char*>(out_im) + (size_t)f0 * n * 8, oplane, oplane_al}};
    rc = slot_out(s, outs, 2, pin);
  }
  if (drain(c)) rc = 1;
  return rc;
}

PDSP_EXPORT int pdsp_fft_forward_real(pdsp_plan* pl, const void* in, int in_dtype, int64_t batch, double* out_re,
                                      double* out_im) {
  if (in_dtype != PDSP_F32 && in_dtype != PDSP_F64) return fail("unknown input dtype %d", in_dtype);
  return host_transform(pl, in, nullptr, in_dtype, batch, out_re, out_im, 0);
}
PDSP_EXPORT int pdsp_fft_forward_complex(pdsp_plan* pl, const double* in_re, const double* in_im, int64_t batch,
                                         double* out_re, double* out_im) {
  return host_transform(pl, in_re, in_im, PDSP_F64, batch, out_re, out_im, 1);
}
PDSP_EXPORT int pdsp_fft_inverse(pdsp_plan* pl, const double* in_re, const double* in_im, int64_t batch, double* out_re,
                                 double* out_im) {
  return host_transform(pl, in_re, in_im, PDSP_F64, batch, out_re, out_im, 2);
}

// op: 0 magnitude(re, im), 1 phase(re, im), 2 applyWindow(a = input, b = window), 3 fftShift(a)
static int host_elementwise(pdsp_ctx* c, const double* re, const double* im, int64_t n, double* out, int op) {
  const bool mag = op == 0;
  if (!c || !re || (!im && op != 3) || !out) return fail("null argument");
  if (n < 0) return fail("negative length");
  if (set_device(c)) return 1;
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const size_t bytes = (size_t)n * 8, al = align256(bytes);
  if (c->tune.fast && 3 * al <= kFastBytes) {
    // small arrays (the reference calls these on one frame's bins): mapped buffer, one launch, no copy-engine hops
    FastLane& f = c->fast;
    memcpy(f.h, re, bytes);
    if (im) memcpy(f.h + al, im, bytes);
    long long fb = (n + 255) / 256;
    if (fb > (long long)c->sm_count * 8) fb = (long long)c->sm_count * 8;
    const double* f_re = reinterpret_cast<const double*>(f.d);
    const double* f_im = reinterpret_cast<const double*>(f.d + al);
    double* f_o = reinterpret_cast<double*>(f.d + 2 * al);
    if (mag)
      PDSP_LAUNCH(k_magnitude, (int)fb, 256, 0, f.stream, f_re, f_im, (long long)n, f_o);
    else if (op == 1)
      PDSP_LAUNCH(k_phase, (int)fb, 256, 0, f.stream, f_re, f_im, (long long)n, f_o);
    else if (op == 2)
      PDSP_LAUNCH(k_apply_window, (int)fb, 256, 0, f.stream, f_re, f_im, (long long)n, f_o);
    else
      PDSP_LAUNCH(k_fft_shift, (int)fb, 256, 0, f.stream, f_re, (long long)n, f_o);
    CU(cudaGetLastError());
    c->launches++;
    CU(cudaStreamSynchronize(f.stream));
    memcpy(out, f.h + 2 * al, bytes);
    c->fast_calls++;
    return 0;
  }
  Slot& s = c->slots[0];
  if (slot_wait(s)) return 1;
  if (ensure(&s.d_in, &s.d_in_cap, 2 * al, false)) return 1;
  if (ensure(&s.d_out, &s.d_out_cap, al, false)) return 1;
  char* din = static_cast<char*>(s.d_in);
  CU(cudaMemcpyAsync(din, re, bytes, cudaMemcpyHostToDevice, s.stream));
  if (im) CU(cudaMemcpyAsync(din + al, im, bytes, cudaMemcpyHostToDevice, s.stream));
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)c->sm_count * 8) blocks = (long long)c->sm_count * 8;
  const double* d_re = reinterpret_cast<const double*>(din);
  const double* d_im = reinterpret_cast<const double*>(din + al);
  double* d_o = static_cast<double*>(s.d_out);
  if (mag)
    PDSP_LAUNCH(k_magnitude, (int)blocks, 256, 0, s.stream, d_re, d_im, (long long)n, d_o);
  else if (op == 1)
    PDSP_LAUNCH(k_phase, (int)blocks, 256, 0, s.stream, d_re, d_im, (long long)n, d_o);
  else if (op == 2)
    PDSP_LAUNCH(k_apply_window, (int)blocks, 256, 0, s.stream, d_re, d_im, (long long)n, d_o);
  else
    PDSP_LAUNCH(k_fft_shift, (int)blocks, 256, 0, s.stream, d_re, (long long)n, d_o);
  CU(cudaGetLastError());
  c->launches++;
  CU(cudaMemcpyAsync(out, s.d_out, bytes, cudaMemcpyDeviceToHost, s.stream));
  CU(cudaStreamSynchronize(s.stream));
  return 0;
}
PDSP_EXPORT int pdsp_magnitude(pdsp_ctx* c, const double* re, const double* im, int64_t n, double* out) {
  return host_elementwise(c, re, im, n, out, 0);
}
PDSP_EXPORT int pdsp_phase(pdsp_ctx* c, const double* re, const double* im, int64_t n, double* out) {
  return host_elementwise(c, re, im, n, out, 1);
}
PDSP_EXPORT int pdsp_apply_window(pdsp_ctx* c, const double* input, const double* window, int64_t n, double* out) {
  return host_elementwise(c, input, window, n, out, 2);
}
PDSP_EXPORT int pdsp_fft_shift(pdsp_ctx* c, const double* input, int64_t n, double* out) {
  return host_elementwise(c, input, nullptr, n, out, 3);
}

PDSP_EXPORT int pdsp_complex_mul_dev(pdsp_ctx* c, int precision, const void* a_re, const void* a_im, const void* b_re,
                                     const void* b_im, int conj_b, double scale, int64_t n, void* out_re, void* out_im,
                                     void* stream) {
  if (!c || !a_re || !a_im || !b_re || !b_im || !out_re || !out_im) return fail("null argument");
  if (precision != PDSP_F32 && precision != PDSP_F64) return fail("unknown precision %d", precision);
  if (n < 0) return fail("negative length");
  if (set_device(c)) return 1;
  if (n == 0) return 0;
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)c->sm_count * 16) blocks = (long long)c->sm_count * 16;
  if (precision == PDSP_F64)
    PDSP_LAUNCH(k_complex_mul<double>, (int)blocks, 256, 0, st, (const double*)a_re, (const double*)a_im,
                (const double*)b_re, (const double*)b_im, conj_b, scale, (long long)n, (double*)out_re, (double*)out_im);
  else
    PDSP_LAUNCH(k_complex_mul<float>, (int)blocks, 256, 0, st, (const float*)a_re, (const float*)a_im, (const float*)b_re,
                (const float*)b_im, conj_b, (float)scale, (long long)n, (float*)out_re, (float*)out_im);
  CU(cudaGetLastError());
  c->launches++;
  return 0;
}

PDSP_EXPORT int pdsp_dev_alloc(pdsp_ctx* c, size_t bytes, void** p) {
  if (!c || !p) return fail("null argument");
  if (set_device(c)) return 1;
  CU(cudaMalloc(p, bytes ? bytes : 1));
  return 0;
}
PDSP_EXPORT int pdsp_dev_free(pdsp_ctx* c, void* p) {
  if (!c) return fail("null argument");
  if (set_device(c)) return 1;
  CU(cudaFree(p));
  return 0;
}
PDSP_EXPORT int pdsp_host_alloc(pdsp_ctx* c, size_t bytes, void** p) {
  if (!c || !p) return fail("null argument");
  if (set_device(c)) return 1;
  CU(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}
PDSP_EXPORT int pdsp_host_free(pdsp_ctx* c, void* p) {
  if (!c) return fail("null argument");
  if (set_device(c)) return 1;
  CU(cudaFreeHost(p));
  return 0;
}
PDSP_EXPORT int pdsp_memcpy_h2d(pdsp_ctx* c, void* d, const void* h, size_t bytes, void* stream) {
  if (!c) return fail("null argument");
  if (set_device(c)) return 1;
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  CU(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st));
  return 0;
}
PDSP_EXPORT int pdsp_memcpy_d2h(pdsp_ctx* c, void* h, const void* d, size_t bytes, void* stream) {
  if (!c) return fail("null argument");
  if (set_device(c)) return 1;
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
  CU(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st));
  return 0;
}

// ------------------------------------------------------------------------------ host copy pool
// Row-wise memcpy split over a few persistent threads: the ingestion ring moves every frame through pinned memory
// (4-8 KB in, 2-8 KB out per frame), and one thread's ~10 GB/s was the ring's bound (profiles/r1: 1.5e6 frames/s).
// Workers spin briefly after a job (a streaming source pushes back to back) and then sleep on a condition variable.
struct CopyPool {
  struct Job {
    char* dst = nullptr;
    const char* src = nullptr;
    size_t rows = 0, row_bytes = 0, dst_stride = 0, src_stride = 0;
  };
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv;
  Job job;
  std::atomic<unsigned> generation{0};
  std::atomic<int> remaining{0};
  std::atomic<bool> stop{false};
  int parts = 1;

  static void copy_part(const Job& j, int part, int parts) {
    const size_t r0 = j.rows * (size_t)part / (size_t)parts, r1 = j.rows * (size_t)(part + 1) / (size_t)parts;
    if (j.dst_stride == j.row_bytes && j.src_stride == j.row_bytes) {
      memcpy(j.dst + r0 * j.row_bytes, j.src + r0 * j.row_bytes, (r1 - r0) * j.row_bytes);
    } else {
      for (size_t r = r0; r < r1; ++r) memcpy(j.dst + r * j.dst_stride, j.src + r * j.src_stride, j.row_bytes);
    }
  }
  void worker(int idx) {
    unsigned seen = 0;
    for (;;) {
      // wait for a new generation: spin for a while, then block
      int spins = 0;
      while (generation.load(std::memory_order_acquire) == seen && !stop.load()) {
        if (++spins < 20000) {
#if defined(__x86_64__) || defined(__i386__)
          __builtin_ia32_pause();
#endif
          continue;
        }
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return generation.load() != seen || stop.load(); });
      }
      if (stop.load()) return;
      seen = generation.load(std::memory_order_acquire);
      copy_part(job, idx + 1, parts);
      remaining.fetch_sub(1, std::memory_order_acq_rel);
    }
  }
  void start(int n_workers) {
    parts = n_workers + 1;
    for (int i = 0; i < n_workers; ++i) workers.emplace_back([this, i] { worker(i); });
  }
  // rows x row_bytes from src (row stride src_stride) to dst (row stride dst_stride); the caller copies part 0
  void run(void* dst, const void* src, size_t rows, size_t row_bytes, size_t dst_stride, size_t src_stride) {
    Job j{static_cast<char*>(dst), static_cast<const char*>(src), rows, row_bytes, dst_stride, src_stride};
    if (workers.empty() || rows * row_bytes < (256u << 10) || rows < (size_t)parts) {
      copy_part(j, 0, 1);
      return;
    }
    job = j;
    remaining.store((int)workers.size(), std::memory_order_release);
    {
      std::lock_guard<std::mutex> lk(mu);
      generation.fetch_add(1, std::memory_order_acq_rel);
    }
    cv.notify_all();
    copy_part(j, 0, parts);
    while (remaining.load(std::memory_order_acquire) != 0) {
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
    }
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop.store(true);
    }
    cv.notify_all();
    for (auto& t : workers) t.join();
  }
};

// ------------------------------------------------------------------------------ ingestion ring
// spectrumStream (src/effect/index.ts:190-194) maps a stream of frames 1:1, in order.  The ring turns that into
// batched launches without the caller assembling batches: frames are copied into a pinned chunk as they arrive;
// a full chunk is sent on its own stream (H2D -> fused kernel -> D2H into pinned memory) while the next chunk
// fills; results are handed back in arrival order.
struct IngestChunk {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  void* h_in = nullptr;   // pinned, cap frames
  void* d_in = nullptr;
  void* d_out = nullptr;  // [amplitude rows | phase rows | peak records], 256-byte aligned sections
  void* h_out = nullptr;  // pinned mirror of d_out
  long long frames = 0;   // frames copied in so far
  long long popped = 0;   // frames already handed back
  bool in_flight = false;
  bool direct = false;    // frames reached d_in by DMA from caller-pinned memory (pdsp_ingest_push_pinned)
};
struct pdsp_ingest {
  pdsp_plan* plan = nullptr;
  pdsp_spectrum_desc desc;
  long long cap = 0;
  int want_amp = 0, want_phase = 0, want_peaks = 0;
  size_t in_frame = 0, row = 0, pk = 0;        // bytes per frame: samples, one output row, one peak record
  size_t a_off = 0, p_off = 0, k_off = 0, out_bytes = 0;
  std::vector<IngestChunk> chunks;
  int fill = 0;  // chunk receiving frames
  int head = 0;  // oldest chunk with results not yet handed back
  std::mutex mu;
  CopyPool pool;  // host copies of push / pop (PDSP_COPY_THREADS helper threads, default 3)
};

static int ingest_submit(pdsp_ingest* g, IngestChunk& ch) {
  if (ch.frames == 0 || ch.in_flight) return 0;
  if (!ch.direct) CU(cudaMemcpyAsync(ch.d_in, ch.h_in, (size_t)ch.frames * g->in_frame, cudaMemcpyHostToDevice, ch.stream));
  pdsp_spectrum_desc d = g->desc;
  d.batch = ch.frames;
  char* dout = static_cast<char*>(ch.d_out);
  if (launch_spectrum(g->plan, &d, ch.d_in, ch.frames, g->want_amp ? dout + g->a_off : nullptr,
                      g->want_phase ? dout + g->p_off : nullptr, g->want_peaks ? dout + g->k_off : nullptr, nullptr, nullptr, 0,
                      ch.stream))
    return 1;
  char* hout = static_cast<char*>(ch.h_out);
  if (g->want_amp)
    CU(cudaMemcpyAsync(hout + g->a_off, dout + g->a_off, (size_t)ch.frames * g->row, cudaMemcpyDeviceToHost, ch.stream));
  if (g->want_phase)
    CU(cudaMemcpyAsync(hout + g->p_off, dout + g->p_off, (size_t)ch.frames * g->row, cudaMemcpyDeviceToHost, ch.stream));
  if (g->want_peaks)
    CU(cudaMemcpyAsync(hout + g->k_off, dout + g->k_off, (size_t)ch.frames * g->pk, cudaMemcpyDeviceToHost, ch.stream));
  CU(cudaEventRecord(ch.done, ch.stream));
  ch.in_flight = true;
  ch.popped = 0;
  return 0;
}

PDSP_EXPORT int pdsp_ingest_open(pdsp_plan* pl, const pdsp_spectrum_desc* d, int want_amplitude, int want_phase,
                                 int want_peaks, int64_t frames_per_chunk, int depth, pdsp_ingest** out) {
  if (!out) return fail("null argument");
  *out = nullptr;
  if (check_desc(pl, d)) return 1;
  if (frames_per_chunk <= 0) return fail("frames_per_chunk must be positive");
  if (depth < 2 || depth > 16) return fail("ring depth must be 2..16, got %d", depth);
  if (d->frame_len <= 0) return fail("the ingestion ring needs a positive frame length");
  if (!want_amplitude && !want_phase && !want_peaks) return fail("no output requested");
  if (set_device(pl->ctx)) return 1;
  pdsp_ingest* g = new pdsp_ingest();
  g->plan = pl;
  g->desc = *d;
  g->desc.hop = d->frame_len;  // frames are packed back to back inside a chunk
  g->cap = frames_per_chunk;
  g->want_amp = want_amplitude != 0, g->want_phase = want_phase != 0, g->want_peaks = want_peaks != 0;
  const int n = pl->n;
  const size_t bins = d->sides == PDSP_SIDES_TWO ? (size_t)n : (size_t)n / 2 + 1;
  g->in_frame = (size_t)d->frame_len * esize(d->sample_dtype);
  g->row = bins * esize(pl->precision);
  g->pk = pl->precision == PDSP_F64 ? sizeof(pdsp_peak_f64) : sizeof(pdsp_peak_f32);
  g->a_off = 0;
  g->p_off = g->a_off + align256(g->want_amp ? (size_t)g->cap * g->row : 0);
  g->k_off = g->p_off + align256(g->want_phase ? (size_t)g->cap * g->row : 0);
  g->out_bytes = g->k_off + align256(g->want_peaks ? (size_t)g->cap * g->pk : 0) + 256;
  g->chunks.resize((size_t)depth);
  int rc = 0;
  for (auto& ch : g->chunks) {
    cudaError_t e = cudaStreamCreateWithFlags(&ch.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ch.done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc(&ch.h_in, (size_t)g->cap * g->in_frame, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaMalloc(&ch.d_in, (size_t)g->cap * g->in_frame + 256);
    if (e == cudaSuccess) e = cudaMalloc(&ch.d_out, g->out_bytes);
    if (e == cudaSuccess) e = cudaHostAlloc(&ch.h_out, g->out_bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
      rc = fail("ingestion ring allocation: %s", cudaGetErrorString(e));
      break;
    }
  }
  if (rc) {
    pdsp_ingest_close(g);
    return 1;
  }
  {
    int helpers = pl->ctx->tune.copy_threads;
    const int hw = (int)std::thread::hardware_concurrency();
    if (helpers > hw - 1) helpers = hw - 1;
    if (helpers > 15) helpers = 15;
    if (helpers > 0) g->pool.start(helpers);
  }
  *out = g;
  return 0;
}

PDSP_EXPORT int pdsp_ingest_close(pdsp_ingest* g) {
  if (!g) return 0;
  if (g->plan && set_device(g->plan->ctx)) return 1;
  for (auto& ch : g->chunks) {
    if (ch.stream) cudaStreamSynchronize(ch.stream);
    if (ch.stream && g->plan) release_stream_work(g->plan, ch.stream);  // large-N scratch planes keyed by this stream
    cudaFreeHost(ch.h_in);
    cudaFree(ch.d_in);
    cudaFree(ch.d_out);
    cudaFreeHost(ch.h_out);
    if (ch.done) cudaEventDestroy(ch.done);
    if (ch.stream) cudaStreamDestroy(ch.stream);
  }
  delete g;
  return 0;
}

// Copies `count` frames (frame_len samples each, `stride` samples apart; 0 = frame_len) into the ring, sending
// every chunk that fills up.  *accepted is how many were taken: fewer than `count` means the ring is full -
// the next chunk still holds results the caller has not collected with pdsp_ingest_pop.
// pinned != 0 (pdsp_ingest_push_pinned): `frames` is pinned / registered host memory and is NOT copied - each run of
// frames is sent to the device by DMA straight from the caller's buffer, which must stay unchanged until those frames
// have been popped.
static int ingest_push_impl(pdsp_ingest* g, const void* frames, int64_t count, int64_t stride, int64_t* accepted, int pinned) {
  if (!g || (!frames && count > 0)) return fail("null argument");
  if (count < 0 || stride < 0) return fail("negative count or stride");
  if (accepted) *accepted = 0;
  if (set_device(g->plan->ctx)) return 1;
  std::lock_guard<std::mutex> lk(g->mu);
  const size_t es = esize(g->desc.sample_dtype);
  const size_t step = (size_t)(stride ? stride : g->desc.frame_len) * es;
  if (pinned && count > 0 && !is_device_visible(frames)) return fail("pdsp_ingest_push_pinned needs pinned (page-locked) host memory");
  const char* src = static_cast<const char*>(frames);
  int64_t done = 0;
  while (done < count) {
    IngestChunk& ch = g->chunks[(size_t)g->fill];
    if (ch.in_flight) break;  // ring full
    long long take = g->cap - ch.frames;
    if (take > count - done) take = count - done;
    if (pinned) {
      // DMA from the caller's memory into the chunk's device buffer (the chunk's host staging is bypassed)
      char* ddst = static_cast<char*>(ch.d_in) + (size_t)ch.frames * g->in_frame;
      if (step == g->in_frame) {
        CU(cudaMemcpyAsync(ddst, src, (size_t)take * g->in_frame, cudaMemcpyHostToDevice, ch.stream));
      } else {
        CU(cudaMemcpy2DAsync(ddst, g->in_frame, src, step, g->in_frame, (size_t)take, cudaMemcpyHostToDevice, ch.stream));
      }
      ch.direct = true;
    } else {
      char* dst = static_cast<char*>(ch.h_in) + (size_t)ch.frames * g->in_frame;
      g->pool.run(dst, src, (size_t)take, g->in_frame, g->in_frame, step);
      if (ch.direct) return fail("one chunk cannot mix copied and pinned pushes: flush between them");
    }
    ch.frames += take;
    done += take;
    src += (size_t)take * step;
    if (ch.frames == g->cap) {
      if (ingest_submit(g, ch)) return 1;
      g->fill = (g->fill + 1) % (int)g->chunks.size();
    }
  }
  if (accepted) *accepted = done;
  return 0;
}
PDSP_EXPORT int pdsp_ingest_push(pdsp_ingest* g, const void* frames, int64_t count, int64_t stride, int64_t* accepted) {
  return ingest_push_impl(g, frames, count, stride, accepted, 0);
}
PDSP_EXPORT int pdsp_ingest_push_pinned(pdsp_ingest* g, const void* frames, int64_t count, int64_t stride, int64_t* accepted) {
  return ingest_push_impl(g, frames, count, stride, accepted, 1);
}

// Non-blocking progress report: frames whose results can be popped without waiting (`finished`: the leading run of
// completed chunks), frames sent and still in flight, frames sitting in the partially filled chunk (not sent yet).
PDSP_EXPORT int pdsp_ingest_ready(pdsp_ingest* g, int64_t* finished, int64_t* in_flight, int64_t* pending) {
  if (!g) return fail("null argument");
  if (set_device(g->plan->ctx)) return 1;
  std::lock_guard<std::mutex> lk(g->mu);
  int64_t fin = 0, fly = 0;
  bool leading = true;
  const int nch = (int)g->chunks.size();
  for (int i = 0; i < nch; ++i) {
    IngestChunk& ch = g->chunks[(size_t)((g->head + i) % nch)];
    if (!ch.in_flight) break;
    bool complete = false;
    if (leading) {
      const cudaError_t q = cudaEventQuery(ch.done);
      if (q == cudaSuccess)
        complete = true;
      else if (q != cudaErrorNotReady)
        return fail("cudaEventQuery: %s", cudaGetErrorString(q));
    }
    if (complete) {
      fin += ch.frames - ch.popped;
    } else {
      leading = false;
      fly += ch.frames - ch.popped;
    }
  }
  const IngestChunk& fc = g->chunks[(size_t)g->fill];
  if (finished) *finished = fin;
  if (in_flight) *in_flight = fly;
  if (pending) *pending = fc.in_flight ? 0 : fc.frames;
  return 0;
}

// Sends the partially filled chunk (end of stream, or a latency bound on the caller's side).
PDSP_EXPORT int pdsp_ingest_flush(pdsp_ingest* g) {
  if (!g) return fail("null argument");
  if (set_device(g->plan->ctx)) return 1;
  std::lock_guard<std::mutex> lk(g->mu);
  IngestChunk& ch = g->chunks[(size_t)g->fill];
  if (ch.in_flight || ch.frames == 0) return 0;
  if (ingest_submit(g, ch)) return 1;
  g->fill = (g->fill + 1) % (int)g->chunks.size();
  return 0;
}

// Hands back up to max_frames finished frames in arrival order (rows appended densely to the caller's arrays,
// which may be NULL for outputs not requested at open).  Blocks on chunks already sent; frames still sitting in
// a partially filled chunk are not waited for.  *got may be 0.
PDSP_EXPORT int pdsp_ingest_pop(pdsp_ingest* g, void* amplitude, void* phase, void* peaks, int64_t max_frames, int64_t* got) {
  if (!g || !got) return fail("null argument");
  *got = 0;
  if (max_frames < 0) return fail("negative max_frames");
  if (set_device(g->plan->ctx)) return 1;
  std::lock_guard<std::mutex> lk(g->mu);
  int64_t n = 0;
  while (n < max_frames) {
    IngestChunk& ch = g->chunks[(size_t)g->head];
    if (!ch.in_flight) break;
    CU(cudaEventSynchronize(ch.done));
    long long take = ch.frames - ch.popped;
    if (take > max_frames - n) take = max_frames - n;
    const char* hout = static_cast<const char*>(ch.h_out);
    if (amplitude && g->want_amp)
      g->pool.run(static_cast<char*>(amplitude) + (size_t)n * g->row, hout + g->a_off + (size_t)ch.popped * g->row, (size_t)take,
                  g->row, g->row, g->row);
    if (phase && g->want_phase)
      g->pool.run(static_cast<char*>(phase) + (size_t)n * g->row, hout + g->p_off + (size_t)ch.popped * g->row, (size_t)take, g->row,
                  g->row, g->row);
    if (peaks && g->want_peaks)
      memcpy(static_cast<char*>(peaks) + (size_t)n * g->pk, hout + g->k_off + (size_t)ch.popped * g->pk, (size_t)take * g->pk);
    ch.popped += take;
    n += take;
    if (ch.popped == ch.frames) {
      ch.in_flight = false;
      ch.direct = false;
      ch.frames = ch.popped = 0;
      g->head = (g->head + 1) % (int)g->chunks.size();
    }
  }
  *got = n;
  return 0;
}

// ------------------------------------------------------------------------------ device groups
// Frames are independent (the reference's spectrumStream is a pure 1:1 map, src/effect/index.ts:190-194), so a box of
// GPUs is used by cutting the frame range into one contiguous block per device.  A group is a set of contexts in ONE
// process with peer access enabled between their devices:
//   * pdsp_group_spectrum      - host frames in, host results out: every device runs its block through its own staging
//                                pipeline on its own host thread (pinned to the CPUs of the device's NUMA node), results
//                                land in the caller's arrays at the block's offset - the gather happens on the way out;
//   * pdsp_group_spectrum_dev  - device-resident blocks: every device's kernel also stores each finished peak record
//                                into the gather buffers of ALL devices (the fused NVLink scatter of
//                                pdsp_spectrum_dev_gather, with plain peer pointers), and amplitude / phase rows are
//                                copied to one root device by peer DMA when the caller asks for them there.
struct pdsp_group {
  std::vector<pdsp_ctx*> ctx;
  std::vector<std::vector<int>> cpus;  // CPUs of each device's NUMA node (empty: unknown, threads are left unpinned)
};

#ifndef PDSP_EMU
#include <sched.h>
#endif
static std::vector<int> numa_cpus_of_device(int device) {
  std::vector<int> cpus;
#ifndef PDSP_EMU
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return cpus;
  for (char* q = bus; *q; ++q)
    if (*q >= 'A' && *q <= 'F') *q = (char)(*q - 'A' + 'a');
  char path[128];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* f = fopen(path, "r");
  int node = -1;
  if (f) {
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
  }
  if (node < 0) return cpus;
  snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
  f = fopen(path, "r");
  if (!f) return cpus;
  char list[4096] = {0};
  if (fgets(list, sizeof list, f)) {
    for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
      int a = 0, b = 0;
      const int got = sscanf(tok, "%d-%d", &a, &b);
      if (got == 1) b = a;
      if (got >= 1)
        for (int k = a; k <= b; ++k) cpus.push_back(k);
    }
  }
  fclose(f);
#else
  (void)device;
#endif
  return cpus;
}
static void pin_thread_to(const std::vector<int>& cpus) {
#ifndef PDSP_EMU
  if (cpus.empty()) return;
  cpu_set_t set;
  CPU_ZERO(&set);
  for (int k : cpus)
    if (k >= 0 && k < CPU_SETSIZE) CPU_SET(k, &set);
  sched_setaffinity(0, sizeof set, &set);  // best effort: an error leaves the thread where it was
#else
  (void)cpus;
#endif
}

PDSP_EXPORT int pdsp_group_create(const int* devices, int n, pdsp_group** out) {
  if (!devices || !out) return fail("null argument");
  *out = nullptr;
  if (n < 1 || n > 8) return fail("a group holds 1..8 devices, got %d", n);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < i; ++k)
      if (devices[i] == devices[k]) return fail("device %d listed twice", devices[i]);
  pdsp_group* g = new pdsp_group();
  for (int i = 0; i < n; ++i) {
    pdsp_ctx* c = nullptr;
    if (pdsp_ctx_create(devices[i], &c)) {
      for (pdsp_ctx* o : g->ctx) pdsp_ctx_destroy(o);
      delete g;
      return 1;
    }
    g->ctx.push_back(c);
    g->cpus.push_back(numa_cpus_of_device(devices[i]));
  }
  // peer access both ways (NVLink / NVSwitch): the fused record scatter and the row gather write peer memory directly
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k) {
      if (i == k) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[i], devices[k]);
      if (!can) continue;
      if (cudaSetDevice(devices[i]) != cudaSuccess) continue;
      const cudaError_t e = cudaDeviceEnablePeerAccess(devices[k], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) fail("peer access %d -> %d: %s", devices[i], devices[k], cudaGetErrorString(e));
      cudaGetLastError();
    }
  *out = g;
  return 0;
}
PDSP_EXPORT int pdsp_group_destroy(pdsp_group* g) {
  if (!g) return 0;
  int rc = 0;
  for (pdsp_ctx* c : g->ctx) rc |= pdsp_ctx_destroy(c);
  delete g;
  return rc;
}
PDSP_EXPORT int pdsp_group_size(const pdsp_group* g) { return g ? (int)g->ctx.size() : 0; }
PDSP_EXPORT pdsp_ctx* pdsp_group_ctx(pdsp_group* g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[(size_t)i] : nullptr; }

// block partition of `batch` frames over `n` devices: device i owns [i*per, min(batch, (i+1)*per)), per = ceil(batch / n)
static void group_block(long long batch, int n, int i, long long* f0, long long* nf) {
  const long long per = (batch + n - 1) / n;
  long long a = per * i, b = a + per;
  if (a > batch) a = batch;
  if (b > batch) b = batch;
  *f0 = a;
  *nf = b - a;
}

PDSP_EXPORT int pdsp_group_spectrum(pdsp_group* g, int32_t size, int precision, const pdsp_spectrum_desc* d, const void* samples,
                                    void* amplitude, void* phase, void* peaks) {
  if (!g || !d) return fail("null argument");
  const int n = (int)g->ctx.size();
  std::vector<pdsp_plan*> plans((size_t)n, nullptr);
  for (int i = 0; i < n; ++i)
    if (pdsp_plan_get(g->ctx[(size_t)i], size, precision, &plans[(size_t)i])) return 1;
  if (check_desc(plans[0], d)) return 1;
  if (d->batch == 0) return 0;
  const int bins = d->sides == PDSP_SIDES_TWO ? size : size / 2 + 1;
  const size_t es = esize(d->sample_dtype), os = esize(precision);
  const size_t pk = precision == PDSP_F64 ? sizeof(pdsp_peak_f64) : sizeof(pdsp_peak_f32);
  std::vector<int> rcs((size_t)n, 0);
  std::vector<std::string> errs((size_t)n);
  auto work = [&](int i) {
    long long f0, nf;
    group_block(d->batch, n, i, &f0, &nf);
    if (nf <= 0) return;
    pin_thread_to(g->cpus[(size_t)i]);
    pdsp_spectrum_desc dd = *d;
    dd.batch = nf;
    const char* src = samples ? static_cast<const char*>(samples) + (size_t)f0 * (size_t)d->hop * es : nullptr;
    rcs[(size_t)i] = pdsp_spectrum(plans[(size_t)i], &dd, src, amplitude ? static_cast<char*>(amplitude) + (size_t)f0 * bins * os : nullptr,
                                   phase ? static_cast<char*>(phase) + (size_t)f0 * bins * os : nullptr,
                                   peaks ? static_cast<char*>(peaks) + (size_t)f0 * pk : nullptr);
    if (rcs[(size_t)i]) errs[(size_t)i] = g_err;  // the message lives in this worker's thread-local slot
  };
  std::vector<std::thread> th;
  for (int i = 1; i < n; ++i) th.emplace_back(work, i);
  {
    // the calling thread serves device 0; its affinity is restored afterwards
#ifndef PDSP_EMU
    cpu_set_t saved;
    const bool have = sched_getaffinity(0, sizeof saved, &saved) == 0;
#endif
    work(0);
#ifndef PDSP_EMU
    if (have) sched_setaffinity(0, sizeof saved, &saved);
#endif
  }
  for (auto& t : th) t.join();
  for (int i = 0; i < n; ++i)
    if (rcs[(size_t)i]) {
      g_err = errs[(size_t)i];
      return 1;
    }
  return 0;
}

// Device-resident sharded form.  Device i holds its block's samples at d_samples[i] (frame f of the block at f*hop) and
// receives its outputs in d_amplitude[i] / d_phase[i] (block-local rows; entries or the arrays themselves may be NULL).
// d_peaks[i] is device i's GATHER buffer, desc->batch records long: every device's kernel writes its block's records into
// all of them (peer stores over NVLink fused into the kernel's finishing loop), so after pdsp_group_sync each device holds
// the peaks of ALL frames.  gather_root >= 0: amplitude / phase rows of all blocks are also copied into
// d_amplitude_all / d_phase_all (desc->batch rows, on device gather_root) by peer DMA queued behind each block's kernel.
PDSP_EXPORT int pdsp_group_spectrum_dev(pdsp_group* g, int32_t size, int precision, const pdsp_spectrum_desc* d,
                                        const void* const* d_samples, void* const* d_amplitude, void* const* d_phase,
                                        void* const* d_peaks, int gather_root, void* d_amplitude_all, void* d_phase_all) {
  if (!g || !d || !d_samples) return fail("null argument");
  const int n = (int)g->ctx.size();
  if (gather_root >= n) return fail("gather root %d out of range (%d devices)", gather_root, n);
  if (gather_root >= 0 && ((d_amplitude_all && !d_amplitude) || (d_phase_all && !d_phase)))
    return fail("a row gather needs the per-device row buffers it copies from");
  std::vector<pdsp_plan*> plans((size_t)n, nullptr);
  for (int i = 0; i < n; ++i)
    if (pdsp_plan_get(g->ctx[(size_t)i], size, precision, &plans[(size_t)i])) return 1;
  if (check_desc(plans[0], d)) return 1;
  if (d->batch == 0) return 0;
  if (d_peaks && size == 1) return fail("the fused peak gather needs an FFT size of at least 2");
  const int bins = d->sides == PDSP_SIDES_TWO ? size : size / 2 + 1;
  const size_t os = esize(precision);
  const size_t pk = precision == PDSP_F64 ? sizeof(pdsp_peak_f64) : sizeof(pdsp_peak_f32);
  void* peers[8] = {nullptr};
  if (d_peaks)
    for (int i = 0; i < n; ++i) {
      if (!d_peaks[i]) return fail("null gather buffer for device %d", i);
      peers[i] = d_peaks[i];
    }
  for (int i = 0; i < n; ++i) {
    long long f0, nf;
    group_block(d->batch, n, i, &f0, &nf);
    if (nf <= 0) continue;
    pdsp_ctx* c = g->ctx[(size_t)i];
    if (set_device(c)) return 1;
    if (!d_samples[i] && d->frame_len > 0) return fail("null samples for device %d", i);
    pdsp_spectrum_desc dd = *d;
    dd.batch = nf;
    void* amp = d_amplitude ? d_amplitude[i] : nullptr;
    void* ph = d_phase ? d_phase[i] : nullptr;
    // the block's own segment of its gather buffer doubles as the kernel's local record workspace
    void* local = d_peaks ? static_cast<char*>(d_peaks[i]) + (size_t)f0 * pk : nullptr;
    PeerSpec ps{peers, n, f0};
    // the kernel writes record f of the block to peer[g] + (offset + f): `local` is that same address on this device,
    // so the self-entry is dropped from the peer list (no duplicate store)
    void* others[8];
    int no = 0;
    for (int k = 0; k < n; ++k)
      if (k != i) others[no++] = peers[k];
    ps.ptrs = others;
    ps.n = d_peaks ? no : 0;
    if (launch_spectrum(plans[(size_t)i], &dd, d_samples[i], nf, amp, ph, local, nullptr, nullptr, 0, c->stream, d_peaks ? &ps : nullptr))
      return 1;
    if (gather_root >= 0) {
      const int root_dev = g->ctx[(size_t)gather_root]->device;
      if (d_amplitude_all && amp)
        CU(cudaMemcpyPeerAsync(static_cast<char*>(d_amplitude_all) + (size_t)f0 * bins * os, root_dev, amp, c->device,
                               (size_t)nf * bins * os, c->stream));
      if (d_phase_all && ph)
        CU(cudaMemcpyPeerAsync(static_cast<char*>(d_phase_all) + (size_t)f0 * bins * os, root_dev, ph, c->device,
                               (size_t)nf * bins * os, c->stream));
    }
  }
  return 0;
}
// waits for everything queued on the group's context streams
PDSP_EXPORT int pdsp_group_sync(pdsp_group* g) {
  if (!g) return fail("null argument");
  for (pdsp_ctx* c : g->ctx) {
    if (set_device(c)) return 1;
    CU(cudaStreamSynchronize(c->stream));
  }
  return 0;
}
