/*
 * pragma_oracle.c - CPU ORACLE for the pragma-dsp FFT/spectrum hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is a plain-C, op-for-op restatement of the
 * reference's TypeScript algorithm (eliesgalvira/pragma-dsp v0.1.0).  It is the checker
 * the CUDA path is compared against; it is never the thing shipped or measured as the
 * product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product library (pragma_dsp_b200/csrc) does not link,
 * include or call anything in oracle/.
 *
 * Parity status: PINNED.  Checked (tests/test_oracle_golden.py) against every golden
 * vector the reference's own tests hold for this path:
 *   - test/reallife/references/{pure_sine,cosine,multi_tone,chirp,special}.json (35 cases,
 *     NumPy 2.4.2) with the tolerances of test/reallife/ *.test.ts,
 *   - test/reallife/references/windows_dsp.json and the regenerated
 *     test/fixtures/pragma-dsp.v0.1.json (scripts/gen_fixtures.py, seed 1337) with the
 *     tolerances of test/fft.test.ts, test/spectrum.test.ts, test/window.test.ts,
 *   - bench/run.ts's guardrail checksums.
 * The reference itself cannot run here (no JavaScript runtime in the image).
 *
 * JS semantics mirrored: IEEE-754 binary64 everywhere, NO fused multiply-add (build with
 * -ffp-contract=off), `x ?? 0` for missing elements (callers pass dense arrays), strict `>`
 * comparisons, `(2*mag)/size` evaluation order.  Third-party arithmetic that is NOT in
 * /root/reference: the JS engine's libm (Math.cos/sin/hypot/atan2; engine unpinned - Node/V8
 * via package.json:70, Bun/JSC in bench/reallife).  It is restated with the C library's
 * cos/sin/hypot/atan2, which may differ from a given JS engine in the last ulp; the
 * reference's tests pin that boundary only to 1e-8..1e-10 absolute.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PO_API __attribute__((visibility("default")))

enum { PO_WIN_RECT = 0, PO_WIN_HANN = 1, PO_WIN_HAMMING = 2, PO_WIN_BLACKMAN = 3 };
enum { PO_SIDES_ONE = 0, PO_SIDES_TWO = 1 };
enum { PO_F32 = 0, PO_F64 = 1 };

typedef struct {
  int32_t index;
  int32_t _pad;
  double frequency;
  double amplitude;
  double phase;
} po_peak;

typedef struct {
  int size;
  int stages;
  uint32_t* bit_reverse; /* src/core/fft.ts:25-38 */
  double** tw_cos;       /* src/core/fft.ts:45-61, one table per stage */
  double** tw_sin;
} po_plan;

/* src/core/fft.ts:16  isPowerOfTwo: n > 0 && (n & (n-1)) === 0 (int32 semantics) */
PO_API int po_is_power_of_two(int32_t n) { return n > 0 && (n & (n - 1)) == 0; }

/* src/core/fft.ts:18-23  nextPowerOfTwo */
PO_API int32_t po_next_power_of_two(int32_t n) {
  if (n <= 1) return 1;
  int32_t p = 1;
  while (p < n) p <<= 1;
  return p;
}

/* src/core/fft.ts:68-75 (constructor), :25-38 (buildBitReverse), :45-61 (buildTwiddles).
 * Returns NULL for a non power-of-two size (the reference throws
 * "FFT size must be power of two, got ${size}"). */
PO_API po_plan* po_plan_create(int size) {
  if (!po_is_power_of_two(size)) return NULL;
  po_plan* p = (po_plan*)calloc(1, sizeof(po_plan));
  p->size = size;
  int bits = (int)lround(log2((double)size)); /* Math.round(Math.log2(size)) */
  p->stages = bits;
  p->bit_reverse = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)size);
  for (int i = 0; i < size; i += 1) {
    uint32_t x = (uint32_t)i, y = 0;
    for (int b = 0; b < bits; b += 1) {
      y = (y << 1) | (x & 1u);
      x >>= 1;
    }
    p->bit_reverse[i] = y;
  }
  p->tw_cos = (double**)calloc((size_t)(bits > 0 ? bits : 1), sizeof(double*));
  p->tw_sin = (double**)calloc((size_t)(bits > 0 ? bits : 1), sizeof(double*));
  for (int stage = 1; stage <= bits; stage += 1) {
    int m = 1 << stage;
    int half = m >> 1;
    double* c = (double*)malloc(sizeof(double) * (size_t)half);
    double* s = (double*)malloc(sizeof(double) * (size_t)half);
    for (int k = 0; k < half; k += 1) {
      /* const angle = (-2 * Math.PI * k) / m;  -- this operation order */
      double angle = (-2.0 * M_PI * (double)k) / (double)m;
      c[k] = cos(angle);
      s[k] = sin(angle);
    }
    p->tw_cos[stage - 1] = c;
    p->tw_sin[stage - 1] = s;
  }
  return p;
}

PO_API void po_plan_destroy(po_plan* p) {
  if (!p) return;
  for (int s = 0; s < p->stages; s += 1) {
    free(p->tw_cos[s]);
    free(p->tw_sin[s]);
  }
  free(p->tw_cos);
  free(p->tw_sin);
  free(p->bit_reverse);
  free(p);
}

PO_API int po_plan_size(const po_plan* p) { return p->size; }

/* src/core/fft.ts:89-151  Radix2Fft.transform.  in_im may be NULL (forward of a real frame,
 * :77-79).  inverse != 0 conjugates the twiddles (sinSign = -1, :122) and multiplies by the
 * precomputed reciprocal 1/size (:142-148). */
PO_API void po_fft_transform(const po_plan* p, const double* in_re, const double* in_im, double* out_re,
                             double* out_im, int inverse) {
  const int size = p->size;
  for (int i = 0; i < size; i += 1) {
    uint32_t j = p->bit_reverse[i];
    out_re[j] = in_re[i];
    out_im[j] = in_im ? in_im[i] : 0.0;
  }
  for (int stage = 0; stage < p->stages; stage += 1) {
    const int m = 1 << (stage + 1);
    const int half = m >> 1;
    const double* c = p->tw_cos[stage];
    const double* s = p->tw_sin[stage];
    const double sin_sign = inverse ? -1.0 : 1.0;
    for (int k = 0; k < size; k += m) {
      for (int j = 0; j < half; j += 1) {
        /* JS evaluates `sinSign * sin[j] * x` left to right: (sinSign*sin[j])*x */
        const double ss = sin_sign * s[j];
        const double t_re = c[j] * out_re[k + j + half] - ss * out_im[k + j + half];
        const double t_im = ss * out_re[k + j + half] + c[j] * out_im[k + j + half];
        const double u_re = out_re[k + j];
        const double u_im = out_im[k + j];
        out_re[k + j] = u_re + t_re;
        out_im[k + j] = u_im + t_im;
        out_re[k + j + half] = u_re - t_re;
        out_im[k + j + half] = u_im - t_im;
      }
    }
  }
  if (inverse) {
    const double scale = 1.0 / (double)size;
    for (int i = 0; i < size; i += 1) {
      out_re[i] = out_re[i] * scale;
      out_im[i] = out_im[i] * scale;
    }
  }
}

/* src/xform/fourier.ts:14-52  createWindow (symmetric, denominator size-1).
 * Returns 0 ok, 1 = "Window size must be positive", 2 = "Unsupported window type". */
PO_API int po_create_window(int type, int size, double* out) {
  if (size <= 0) return 1;
  if (size == 1) {
    out[0] = 1.0;
    return 0;
  }
  switch (type) {
    case PO_WIN_RECT:
      for (int i = 0; i < size; i += 1) out[i] = 1.0;
      return 0;
    case PO_WIN_HANN:
      for (int i = 0; i < size; i += 1) out[i] = 0.5 * (1.0 - cos((2.0 * M_PI * (double)i) / (double)(size - 1)));
      return 0;
    case PO_WIN_HAMMING:
      for (int i = 0; i < size; i += 1) out[i] = 0.54 - 0.46 * cos((2.0 * M_PI * (double)i) / (double)(size - 1));
      return 0;
    case PO_WIN_BLACKMAN:
      for (int i = 0; i < size; i += 1) {
        const double f = (2.0 * M_PI * (double)i) / (double)(size - 1);
        out[i] = 0.42 - 0.5 * cos(f) + 0.08 * cos(2.0 * f);
      }
      return 0;
    default:
      return 2;
  }
}

/* src/xform/fourier.ts:54-67  applyWindow */
PO_API void po_apply_window(const double* input, const double* window, int n, double* out) {
  for (int i = 0; i < n; i += 1) out[i] = input[i] * window[i];
}

/* src/xform/fourier.ts:98-109  magnitude = Math.hypot(re, im) over ALL bins */
PO_API void po_magnitude(const double* re, const double* im, int n, double* out) {
  for (int i = 0; i < n; i += 1) out[i] = hypot(re[i], im[i]);
}

/* src/xform/fourier.ts:111-120  phase = Math.atan2(im, re) over ALL bins */
PO_API void po_phase(const double* re, const double* im, int n, double* out) {
  for (int i = 0; i < n; i += 1) out[i] = atan2(im[i], re[i]);
}

/* src/xform/fourier.ts:122-134  fftShift: out[i] = in[(i + floor(n/2)) % n] */
PO_API void po_fft_shift(const double* in, int n, double* out) {
  const int mid = n / 2;
  for (int i = 0; i < n; i += 1) out[i] = in[(i + mid) % n];
}

/* src/xform/fourier.ts:147-165  binFrequencies; the quotient sampleRate/size is formed first.
 * Returns bin count, or -1 / -2 for the reference's two throws. */
PO_API int po_bin_frequencies(int size, double sample_rate, int sides, double* out) {
  if (size <= 0) return -1;
  if (!(sample_rate > 0)) return -2;
  const int bins = sides == PO_SIDES_ONE ? size / 2 + 1 : size;
  const double scale = sample_rate / (double)size;
  if (out)
    for (int i = 0; i < bins; i += 1) out[i] = (double)i * scale;
  return bins;
}

/* src/public/spectrum.ts:45-61  scaleAmplitudeOneSided */
PO_API void po_scale_amplitude_one_sided(const double* mag, int size, double* out) {
  const int bins = size / 2 + 1;
  const int nyquist = size % 2 == 0 ? size / 2 : -1;
  for (int k = 0; k < bins; k += 1) {
    const double m = mag[k];
    if (k == 0 || k == nyquist)
      out[k] = m / (double)size;
    else
      out[k] = (2.0 * m) / (double)size;
  }
}

/* src/public/spectrum.ts:63-72  scaleAmplitudeTwoSided */
PO_API void po_scale_amplitude_two_sided(const double* mag, int size, double* out) {
  for (int k = 0; k < size; k += 1) out[k] = mag[k] / (double)size;
}

/* src/public/spectrum.ts:74-105  findPeak (phase is filled by the caller, :134) */
PO_API void po_find_peak(const double* amplitude, const double* frequencies, int len, po_peak* peak) {
  int max_index = 0;
  double max_value = len > 0 ? amplitude[0] : 0.0;
  int has_non_dc = 0;
  int non_dc_index = 0;
  double non_dc_value = 0.0;
  for (int i = 1; i < len; i += 1) {
    const double v = amplitude[i];
    if (v > non_dc_value) {
      non_dc_value = v;
      non_dc_index = i;
    }
    if (v > 0) has_non_dc = 1;
    if (v > max_value) {
      max_value = v;
      max_index = i;
    }
  }
  const int index = has_non_dc ? non_dc_index : max_index;
  peak->index = index;
  peak->_pad = 0;
  peak->frequency = index < len ? frequencies[index] : 0.0;
  peak->amplitude = index < len ? amplitude[index] : 0.0;
  peak->phase = 0.0;
}

/* src/public/spectrum.ts:107-142  spectrum() for one frame, with the plan and the window
 * supplied by the caller (that is spectrumWithService, src/effect/index.ts:143-179, which the
 * reference's tests require to be bit-identical to spectrum()).
 *   samples/len : the caller's samples (already widened to double; Float32Array elements widen
 *                 exactly), buildFrame zero-pads / truncates to plan->size (:36-43)
 *   amplitude, phase_out, frequencies : bins = size/2+1 (one) or size (two); any may be NULL
 *   scratch : 5*size doubles of workspace */
/* lean != 0 (bench.py's reference arm only, so that it computes the SAME outputs as the GPU arm it is compared
 * with): Math.atan2 runs over all bins only when the phase array is requested; with a peak but no phase array it
 * runs once, at the peak bin.  lean == 0 is spectrum() as written (all phases always, src/public/spectrum.ts:121-122). */
static void po_spectrum_frame_ex(const po_plan* p, const double* window, const double* samples, int len,
                                 double sample_rate, int sides, double* amplitude, double* phase_out,
                                 double* frequencies, po_peak* peak, double* scratch, int lean) {
  const int size = p->size;
  double* frame = scratch;
  double* re = scratch + size;
  double* im = scratch + 2 * (size_t)size;
  double* mag = scratch + 3 * (size_t)size;
  double* ang = scratch + 4 * (size_t)size;
  const int limit = len < size ? len : size;
  for (int i = 0; i < limit; i += 1) frame[i] = samples[i];
  for (int i = limit; i < size; i += 1) frame[i] = 0.0;
  for (int i = 0; i < size; i += 1) frame[i] = frame[i] * window[i]; /* applyWindow */
  po_fft_transform(p, frame, NULL, re, im, 0);
  po_magnitude(re, im, size, mag);
  if (!lean || phase_out) po_phase(re, im, size, ang);
  const int bins = sides == PO_SIDES_ONE ? size / 2 + 1 : size;
  double* amp = amplitude ? amplitude : frame; /* frame is free again */
  if (sides == PO_SIDES_ONE)
    po_scale_amplitude_one_sided(mag, size, amp);
  else
    po_scale_amplitude_two_sided(mag, size, amp);
  if (phase_out) memcpy(phase_out, ang, sizeof(double) * (size_t)bins);
  /* findPeak reads frequencies[index] only: compute it in place, same op order */
  const double scale = sample_rate / (double)size;
  if (frequencies)
    for (int i = 0; i < bins; i += 1) frequencies[i] = (double)i * scale;
  if (peak) {
    /* inline findPeak without materialising the axis */
    int max_index = 0;
    double max_value = amp[0];
    int has_non_dc = 0, non_dc_index = 0;
    double non_dc_value = 0.0;
    for (int i = 1; i < bins; i += 1) {
      const double v = amp[i];
      if (v > non_dc_value) {
        non_dc_value = v;
        non_dc_index = i;
      }
      if (v > 0) has_non_dc = 1;
      if (v > max_value) {
        max_value = v;
        max_index = i;
      }
    }
    const int index = has_non_dc ? non_dc_index : max_index;
    peak->index = index;
    peak->_pad = 0;
    peak->frequency = (double)index * scale;
    peak->amplitude = amp[index];
    peak->phase = (lean && !phase_out) ? atan2(im[index], re[index]) : ang[index]; /* peak.phase = phaseBins[peak.index] */
  }
}
PO_API void po_spectrum_frame(const po_plan* p, const double* window, const double* samples, int len,
                              double sample_rate, int sides, double* amplitude, double* phase_out,
                              double* frequencies, po_peak* peak, double* scratch) {
  po_spectrum_frame_ex(p, window, samples, len, sample_rate, sides, amplitude, phase_out, frequencies, peak, scratch, 0);
}

/* Batched driver over `batch` frames taken at samples + f*hop (element units), each
 * frame_len samples long, dtype f32 or f64 (the Float32Array -> Float64Array widening of
 * src/effect/index.ts:72-79).  Output rows are dense: bins per frame.  Any output may be NULL.
 * threads <= 1 runs single-threaded like the JS reference; threads > 1 splits frames
 * statically over OpenMP threads (used only by bench.py's reference arm). Returns threads used. */
static int po_spectrum_batch_impl(const po_plan* p, int window_type, const void* samples, int dtype, int frame_len,
                                  long long hop, long long batch, double sample_rate, int sides, double* amplitude,
                                  double* phase_out, po_peak* peaks, int threads, int lean) {
  const int size = p->size;
  const int bins = sides == PO_SIDES_ONE ? size / 2 + 1 : size;
  double* window = (double*)malloc(sizeof(double) * (size_t)size);
  po_create_window(window_type, size, window);
  int used = 1;
#ifdef _OPENMP
  if (threads > 1) used = threads;
#pragma omp parallel num_threads(used)
#endif
  {
    double* scratch = (double*)malloc(sizeof(double) * (size_t)size * 5);
    double* wide = (double*)malloc(sizeof(double) * (size_t)(frame_len > 0 ? frame_len : 1));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (long long f = 0; f < batch; f += 1) {
      const double* src;
      if (dtype == PO_F32) {
        const float* s = (const float*)samples + f * hop;
        const int lim = frame_len < size ? frame_len : size;
        for (int i = 0; i < lim; i += 1) wide[i] = (double)s[i];
        src = wide;
      } else {
        src = (const double*)samples + f * hop;
      }
      po_spectrum_frame_ex(p, window, src, frame_len, sample_rate, sides,
                           amplitude ? amplitude + f * (long long)bins : NULL,
                           phase_out ? phase_out + f * (long long)bins : NULL, NULL, peaks ? peaks + f : NULL,
                           scratch, lean);
    }
    free(scratch);
    free(wide);
  }
  free(window);
  return used;
}
PO_API int po_spectrum_batch(const po_plan* p, int window_type, const void* samples, int dtype, int frame_len,
                             long long hop, long long batch, double sample_rate, int sides, double* amplitude,
                             double* phase_out, po_peak* peaks, int threads) {
  return po_spectrum_batch_impl(p, window_type, samples, dtype, frame_len, hop, batch, sample_rate, sides, amplitude,
                                phase_out, peaks, threads, 0);
}
/* bench.py --impl reference: only the outputs the GPU arm produces (see po_spectrum_frame_ex) */
PO_API int po_spectrum_batch_lean(const po_plan* p, int window_type, const void* samples, int dtype, int frame_len,
                                  long long hop, long long batch, double sample_rate, int sides, double* amplitude,
                                  double* phase_out, po_peak* peaks, int threads) {
  return po_spectrum_batch_impl(p, window_type, samples, dtype, frame_len, hop, batch, sample_rate, sides, amplitude,
                                phase_out, peaks, threads, 1);
}

/* Batched forward/inverse transform, frames contiguous (stride = size). in_im may be NULL. */
PO_API int po_fft_batch(const po_plan* p, const double* in_re, const double* in_im, double* out_re,
                        double* out_im, long long batch, int inverse, int threads) {
  const long long size = p->size;
  int used = 1;
#ifdef _OPENMP
  if (threads > 1) used = threads;
#pragma omp parallel for schedule(static) num_threads(used)
#endif
  for (long long f = 0; f < batch; f += 1)
    po_fft_transform(p, in_re + f * size, in_im ? in_im + f * size : NULL, out_re + f * size,
                     out_im + f * size, inverse);
  return used;
}

/* bench/run.ts:20-25  guardrail checksum: sum of re*0.001 then im*0.002, interleaved per bin */
PO_API double po_bench_checksum(const double* re, const double* im, int n) {
  double checksum = 0;
  for (int i = 0; i < n; i += 1) {
    checksum += re[i] * 0.001;
    checksum += im[i] * 0.002;
  }
  return checksum;
}

PO_API int po_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
