/*
 * pragma_b200.h - C ABI of libpragma_b200.so, the B200 (sm_100a) implementation of
 * pragma-dsp's FFT/spectrum hot path.
 *
 * The reference (eliesgalvira/pragma-dsp v0.1.0) is pure TypeScript and has no FFI seam; the
 * seams a drop-in replaces are its ES-module exports.  Every entry point below cites the
 * reference interface it stands behind (paths relative to the reference root).  The Node-API
 * addon (napi/pragma_napi.cc) and the Python host (pragma_dsp_b200/) are thin shims over
 * exactly these symbols; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; pdsp_last_error() returns the
 *     message of the calling thread's last failure ("pragma-dsp/b200: ...").
 *   - argument validation that carries the reference's error messages (power-of-two size,
 *     input length) stays in the host language; the library re-checks and fails with its own text.
 *   - "host" entry points take ordinary host pointers (JS typed-array memory), stage through
 *     pinned buffers owned by the context, run on the context's stream and return after the
 *     results are in the caller's arrays (the reference is synchronous).  Jobs of up to 256 KB (input +
 *     outputs: Radix2Fft.forward or spectrum() on one or a few frames) take the single-call fast lane:
 *     one kernel launch reading and writing a host-mapped pinned buffer, completion by a doorbell word;
 *     larger jobs are cut into chunks and pipelined (H2D, kernel, D2H) over three streams.
 *   - "dev" entry points take device pointers and a CUDA stream and only enqueue work.
 *   - there is no CPU execution path: without a CUDA device pdsp_ctx_create fails.
 */
#ifndef PRAGMA_B200_H
#define PRAGMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDSP_ABI_VERSION 1

typedef struct pdsp_ctx pdsp_ctx;   /* one per device: stream, pinned + device staging, plan cache */
typedef struct pdsp_plan pdsp_plan; /* one per (size, precision): device twiddle / window tables */

enum pdsp_precision { PDSP_F32 = 0, PDSP_F64 = 1 };                                  /* compute + output type */
enum pdsp_window { PDSP_WIN_RECT = 0, PDSP_WIN_HANN = 1, PDSP_WIN_HAMMING = 2, PDSP_WIN_BLACKMAN = 3 };
enum pdsp_sides { PDSP_SIDES_ONE = 0, PDSP_SIDES_TWO = 1 };

/* findPeak result (src/public/spectrum.ts:15-20 SpectrumPeak).  Layout by plan precision. */
typedef struct { int32_t index; int32_t _pad; double frequency, amplitude, phase; } pdsp_peak_f64; /* 32 B */
typedef struct { int32_t index; float frequency, amplitude, phase; } pdsp_peak_f32;                /* 16 B */

/* spectrum() options (src/public/spectrum.ts:29-34 SpectrumOptions; src/effect/index.ts:53-58) plus
 * the batch addressing the reference leaves to the caller (spectrumStream frames). */
typedef struct {
  int32_t sample_dtype;  /* pdsp_precision of the samples buffer (Float32Array / Float64Array) */
  int32_t frame_len;     /* samples per frame; < fft size zero-pads, > fft size truncates (buildFrame) */
  int64_t hop;           /* distance between frame starts, in samples (= frame_len for disjoint frames) */
  int64_t batch;         /* number of frames */
  int32_t window;        /* pdsp_window */
  int32_t sides;         /* pdsp_sides */
  double sample_rate;    /* > 0 */
  int32_t raw_magnitude; /* 1: amplitude output is the unscaled |X| (magnitude(), xform/fourier.ts:98-109) */
  int32_t fft_shift;     /* 1 (two-sided only): rows are stored fftShift-ed, bin k at (k + N/2) mod N - fftShift
                            (src/xform/fourier.ts:122-134) fused into the store; peak.index stays the unshifted bin */
} pdsp_spectrum_desc;

/* ---- library / context ------------------------------------------------------------------ */
int pdsp_abi_version(void);
const char* pdsp_last_error(void);
int pdsp_device_count(int* count);
int pdsp_ctx_create(int device, pdsp_ctx** ctx);
int pdsp_ctx_destroy(pdsp_ctx* ctx);
int pdsp_ctx_sync(pdsp_ctx* ctx);
int pdsp_ctx_device(const pdsp_ctx* ctx);
/* Tuning / test hook: sets one tunable of the context ("staged", "chunk_bytes", "big_tma", "big_interleave",
 * "big_prefetch", "big_chunk", "big_factors", "big_resident", "fast", "doorbell", "copy_threads", "big_v2", "variant"); value NULL or "" restores the default.  The same
 * tunables are read ONCE from the environment (PDSP_<KEY>) when the context is created; nothing on the launch
 * path reads the environment.  Results do not depend on them (parity tests run every setting). */
int pdsp_ctx_tune(pdsp_ctx* ctx, const char* key, const char* value);
int pdsp_ctx_sm_count(const pdsp_ctx* ctx);
/* number of kernel launches issued through this context so far (bench.py's gpu_launches) */
int64_t pdsp_ctx_launch_count(const pdsp_ctx* ctx);
/* number of host calls served by the single-call fast lane (small jobs: one launch through a host-mapped pinned
 * buffer, completion by doorbell; see "host entry points" above) */
int64_t pdsp_ctx_fast_call_count(const pdsp_ctx* ctx);
/* an empty launch through the fast lane (launch + doorbell round trip): the floor under any one-frame call's latency */
int pdsp_ctx_ping(pdsp_ctx* ctx);

/* ---- pure host helpers (no device work) -------------------------------------------------- */
/* src/core/fft.ts:16 isPowerOfTwo, :18-23 nextPowerOfTwo */
int pdsp_is_power_of_two(int32_t n);
int32_t pdsp_next_power_of_two(int32_t n);
/* src/xform/fourier.ts:14-52 createWindow(type, size) -> out[size] (binary64, symmetric form) */
int pdsp_create_window(int window, int32_t size, double* out);
/* src/xform/fourier.ts:147-165 binFrequencies(size, sampleRate, sides) -> out[bins]; *bins is set */
int pdsp_bin_frequencies(int32_t size, double sample_rate, int sides, double* out, int32_t* bins);

/* ---- plans: new Radix2Fft(size) / new FFT(size) (src/core/fft.ts:68-75, xform/fourier.ts:73-79);
 *      cached per context like FourierLive's Map (src/effect/index.ts:30-40) ------------------ */
int pdsp_plan_get(pdsp_ctx* ctx, int32_t size, int precision, pdsp_plan** plan);
int32_t pdsp_plan_size(const pdsp_plan* plan);
/* Frees the scratch planes the large-transform path keeps per (plan, stream); call it before destroying a stream that
 * was passed to a *_dev entry point with an FFT size above 16384 (NULL = the context's stream).  The ingestion ring
 * does this for its own streams in pdsp_ingest_close. */
int pdsp_plan_release_stream(pdsp_plan* plan, void* stream);
int pdsp_plan_precision(const pdsp_plan* plan);

/* ---- Radix2Fft / FFT methods, host buffers ------------------------------------------------- */
/* forward(input, out) (src/core/fft.ts:77-79): `batch` real frames of plan size, contiguous;
 * in_dtype says whether `in` is float or double; out planes are double[batch*size], all N bins. */
int pdsp_fft_forward_real(pdsp_plan* plan, const void* in, int in_dtype, int64_t batch, double* out_re,
                          double* out_im);
/* forwardComplex(input, out) (src/core/fft.ts:81-83) */
int pdsp_fft_forward_complex(pdsp_plan* plan, const double* in_re, const double* in_im, int64_t batch,
                             double* out_re, double* out_im);
/* inverse(input, out) (src/core/fft.ts:85-87): conjugate transform times 1/N */
int pdsp_fft_inverse(pdsp_plan* plan, const double* in_re, const double* in_im, int64_t batch, double* out_re,
                     double* out_im);
/* magnitude(c, out) / phase(c, out) (src/xform/fourier.ts:98-120): elementwise over n values */
int pdsp_magnitude(pdsp_ctx* ctx, const double* re, const double* im, int64_t n, double* out);
int pdsp_phase(pdsp_ctx* ctx, const double* re, const double* im, int64_t n, double* out);
/* applyWindow(input, window, out) (src/xform/fourier.ts:54-67): out[i] = input[i] * window[i] */
int pdsp_apply_window(pdsp_ctx* ctx, const double* input, const double* window, int64_t n, double* out);
/* fftShift(input, out) (src/xform/fourier.ts:122-134): out[i] = input[(i + floor(n/2)) % n];
 * fftShiftComplex (:136-145) is this applied to each plane */
int pdsp_fft_shift(pdsp_ctx* ctx, const double* input, int64_t n, double* out);

/* ---- spectrum() / spectrumFx / spectrumStream chunk, host buffers -------------------------
 * (src/public/spectrum.ts:107-142, src/effect/index.ts:143-194).  Outputs are in the plan's
 * precision (float for PDSP_F32, double for PDSP_F64), dense rows of `bins` = N/2+1 (one) or N
 * (two) per frame; amplitude / phase / peaks may each be NULL. peaks points to pdsp_peak_f32[batch]
 * or pdsp_peak_f64[batch].  FFT sizes up to 16384 run as one fused kernel; larger sizes (up to 2^28) run as
 * frame-build + multi-pass transform + epilogue kernels with the same results. */
int pdsp_spectrum(pdsp_plan* plan, const pdsp_spectrum_desc* desc, const void* samples, void* amplitude,
                  void* phase, void* peaks);

/* ---- device-resident variants: pointers are device memory, work is enqueued on `stream`
 *      (a cudaStream_t passed as void*; NULL = the context's own stream, which is created non-blocking and is therefore
 *      NOT ordered with work on the legacy default stream - whose handle is also 0.  A caller working on the default
 *      stream passes cudaStreamLegacy ((void*)0x1) or cudaStreamPerThread ((void*)0x2), or synchronises first).  Scratch planes of the large
 *      transforms are kept per (plan, stream): calls on different streams may overlap, calls on one
 *      stream are ordered by it. ------------------------------------------------------------- */
int pdsp_spectrum_dev(pdsp_plan* plan, const pdsp_spectrum_desc* desc, const void* d_samples, void* d_amplitude,
                      void* d_phase, void* d_peaks, void* stream);
/* Frame-sharded multi-GPU form (BASELINE config C5; no reference counterpart - the reference is
 * single-process): as pdsp_spectrum_dev, and in addition every finished peak record of local frame f is
 * stored by the same kernel into peer_peaks[g] + (record_offset + f) records for g < n_peers - buffers of
 * the other ranks' GPUs mapped with pdsp_ipc_open (NVLink peer stores fused into the epilogue; no
 * separate all-gather).  d_peaks is the local workspace (required).  The caller synchronises the ranks
 * before reading a gathered buffer. */
int pdsp_spectrum_dev_gather(pdsp_plan* plan, const pdsp_spectrum_desc* desc, const void* d_samples, void* d_amplitude,
                             void* d_phase, void* d_peaks, void* const* peer_peaks, int n_peers, int64_t record_offset,
                             void* stream);
/* CUDA IPC plumbing for the peer buffers: export a device allocation made by pdsp_dev_alloc as a 64-byte
 * handle, open / close another process's handle on this context's device. */
int pdsp_ipc_export(pdsp_ctx* ctx, void* d_ptr, unsigned char handle[64]);
int pdsp_ipc_open(pdsp_ctx* ctx, const unsigned char handle[64], void** d_ptr);
int pdsp_ipc_close(pdsp_ctx* ctx, void* d_ptr);
/* real forward: out planes in plan precision; full != 0 writes all N bins, else N/2+1 */
int pdsp_fft_forward_real_dev(pdsp_plan* plan, const void* d_in, int in_dtype, int64_t batch, void* d_out_re,
                              void* d_out_im, int full, void* stream);
/* complex forward / inverse on planar arrays in plan precision; d_in_im may be NULL */
int pdsp_fft_complex_dev(pdsp_plan* plan, const void* d_in_re, const void* d_in_im, int64_t batch, void* d_out_re,
                         void* d_out_im, int inverse, void* stream);

/* "next" row: frequency-domain product on device-resident planar arrays, so that
 * forward -> multiply -> inverse (FFT convolution, test/fluent/chain.test.ts:287-316) needs no host round
 * trip: out = a * (conj_b ? conj(b) : b) * scale (src/math/complex.ts:87-105 mul, conj, scale fused). */
int pdsp_complex_mul_dev(pdsp_ctx* ctx, int precision, const void* d_a_re, const void* d_a_im, const void* d_b_re,
                         const void* d_b_im, int conj_b, double scale, int64_t n, void* d_out_re, void* d_out_im,
                         void* stream);

/* ---- device / pinned-host buffers owned by the library (what createComplexArray hands out
 *      behind the addon: src/core/fft.ts:6-14) ------------------------------------------------ */
int pdsp_dev_alloc(pdsp_ctx* ctx, size_t bytes, void** d_ptr);
int pdsp_dev_free(pdsp_ctx* ctx, void* d_ptr);
int pdsp_host_alloc(pdsp_ctx* ctx, size_t bytes, void** h_ptr); /* pinned */
int pdsp_host_free(pdsp_ctx* ctx, void* h_ptr);
int pdsp_memcpy_h2d(pdsp_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, void* stream);
int pdsp_memcpy_d2h(pdsp_ctx* ctx, void* h_dst, const void* d_src, size_t bytes, void* stream);

/* ---- ingestion ring: spectrumStream (src/effect/index.ts:190-194) maps a stream of frames
 *      1:1 and in order.  The ring batches them without the caller assembling batches: frames are copied
 *      into a pinned chunk as they arrive, a full chunk goes out on its own stream (H2D, fused kernel, D2H)
 *      while the next one fills, and results come back in arrival order.  desc->hop and desc->batch are
 *      ignored (frames are frame_len samples, packed by the ring). ------------------------------------- */
typedef struct pdsp_ingest pdsp_ingest;
int pdsp_ingest_open(pdsp_plan* plan, const pdsp_spectrum_desc* desc, int want_amplitude, int want_phase, int want_peaks,
                     int64_t frames_per_chunk, int depth, pdsp_ingest** ring);
/* copies `count` frames, `stride` samples apart (0 = frame_len); *accepted < count means the ring is full */
int pdsp_ingest_push(pdsp_ingest* ring, const void* frames, int64_t count, int64_t stride, int64_t* accepted);
/* as pdsp_ingest_push for frames in pinned (page-locked) host memory - e.g. an ArrayBuffer from pdsp_host_alloc: no host
 * copy, the frames go to the device by DMA straight from `frames`, which must stay unchanged until they have been popped.
 * A chunk holds either copied or pinned frames (flush between the two kinds). */
int pdsp_ingest_push_pinned(pdsp_ingest* ring, const void* frames, int64_t count, int64_t stride, int64_t* accepted);
/* non-blocking progress: frames that pdsp_ingest_pop would return without waiting, frames sent and still in flight,
 * frames waiting in the partially filled chunk (any pointer may be NULL).  What a latency-bound caller polls:
 * flush when nothing is in flight, pop what is finished. */
int pdsp_ingest_ready(pdsp_ingest* ring, int64_t* finished, int64_t* in_flight, int64_t* pending);
/* sends the partially filled chunk (end of stream / latency bound) */
int pdsp_ingest_flush(pdsp_ingest* ring);
/* up to max_frames finished frames, in arrival order, appended densely to the caller's arrays (NULL = skip);
 * waits for chunks already sent, never for a partially filled one; *got may be 0 */
int pdsp_ingest_pop(pdsp_ingest* ring, void* amplitude, void* phase, void* peaks, int64_t max_frames, int64_t* got);
int pdsp_ingest_close(pdsp_ingest* ring);

/* ---- device groups: frame sharding over the GPUs of one box (BASELINE config C5).  No reference
 *      counterpart - the reference is single-threaded; frames are independent (spectrumStream is a pure 1:1 map,
 *      src/effect/index.ts:190-194), so device i of n owns the contiguous block [i*ceil(F/n), (i+1)*ceil(F/n)) of the
 *      frame range.  One process, one context per device, peer access enabled between the devices. -------------------- */
typedef struct pdsp_group pdsp_group;
int pdsp_group_create(const int* devices, int n, pdsp_group** group); /* 1..8 distinct device ordinals */
int pdsp_group_destroy(pdsp_group* group);
int pdsp_group_size(const pdsp_group* group);
pdsp_ctx* pdsp_group_ctx(pdsp_group* group, int i); /* the i-th device's context (owned by the group) */
/* pdsp_spectrum over all devices of the group: host frames in, host results out.  Every device runs its block through
 * its own staging pipeline on its own host thread (pinned to the CPUs of the device's NUMA node); results land in the
 * caller's arrays at the block's offset, i.e. gathered on the way out.  Same arguments and results as pdsp_spectrum
 * (size / precision select each device's plan). */
int pdsp_group_spectrum(pdsp_group* group, int32_t size, int precision, const pdsp_spectrum_desc* desc, const void* samples,
                        void* amplitude, void* phase, void* peaks);
/* Device-resident sharded form (desc->batch = frames of the whole job).  Device i holds its block's samples at
 * d_samples[i] and gets block-local rows in d_amplitude[i] / d_phase[i] (arrays or entries may be NULL).  d_peaks[i] is
 * device i's GATHER buffer of desc->batch records: every device's kernel stores its block's records into all of them
 * (NVLink peer stores fused into the kernel - the all-gather of per-frame peaks), so after pdsp_group_sync every device
 * holds the peaks of all frames.  gather_root >= 0 additionally copies all blocks' amplitude / phase rows into
 * d_amplitude_all / d_phase_all (desc->batch rows on device gather_root; either may be NULL) by peer DMA queued behind
 * each block's kernel: the spectra gather "where the caller requests them on one device".  Work is queued on each
 * context's own stream. */
int pdsp_group_spectrum_dev(pdsp_group* group, int32_t size, int precision, const pdsp_spectrum_desc* desc,
                            const void* const* d_samples, void* const* d_amplitude, void* const* d_phase, void* const* d_peaks,
                            int gather_root, void* d_amplitude_all, void* d_phase_all);
int pdsp_group_sync(pdsp_group* group);

#ifdef __cplusplus
}
#endif
#endif /* PRAGMA_B200_H */
