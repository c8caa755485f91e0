// pragma_napi.cc - Node-API addon: the thin shim between the TypeScript host (ts/) and the C ABI
// (include/pragma_b200.h).  One addon function per C entry point; typed-array memory is borrowed
// for the duration of the synchronous call (napi_get_typedarray_info), results land in the
// caller's Float64Array/Float32Array before the call returns, CUDA failures become JS Errors
// carrying pdsp_last_error().  Argument validation with the reference's messages stays in TS.
//
// Not executable in this image (no Node.js); compile-checked against napi/node_api_min.h by
// tests/test_napi_shim.py.  Build: g++ -shared -fPIC -Iinclude napi/pragma_napi.cc
//        -Lpragma_dsp_b200 -lpragma_b200 -o pragma_b200.node
#ifdef PDSP_HAVE_NODE_API_H
#include <node_api.h>
#else
#include "node_api_min.h"
#endif
#include <string.h>

#include "../include/pragma_b200.h"

namespace {

struct TA {
  void* data = nullptr;
  size_t length = 0;
  napi_typedarray_type type = napi_uint8_array;
  bool present = false;
};

bool get_ta(napi_env env, napi_value v, TA* out) {
  napi_valuetype t;
  if (napi_typeof(env, v, &t) != napi_ok) return false;
  if (t == napi_undefined || t == napi_null) return true;  // optional output
  bool is = false;
  if (napi_is_typedarray(env, v, &is) != napi_ok || !is) return false;
  napi_value ab;
  size_t off;
  if (napi_get_typedarray_info(env, v, &out->type, &out->length, &out->data, &ab, &off) != napi_ok) return false;
  out->present = true;
  return true;
}

napi_value fail(napi_env env, const char* msg) {
  napi_throw_error(env, nullptr, msg);
  return nullptr;
}
napi_value fail_last(napi_env env) { return fail(env, pdsp_last_error()); }
napi_value undefined(napi_env env) {
  napi_value u;
  napi_get_undefined(env, &u);
  return u;
}
int dtype_of(const TA& a) { return a.type == napi_float64_array ? PDSP_F64 : PDSP_F32; }
bool is_float(const TA& a) { return a.type == napi_float64_array || a.type == napi_float32_array; }

#define ARGS(n)                                                                   \
  size_t argc = n;                                                                \
  napi_value argv[n];                                                             \
  if (napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr) != napi_ok || argc < n) \
    return fail(env, "pragma-dsp/b200: wrong number of arguments");

void ctx_finalize(napi_env, void* data, void*) { pdsp_ctx_destroy(static_cast<pdsp_ctx*>(data)); }

// contextCreate(device) -> external<pdsp_ctx>
napi_value ContextCreate(napi_env env, napi_callback_info info) {
  ARGS(1)
  int32_t dev = 0;
  napi_get_value_int32(env, argv[0], &dev);
  pdsp_ctx* c = nullptr;
  if (pdsp_ctx_create(dev, &c)) return fail_last(env);
  napi_value ext;
  napi_create_external(env, c, ctx_finalize, nullptr, &ext);
  return ext;
}

// planGet(ctx, size, precision) -> external<pdsp_plan>   (new Radix2Fft(size) / FourierLive.fft(size))
napi_value PlanGet(napi_env env, napi_callback_info info) {
  ARGS(3)
  void* c = nullptr;
  int32_t size = 0, prec = PDSP_F64;
  napi_get_value_external(env, argv[0], &c);
  napi_get_value_int32(env, argv[1], &size);
  napi_get_value_int32(env, argv[2], &prec);
  pdsp_plan* p = nullptr;
  if (pdsp_plan_get(static_cast<pdsp_ctx*>(c), size, prec, &p)) return fail_last(env);
  napi_value ext;
  napi_create_external(env, p, nullptr, nullptr, &ext);  // owned by the context
  return ext;
}

// fftForwardReal(plan, input: Float32Array|Float64Array, outReal: Float64Array, outImag: Float64Array)
napi_value FftForwardReal(napi_env env, napi_callback_info info) {
  ARGS(4)
  void* p = nullptr;
  napi_get_value_external(env, argv[0], &p);
  TA in, re, im;
  if (!get_ta(env, argv[1], &in) || !get_ta(env, argv[2], &re) || !get_ta(env, argv[3], &im) || !in.present ||
      !re.present || !im.present || !is_float(in) || re.type != napi_float64_array || im.type != napi_float64_array)
    return fail(env, "pragma-dsp/b200: expected (plan, Float32Array|Float64Array, Float64Array, Float64Array)");
  const size_t n = (size_t)pdsp_plan_size(static_cast<pdsp_plan*>(p));
  if (n == 0 || in.length % n || re.length < in.length || im.length < in.length)
    return fail(env, "pragma-dsp/b200: array lengths do not match the plan size");
  if (pdsp_fft_forward_real(static_cast<pdsp_plan*>(p), in.data, dtype_of(in), (int64_t)(in.length / n),
                            static_cast<double*>(re.data), static_cast<double*>(im.data)))
    return fail_last(env);
  return undefined(env);
}

napi_value complex_common(napi_env env, napi_callback_info info, bool inverse) {
  ARGS(5)
  void* p = nullptr;
  napi_get_value_external(env, argv[0], &p);
  TA ire, iim, ore, oim;
  if (!get_ta(env, argv[1], &ire) || !get_ta(env, argv[2], &iim) || !get_ta(env, argv[3], &ore) ||
      !get_ta(env, argv[4], &oim) || ire.type != napi_float64_array || iim.type != napi_float64_array ||
      ore.type != napi_float64_array || oim.type != napi_float64_array)
    return fail(env, "pragma-dsp/b200: expected (plan, Float64Array x4)");
  const size_t n = (size_t)pdsp_plan_size(static_cast<pdsp_plan*>(p));
  if (n == 0 || ire.length % n || iim.length != ire.length || ore.length < ire.length || oim.length < ire.length)
    return fail(env, "pragma-dsp/b200: array lengths do not match the plan size");
  const int64_t batch = (int64_t)(ire.length / n);
  const int rc = inverse ? pdsp_fft_inverse(static_cast<pdsp_plan*>(p), static_cast<double*>(ire.data),
                                            static_cast<double*>(iim.data), batch, static_cast<double*>(ore.data),
                                            static_cast<double*>(oim.data))
                         : pdsp_fft_forward_complex(static_cast<pdsp_plan*>(p), static_cast<double*>(ire.data),
                                                    static_cast<double*>(iim.data), batch,
                                                    static_cast<double*>(ore.data), static_cast<double*>(oim.data));
  if (rc) return fail_last(env);
  return undefined(env);
}
napi_value FftForwardComplex(napi_env env, napi_callback_info info) { return complex_common(env, info, false); }
napi_value FftInverse(napi_env env, napi_callback_info info) { return complex_common(env, info, true); }

napi_value elementwise(napi_env env, napi_callback_info info, bool mag) {
  ARGS(4)
  void* c = nullptr;
  napi_get_value_external(env, argv[0], &c);
  TA re, im, out;
  if (!get_ta(env, argv[1], &re) || !get_ta(env, argv[2], &im) || !get_ta(env, argv[3], &out) ||
      re.type != napi_float64_array || im.type != napi_float64_array || out.type != napi_float64_array ||
      im.length < re.length || out.length < re.length)
    return fail(env, "pragma-dsp/b200: expected (ctx, Float64Array, Float64Array, Float64Array)");
  const int rc = mag ? pdsp_magnitude(static_cast<pdsp_ctx*>(c), static_cast<double*>(re.data),
                                      static_cast<double*>(im.data), (int64_t)re.length, static_cast<double*>(out.data))
                     : pdsp_phase(static_cast<pdsp_ctx*>(c), static_cast<double*>(re.data),
                                  static_cast<double*>(im.data), (int64_t)re.length, static_cast<double*>(out.data));
  if (rc) return fail_last(env);
  return undefined(env);
}
napi_value Magnitude(napi_env env, napi_callback_info info) { return elementwise(env, info, true); }
napi_value Phase(napi_env env, napi_callback_info info) { return elementwise(env, info, false); }

int32_t get_i32(napi_env env, napi_value obj, const char* key, int32_t dflt) {
  napi_value v;
  napi_valuetype t;
  if (napi_get_named_property(env, obj, key, &v) != napi_ok || napi_typeof(env, v, &t) != napi_ok || t != napi_number)
    return dflt;
  int32_t r = dflt;
  napi_get_value_int32(env, v, &r);
  return r;
}
double get_f64(napi_env env, napi_value obj, const char* key, double dflt) {
  napi_value v;
  napi_valuetype t;
  if (napi_get_named_property(env, obj, key, &v) != napi_ok || napi_typeof(env, v, &t) != napi_ok || t != napi_number)
    return dflt;
  double r = dflt;
  napi_get_value_double(env, v, &r);
  return r;
}

// spectrum(plan, samples, {frameLen, hop, batch, window, sides, sampleRate, rawMagnitude},
//          amplitude|null, phase|null, peaks: Uint8Array|null)
napi_value Spectrum(napi_env env, napi_callback_info info) {
  ARGS(6)
  void* p = nullptr;
  napi_get_value_external(env, argv[0], &p);
  TA s, amp, ph, pk;
  if (!get_ta(env, argv[1], &s) || !get_ta(env, argv[3], &amp) || !get_ta(env, argv[4], &ph) ||
      !get_ta(env, argv[5], &pk) || !s.present || !is_float(s))
    return fail(env, "pragma-dsp/b200: expected (plan, Float32Array|Float64Array, options, out arrays)");
  pdsp_spectrum_desc d;
  memset(&d, 0, sizeof d);
  d.sample_dtype = dtype_of(s);
  d.frame_len = get_i32(env, argv[2], "frameLen", (int32_t)s.length);
  d.hop = get_i32(env, argv[2], "hop", d.frame_len);
  d.batch = get_i32(env, argv[2], "batch", 1);
  d.window = get_i32(env, argv[2], "window", PDSP_WIN_RECT);
  d.sides = get_i32(env, argv[2], "sides", PDSP_SIDES_ONE);
  d.sample_rate = get_f64(env, argv[2], "sampleRate", 1.0);
  d.raw_magnitude = get_i32(env, argv[2], "rawMagnitude", 0);
  d.fft_shift = get_i32(env, argv[2], "shift", 0);
  if (d.batch > 0 && d.frame_len > 0 && (size_t)((d.batch - 1) * d.hop + d.frame_len) > s.length)
    return fail(env, "pragma-dsp/b200: frames exceed the samples buffer");
  pdsp_plan* plan = static_cast<pdsp_plan*>(p);
  const size_t n = (size_t)pdsp_plan_size(plan);
  const size_t bins = d.sides == PDSP_SIDES_TWO ? n : n / 2 + 1;
  const napi_typedarray_type want = pdsp_plan_precision(plan) == PDSP_F64 ? napi_float64_array : napi_float32_array;
  const size_t pk_size = pdsp_plan_precision(plan) == PDSP_F64 ? sizeof(pdsp_peak_f64) : sizeof(pdsp_peak_f32);
  if ((amp.present && (amp.type != want || amp.length < bins * (size_t)d.batch)) ||
      (ph.present && (ph.type != want || ph.length < bins * (size_t)d.batch)) ||
      (pk.present && (pk.type != napi_uint8_array || pk.length < pk_size * (size_t)d.batch)))
    return fail(env, "pragma-dsp/b200: output arrays have the wrong type or length for this plan");
  if (pdsp_spectrum(plan, &d, s.data, amp.data, ph.data, pk.data)) return fail_last(env);
  return undefined(env);
}

// createWindow(type, size, out: Float64Array) ; binFrequencies(size, sampleRate, sides, out: Float64Array)
napi_value CreateWindow(napi_env env, napi_callback_info info) {
  ARGS(3)
  int32_t type = 0, size = 0;
  napi_get_value_int32(env, argv[0], &type);
  napi_get_value_int32(env, argv[1], &size);
  TA out;
  if (!get_ta(env, argv[2], &out) || out.type != napi_float64_array || (int64_t)out.length < size)
    return fail(env, "pragma-dsp/b200: expected (type, size, Float64Array(size))");
  if (pdsp_create_window(type, size, static_cast<double*>(out.data))) return fail_last(env);
  return undefined(env);
}
napi_value BinFrequencies(napi_env env, napi_callback_info info) {
  ARGS(4)
  int32_t size = 0, sides = 0;
  double fs = 1;
  napi_get_value_int32(env, argv[0], &size);
  napi_get_value_double(env, argv[1], &fs);
  napi_get_value_int32(env, argv[2], &sides);
  TA out;
  if (!get_ta(env, argv[3], &out) || out.type != napi_float64_array)
    return fail(env, "pragma-dsp/b200: expected (size, sampleRate, sides, Float64Array)");
  int32_t bins = 0;
  if (pdsp_bin_frequencies(size, fs, sides, nullptr, &bins)) return fail_last(env);
  if ((int64_t)out.length < bins) return fail(env, "pragma-dsp/b200: output array too short");
  if (pdsp_bin_frequencies(size, fs, sides, static_cast<double*>(out.data), &bins)) return fail_last(env);
  return undefined(env);
}

// hostAlloc(ctx, bytes) -> ArrayBuffer backed by pinned memory (createComplexArray hands these out so
// forward(input, out) DMA-writes straight into `out`; still an ordinary Float64Array to JS)
struct PinnedHint {
  pdsp_ctx* ctx;
};
void pinned_finalize(napi_env, void* data, void* hint) {
  PinnedHint* h = static_cast<PinnedHint*>(hint);
  pdsp_host_free(h->ctx, data);
  delete h;
}
napi_value HostAlloc(napi_env env, napi_callback_info info) {
  ARGS(2)
  void* c = nullptr;
  int64_t bytes = 0;
  napi_get_value_external(env, argv[0], &c);
  napi_get_value_int64(env, argv[1], &bytes);
  void* ptr = nullptr;
  if (bytes < 0 || pdsp_host_alloc(static_cast<pdsp_ctx*>(c), (size_t)bytes, &ptr)) return fail_last(env);
  memset(ptr, 0, (size_t)bytes);
  napi_value ab;
  PinnedHint* h = new PinnedHint{static_cast<pdsp_ctx*>(c)};
  if (napi_create_external_arraybuffer(env, ptr, (size_t)bytes, pinned_finalize, h, &ab) != napi_ok) {
    pdsp_host_free(static_cast<pdsp_ctx*>(c), ptr);
    delete h;
    return fail(env, "pragma-dsp/b200: external array buffers are not allowed by this runtime");
  }
  return ab;
}

// ---- ingestion ring (pdsp_ingest_*): what ts/effect/index.ts spectrumStream pushes Stream<Float32Array> frames into
void ingest_finalize(napi_env, void* data, void*) { pdsp_ingest_close(static_cast<pdsp_ingest*>(data)); }
// ingestOpen(plan, {frameLen, window, sides, sampleRate, sampleDtype}, wantAmp, wantPhase, wantPeaks, framesPerChunk, depth)
napi_value IngestOpen(napi_env env, napi_callback_info info) {
  ARGS(7)
  void* p = nullptr;
  napi_get_value_external(env, argv[0], &p);
  pdsp_spectrum_desc d;
  memset(&d, 0, sizeof d);
  d.sample_dtype = get_i32(env, argv[1], "sampleDtype", PDSP_F32);
  d.frame_len = get_i32(env, argv[1], "frameLen", 0);
  d.hop = d.frame_len;
  d.window = get_i32(env, argv[1], "window", PDSP_WIN_RECT);
  d.sides = get_i32(env, argv[1], "sides", PDSP_SIDES_ONE);
  d.sample_rate = get_f64(env, argv[1], "sampleRate", 1.0);
  int32_t wa = 0, wp = 0, wk = 0, depth = 3;
  int64_t per_chunk = 0;
  napi_get_value_int32(env, argv[2], &wa);
  napi_get_value_int32(env, argv[3], &wp);
  napi_get_value_int32(env, argv[4], &wk);
  napi_get_value_int64(env, argv[5], &per_chunk);
  napi_get_value_int32(env, argv[6], &depth);
  pdsp_ingest* ring = nullptr;
  if (pdsp_ingest_open(static_cast<pdsp_plan*>(p), &d, wa, wp, wk, per_chunk, depth, &ring)) return fail_last(env);
  napi_value ext;
  napi_create_external(env, ring, ingest_finalize, nullptr, &ext);
  return ext;
}
// ingestPush(ring, frames: Float32Array|Float64Array holding whole frames, count) -> frames accepted
napi_value IngestPush(napi_env env, napi_callback_info info) {
  ARGS(3)
  void* r = nullptr;
  napi_get_value_external(env, argv[0], &r);
  TA s;
  int64_t count = 0, accepted = 0;
  napi_get_value_int64(env, argv[2], &count);
  if (!get_ta(env, argv[1], &s) || !s.present || !is_float(s)) return fail(env, "pragma-dsp/b200: expected (ring, frames, count)");
  if (pdsp_ingest_push(static_cast<pdsp_ingest*>(r), s.data, count, 0, &accepted)) return fail_last(env);
  napi_value v;
  napi_create_double(env, (double)accepted, &v);
  return v;
}
napi_value IngestFlush(napi_env env, napi_callback_info info) {
  ARGS(1)
  void* r = nullptr;
  napi_get_value_external(env, argv[0], &r);
  if (pdsp_ingest_flush(static_cast<pdsp_ingest*>(r))) return fail_last(env);
  return undefined(env);
}
// ingestPop(ring, amplitude|null, phase|null, peaks: Uint8Array|null, maxFrames) -> frames returned
napi_value IngestPop(napi_env env, napi_callback_info info) {
  ARGS(5)
  void* r = nullptr;
  napi_get_value_external(env, argv[0], &r);
  TA amp, ph, pk;
  int64_t max_frames = 0, got = 0;
  napi_get_value_int64(env, argv[4], &max_frames);
  if (!get_ta(env, argv[1], &amp) || !get_ta(env, argv[2], &ph) || !get_ta(env, argv[3], &pk))
    return fail(env, "pragma-dsp/b200: expected (ring, amplitude|null, phase|null, peaks|null, maxFrames)");
  if (pdsp_ingest_pop(static_cast<pdsp_ingest*>(r), amp.data, ph.data, pk.data, max_frames, &got)) return fail_last(env);
  napi_value v;
  napi_create_double(env, (double)got, &v);
  return v;
}

napi_value AbiVersion(napi_env env, napi_callback_info) {
  napi_value v;
  napi_create_int32(env, pdsp_abi_version(), &v);
  return v;
}

void def(napi_env env, napi_value exports, const char* name, napi_callback cb) {
  napi_value fn;
  napi_create_function(env, name, strlen(name), cb, nullptr, &fn);
  napi_set_named_property(env, exports, name, fn);
}

}  // namespace

extern "C" __attribute__((visibility("default"))) napi_value napi_register_module_v1(napi_env env, napi_value exports) {
  def(env, exports, "abiVersion", AbiVersion);
  def(env, exports, "contextCreate", ContextCreate);
  def(env, exports, "planGet", PlanGet);
  def(env, exports, "fftForwardReal", FftForwardReal);
  def(env, exports, "fftForwardComplex", FftForwardComplex);
  def(env, exports, "fftInverse", FftInverse);
  def(env, exports, "magnitude", Magnitude);
  def(env, exports, "phase", Phase);
  def(env, exports, "spectrum", Spectrum);
  def(env, exports, "createWindow", CreateWindow);
  def(env, exports, "binFrequencies", BinFrequencies);
  def(env, exports, "hostAlloc", HostAlloc);
  def(env, exports, "ingestOpen", IngestOpen);
  def(env, exports, "ingestPush", IngestPush);
  def(env, exports, "ingestFlush", IngestFlush);
  def(env, exports, "ingestPop", IngestPop);
  return exports;
}
