// pragma_napi.cc - Node-API addon: the thin shim between the TypeScript host (ts/) and the C ABI
// (include/pragma_b200.h).  One addon function per C entry point; typed-array memory is borrowed
// for the duration of the synchronous call (napi_get_typedarray_info), results land in the
// caller's Float64Array/Float32Array before the call returns, CUDA failures become JS Errors
// carrying pdsp_last_error().  Argument validation with the reference's messages stays in TS; the
// shim itself checks every typed array's element type and length before a pointer reaches native code.
//
// Lifetime: every handle handed to JS is a tagged box.  Plans, ingestion rings and pinned buffers hold a
// reference on their context's box, and the context is destroyed when the last reference goes - so the
// order in which the garbage collector (or environment teardown) runs the finalisers does not matter.
//
// No Node.js in this image: the shim is compiled against napi/node_api_min.h and EXECUTED under the mock
// Node-API host of tests/napi_host (tests/test_napi_shim.py: against the emulated C-ABI library on the CPU,
// against libpragma_b200.so under -m gpu).  Build for Node:
//   g++ -shared -fPIC -Iinclude -DPDSP_HAVE_NODE_API_H napi/pragma_napi.cc -Lpragma_dsp_b200 -lpragma_b200 -o pragma_b200.node
#ifdef PDSP_HAVE_NODE_API_H
#include <node_api.h>
#else
#include "node_api_min.h"
#endif
#include <string.h>

#include "../include/pragma_b200.h"

namespace {

// ---------------------------------------------------------------------------------------- handles
enum : uint32_t { TAG_CTX = 0x70637478u, TAG_PLAN = 0x70706c6eu, TAG_RING = 0x7072696eu };

struct CtxBox {
  uint32_t tag = TAG_CTX;
  pdsp_ctx* ctx = nullptr;
  int refs = 1;  // the JS handle itself + one per plan / ring / pinned buffer alive
};
void ctx_unref(CtxBox* b) {
  if (--b->refs == 0) {
    pdsp_ctx_destroy(b->ctx);
    b->tag = 0;
    delete b;
  }
}
struct PlanBox {
  uint32_t tag = TAG_PLAN;
  pdsp_plan* plan = nullptr;  // owned by the context (plan cache)
  CtxBox* owner = nullptr;
};
struct RingBox {
  uint32_t tag = TAG_RING;
  pdsp_ingest* ring = nullptr;
  CtxBox* owner = nullptr;
  size_t frame_len = 0, bins = 0, peak_bytes = 0;
  napi_typedarray_type sample_type = napi_float32_array, out_type = napi_float64_array;
  bool want_amp = false, want_phase = false, want_peaks = false;
};

void ctx_finalize(napi_env, void* data, void*) { ctx_unref(static_cast<CtxBox*>(data)); }
void plan_finalize(napi_env, void* data, void*) {
  PlanBox* p = static_cast<PlanBox*>(data);
  CtxBox* o = p->owner;
  p->tag = 0;
  delete p;
  ctx_unref(o);
}
void ring_finalize(napi_env, void* data, void*) {
  RingBox* r = static_cast<RingBox*>(data);
  pdsp_ingest_close(r->ring);  // the context is still alive: this box holds a reference on it
  CtxBox* o = r->owner;
  r->tag = 0;
  delete r;
  ctx_unref(o);
}

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
bool is_f64(const TA& a) { return a.present && a.type == napi_float64_array; }

template <typename Box>
Box* unbox(napi_env env, napi_value v, uint32_t tag) {
  napi_valuetype t;
  void* p = nullptr;
  if (napi_typeof(env, v, &t) != napi_ok || t != napi_external || napi_get_value_external(env, v, &p) != napi_ok || !p)
    return nullptr;
  Box* b = static_cast<Box*>(p);
  return b->tag == tag ? b : nullptr;
}
#define CTX_ARG(var, v)                              \
  CtxBox* var = unbox<CtxBox>(env, v, TAG_CTX);      \
  if (!var) return fail(env, "pragma-dsp/b200: expected a context handle");
#define PLAN_ARG(var, v)                             \
  PlanBox* var = unbox<PlanBox>(env, v, TAG_PLAN);   \
  if (!var) return fail(env, "pragma-dsp/b200: expected a plan handle");
#define RING_ARG(var, v)                             \
  RingBox* var = unbox<RingBox>(env, v, TAG_RING);   \
  if (!var) return fail(env, "pragma-dsp/b200: expected an ingestion ring handle");

#define ARGS(n)                                                                   \
  size_t argc = n;                                                                \
  napi_value argv[n];                                                             \
  if (napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr) != napi_ok || argc < n) \
    return fail(env, "pragma-dsp/b200: wrong number of arguments");

// contextCreate(device) -> external<ctx>
napi_value ContextCreate(napi_env env, napi_callback_info info) {
  ARGS(1)
  int32_t dev = 0;
  napi_get_value_int32(env, argv[0], &dev);
  pdsp_ctx* c = nullptr;
  if (pdsp_ctx_create(dev, &c)) return fail_last(env);
  CtxBox* box = new CtxBox();
  box->ctx = c;
  napi_value ext;
  if (napi_create_external(env, box, ctx_finalize, nullptr, &ext) != napi_ok) {
    ctx_unref(box);
    return fail(env, "pragma-dsp/b200: napi_create_external failed");
  }
  return ext;
}

// planGet(ctx, size, precision) -> external<plan>   (new Radix2Fft(size) / FourierLive.fft(size))
napi_value PlanGet(napi_env env, napi_callback_info info) {
  ARGS(3)
  CTX_ARG(c, argv[0])
  int32_t size = 0, prec = PDSP_F64;
  napi_get_value_int32(env, argv[1], &size);
  napi_get_value_int32(env, argv[2], &prec);
  pdsp_plan* p = nullptr;
  if (pdsp_plan_get(c->ctx, size, prec, &p)) return fail_last(env);
  PlanBox* box = new PlanBox();
  box->plan = p;
  box->owner = c;
  c->refs++;
  napi_value ext;
  if (napi_create_external(env, box, plan_finalize, nullptr, &ext) != napi_ok) {
    plan_finalize(env, box, nullptr);
    return fail(env, "pragma-dsp/b200: napi_create_external failed");
  }
  return ext;
}

// fftForwardReal(plan, input: Float32Array|Float64Array, outReal: Float64Array, outImag: Float64Array)
napi_value FftForwardReal(napi_env env, napi_callback_info info) {
  ARGS(4)
  PLAN_ARG(p, argv[0])
  TA in, re, im;
  if (!get_ta(env, argv[1], &in) || !get_ta(env, argv[2], &re) || !get_ta(env, argv[3], &im) || !in.present ||
      !is_float(in) || !is_f64(re) || !is_f64(im))
    return fail(env, "pragma-dsp/b200: expected (plan, Float32Array|Float64Array, Float64Array, Float64Array)");
  const size_t n = (size_t)pdsp_plan_size(p->plan);
  if (n == 0 || in.length % n || re.length < in.length || im.length < in.length)
    return fail(env, "pragma-dsp/b200: array lengths do not match the plan size");
  if (pdsp_fft_forward_real(p->plan, in.data, dtype_of(in), (int64_t)(in.length / n), static_cast<double*>(re.data),
                            static_cast<double*>(im.data)))
    return fail_last(env);
  return undefined(env);
}

napi_value complex_common(napi_env env, napi_callback_info info, bool inverse) {
  ARGS(5)
  PLAN_ARG(p, argv[0])
  TA ire, iim, ore, oim;
  if (!get_ta(env, argv[1], &ire) || !get_ta(env, argv[2], &iim) || !get_ta(env, argv[3], &ore) ||
      !get_ta(env, argv[4], &oim) || !is_f64(ire) || !is_f64(iim) || !is_f64(ore) || !is_f64(oim))
    return fail(env, "pragma-dsp/b200: expected (plan, Float64Array x4)");
  const size_t n = (size_t)pdsp_plan_size(p->plan);
  if (n == 0 || ire.length % n || iim.length != ire.length || ore.length < ire.length || oim.length < ire.length)
    return fail(env, "pragma-dsp/b200: array lengths do not match the plan size");
  const int64_t batch = (int64_t)(ire.length / n);
  const int rc = inverse ? pdsp_fft_inverse(p->plan, static_cast<double*>(ire.data), static_cast<double*>(iim.data), batch,
                                            static_cast<double*>(ore.data), static_cast<double*>(oim.data))
                         : pdsp_fft_forward_complex(p->plan, static_cast<double*>(ire.data), static_cast<double*>(iim.data),
                                                    batch, static_cast<double*>(ore.data), static_cast<double*>(oim.data));
  if (rc) return fail_last(env);
  return undefined(env);
}
napi_value FftForwardComplex(napi_env env, napi_callback_info info) { return complex_common(env, info, false); }
napi_value FftInverse(napi_env env, napi_callback_info info) { return complex_common(env, info, true); }

// magnitude / phase (ctx, re, im, out) ; applyWindow (ctx, input, window, out): elementwise over re.length values
napi_value elementwise(napi_env env, napi_callback_info info, int op) {
  ARGS(4)
  CTX_ARG(c, argv[0])
  TA a, b, out;
  if (!get_ta(env, argv[1], &a) || !get_ta(env, argv[2], &b) || !get_ta(env, argv[3], &out) || !is_f64(a) || !is_f64(b) ||
      !is_f64(out) || b.length < a.length || out.length < a.length)
    return fail(env, "pragma-dsp/b200: expected (ctx, Float64Array, Float64Array, Float64Array) of matching lengths");
  double* x = static_cast<double*>(a.data);
  double* y = static_cast<double*>(b.data);
  double* o = static_cast<double*>(out.data);
  const int64_t n = (int64_t)a.length;
  const int rc = op == 0 ? pdsp_magnitude(c->ctx, x, y, n, o) : op == 1 ? pdsp_phase(c->ctx, x, y, n, o) : pdsp_apply_window(c->ctx, x, y, n, o);
  if (rc) return fail_last(env);
  return undefined(env);
}
napi_value Magnitude(napi_env env, napi_callback_info info) { return elementwise(env, info, 0); }
napi_value Phase(napi_env env, napi_callback_info info) { return elementwise(env, info, 1); }
napi_value ApplyWindow(napi_env env, napi_callback_info info) { return elementwise(env, info, 2); }

// fftShift(ctx, input: Float64Array, out: Float64Array)   (fftShiftComplex = one call per plane, in TS)
napi_value FftShift(napi_env env, napi_callback_info info) {
  ARGS(3)
  CTX_ARG(c, argv[0])
  TA in, out;
  if (!get_ta(env, argv[1], &in) || !get_ta(env, argv[2], &out) || !is_f64(in) || !is_f64(out) || out.length < in.length)
    return fail(env, "pragma-dsp/b200: expected (ctx, Float64Array, Float64Array) of matching lengths");
  if (in.length && pdsp_fft_shift(c->ctx, static_cast<double*>(in.data), (int64_t)in.length, static_cast<double*>(out.data)))
    return fail_last(env);
  return undefined(env);
}

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

// spectrum(plan, samples, {frameLen, hop, batch, window, sides, sampleRate, rawMagnitude, shift},
//          amplitude|null, phase|null, peaks: Uint8Array|null)
napi_value Spectrum(napi_env env, napi_callback_info info) {
  ARGS(6)
  PLAN_ARG(p, argv[0])
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
  if (d.batch < 0 || d.frame_len < 0 || d.hop < 0) return fail(env, "pragma-dsp/b200: negative frame length, hop or batch");
  if (d.batch > 0 && d.frame_len > 0 && (size_t)((d.batch - 1) * d.hop + d.frame_len) > s.length)
    return fail(env, "pragma-dsp/b200: frames exceed the samples buffer");
  const size_t n = (size_t)pdsp_plan_size(p->plan);
  const size_t bins = d.sides == PDSP_SIDES_TWO ? n : n / 2 + 1;
  const napi_typedarray_type want = pdsp_plan_precision(p->plan) == PDSP_F64 ? napi_float64_array : napi_float32_array;
  const size_t pk_size = pdsp_plan_precision(p->plan) == PDSP_F64 ? sizeof(pdsp_peak_f64) : sizeof(pdsp_peak_f32);
  if ((amp.present && (amp.type != want || amp.length < bins * (size_t)d.batch)) ||
      (ph.present && (ph.type != want || ph.length < bins * (size_t)d.batch)) ||
      (pk.present && (pk.type != napi_uint8_array || pk.length < pk_size * (size_t)d.batch)))
    return fail(env, "pragma-dsp/b200: output arrays have the wrong type or length for this plan");
  if (pdsp_spectrum(p->plan, &d, s.data, amp.data, ph.data, pk.data)) return fail_last(env);
  return undefined(env);
}

// createWindow(type, size, out: Float64Array) ; binFrequencies(size, sampleRate, sides, out: Float64Array)
napi_value CreateWindow(napi_env env, napi_callback_info info) {
  ARGS(3)
  int32_t type = 0, size = 0;
  napi_get_value_int32(env, argv[0], &type);
  napi_get_value_int32(env, argv[1], &size);
  TA out;
  if (!get_ta(env, argv[2], &out) || !is_f64(out) || (int64_t)out.length < size)
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
  if (!get_ta(env, argv[3], &out) || !is_f64(out)) return fail(env, "pragma-dsp/b200: expected (size, sampleRate, sides, Float64Array)");
  int32_t bins = 0;
  if (pdsp_bin_frequencies(size, fs, sides, nullptr, &bins)) return fail_last(env);
  if ((int64_t)out.length < bins) return fail(env, "pragma-dsp/b200: output array too short");
  if (pdsp_bin_frequencies(size, fs, sides, static_cast<double*>(out.data), &bins)) return fail_last(env);
  return undefined(env);
}

// hostAlloc(ctx, bytes) -> ArrayBuffer backed by pinned memory (createComplexArray hands these out so
// forward(input, out) DMA-writes straight into `out`; still an ordinary Float64Array to JS)
void pinned_finalize(napi_env, void* data, void* hint) {
  CtxBox* owner = static_cast<CtxBox*>(hint);
  pdsp_host_free(owner->ctx, data);
  ctx_unref(owner);
}
napi_value HostAlloc(napi_env env, napi_callback_info info) {
  ARGS(2)
  CTX_ARG(c, argv[0])
  int64_t bytes = 0;
  napi_get_value_int64(env, argv[1], &bytes);
  void* ptr = nullptr;
  if (bytes < 0) return fail(env, "pragma-dsp/b200: negative size");
  if (pdsp_host_alloc(c->ctx, (size_t)bytes, &ptr)) return fail_last(env);
  memset(ptr, 0, (size_t)bytes);
  napi_value ab;
  c->refs++;
  if (napi_create_external_arraybuffer(env, ptr, (size_t)bytes, pinned_finalize, c, &ab) != napi_ok) {
    pinned_finalize(env, ptr, c);
    return fail(env, "pragma-dsp/b200: external array buffers are not allowed by this runtime");
  }
  return ab;
}

// ---- ingestion ring (pdsp_ingest_*): what ts/effect/index.ts spectrumStream pushes Stream<Float32Array> frames into
// ingestOpen(plan, {frameLen, window, sides, sampleRate, sampleDtype}, wantAmp, wantPhase, wantPeaks, framesPerChunk, depth)
napi_value IngestOpen(napi_env env, napi_callback_info info) {
  ARGS(7)
  PLAN_ARG(p, argv[0])
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
  if (pdsp_ingest_open(p->plan, &d, wa, wp, wk, per_chunk, depth, &ring)) return fail_last(env);
  RingBox* box = new RingBox();
  box->ring = ring;
  box->owner = p->owner;
  box->owner->refs++;
  const size_t n = (size_t)pdsp_plan_size(p->plan);
  const bool f64 = pdsp_plan_precision(p->plan) == PDSP_F64;
  box->frame_len = (size_t)d.frame_len;
  box->bins = d.sides == PDSP_SIDES_TWO ? n : n / 2 + 1;
  box->peak_bytes = f64 ? sizeof(pdsp_peak_f64) : sizeof(pdsp_peak_f32);
  box->sample_type = d.sample_dtype == PDSP_F64 ? napi_float64_array : napi_float32_array;
  box->out_type = f64 ? napi_float64_array : napi_float32_array;
  box->want_amp = wa != 0, box->want_phase = wp != 0, box->want_peaks = wk != 0;
  napi_value ext;
  if (napi_create_external(env, box, ring_finalize, nullptr, &ext) != napi_ok) {
    ring_finalize(env, box, nullptr);
    return fail(env, "pragma-dsp/b200: napi_create_external failed");
  }
  return ext;
}
// ingestPush(ring, frames: typed array of the ring's sample type holding `count` whole frames, count) -> frames accepted
napi_value IngestPush(napi_env env, napi_callback_info info) {
  ARGS(3)
  RING_ARG(r, argv[0])
  TA s;
  int64_t count = 0, accepted = 0;
  napi_get_value_int64(env, argv[2], &count);
  if (!get_ta(env, argv[1], &s) || !s.present) return fail(env, "pragma-dsp/b200: expected (ring, frames, count)");
  if (s.type != r->sample_type) return fail(env, "pragma-dsp/b200: frames must have the ring's sample type");
  if (count < 0 || (size_t)count * r->frame_len > s.length) return fail(env, "pragma-dsp/b200: frames array shorter than count * frameLen");
  if (pdsp_ingest_push(r->ring, s.data, count, 0, &accepted)) return fail_last(env);
  napi_value v;
  napi_create_double(env, (double)accepted, &v);
  return v;
}
napi_value IngestFlush(napi_env env, napi_callback_info info) {
  ARGS(1)
  RING_ARG(r, argv[0])
  if (pdsp_ingest_flush(r->ring)) return fail_last(env);
  return undefined(env);
}
// ingestPop(ring, amplitude|null, phase|null, peaks: Uint8Array|null, maxFrames) -> frames returned
napi_value IngestPop(napi_env env, napi_callback_info info) {
  ARGS(5)
  RING_ARG(r, argv[0])
  TA amp, ph, pk;
  int64_t max_frames = 0, got = 0;
  napi_get_value_int64(env, argv[4], &max_frames);
  if (!get_ta(env, argv[1], &amp) || !get_ta(env, argv[2], &ph) || !get_ta(env, argv[3], &pk))
    return fail(env, "pragma-dsp/b200: expected (ring, amplitude|null, phase|null, peaks|null, maxFrames)");
  if (max_frames < 0) return fail(env, "pragma-dsp/b200: negative maxFrames");
  const size_t rows = (size_t)max_frames * r->bins;
  if ((amp.present && (amp.type != r->out_type || amp.length < rows)) || (ph.present && (ph.type != r->out_type || ph.length < rows)) ||
      (pk.present && (pk.type != napi_uint8_array || pk.length < (size_t)max_frames * r->peak_bytes)))
    return fail(env, "pragma-dsp/b200: output arrays have the wrong type or length for maxFrames");
  if (pdsp_ingest_pop(r->ring, amp.data, ph.data, pk.data, max_frames, &got)) return fail_last(env);
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
  def(env, exports, "applyWindow", ApplyWindow);
  def(env, exports, "fftShift", FftShift);
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
