// mock_napi.cc - a small Node-API HOST for tests (TEST INFRASTRUCTURE ONLY).
//
// This image has no Node.js, so napi/pragma_napi.cc could only ever be compiled.  This file implements the Node-API
// functions the addon uses (the subset declared in napi/node_api_min.h) over a toy value model - numbers, objects with
// named properties, functions, externals with finalisers, array buffers (plain and external) and typed arrays - loads
// the addon with dlopen like Node does, calls napi_register_module_v1, and lets a test driver (Python, ctypes) build
// argument values, call the exported functions, look at results and pending exceptions, and run finalisers in any
// order (garbage collection and environment teardown give no ordering guarantee).
//
// Loaded with RTLD_GLOBAL so that the addon's undefined napi_* symbols resolve against this library.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../napi/node_api_min.h"

namespace {
enum Kind { K_UNDEFINED, K_NULL, K_BOOL, K_NUMBER, K_STRING, K_OBJECT, K_FUNCTION, K_EXTERNAL, K_ARRAYBUFFER, K_TYPEDARRAY };
}

struct napi_value__ {
  Kind kind = K_UNDEFINED;
  double num = 0;
  std::string str;
  std::map<std::string, napi_value> props;
  napi_callback cb = nullptr;
  void* cb_data = nullptr;
  // external / external array buffer
  void* data = nullptr;
  size_t byte_length = 0;
  napi_finalize fin = nullptr;
  void* fin_hint = nullptr;
  bool finalized = false;
  std::vector<unsigned char> owned;  // plain ArrayBuffer storage
  // typed array
  napi_typedarray_type ta_type = napi_uint8_array;
  size_t ta_length = 0;
  size_t ta_offset = 0;
  napi_value ta_buffer = nullptr;
};

struct napi_env__ {
  std::vector<std::unique_ptr<napi_value__>> heap;
  bool pending = false;
  std::string message;
  void* addon = nullptr;
  long finalizers_run = 0;
  napi_value make(Kind k) {
    heap.emplace_back(new napi_value__());
    heap.back()->kind = k;
    return heap.back().get();
  }
};

struct napi_callback_info__ {
  size_t argc;
  napi_value* argv;
  napi_value this_arg;
  void* data;
};

static size_t elem_size(napi_typedarray_type t) {
  switch (t) {
    case napi_int8_array: case napi_uint8_array: case napi_uint8_clamped_array: return 1;
    case napi_int16_array: case napi_uint16_array: return 2;
    case napi_int32_array: case napi_uint32_array: case napi_float32_array: return 4;
    default: return 8;
  }
}

// ------------------------------------------------------------------------------------------ Node-API subset
extern "C" {
napi_status napi_get_cb_info(napi_env, napi_callback_info info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data) {
  if (argc) {
    const size_t cap = *argc;
    for (size_t i = 0; i < cap && argv; ++i) argv[i] = i < info->argc ? info->argv[i] : nullptr;
    *argc = info->argc;
  }
  if (this_arg) *this_arg = info->this_arg;
  if (data) *data = info->data;
  return napi_ok;
}
napi_status napi_typeof(napi_env, napi_value v, napi_valuetype* result) {
  if (!v || !result) return napi_invalid_arg;
  switch (v->kind) {
    case K_UNDEFINED: *result = napi_undefined; break;
    case K_NULL: *result = napi_null; break;
    case K_BOOL: *result = napi_boolean; break;
    case K_NUMBER: *result = napi_number; break;
    case K_STRING: *result = napi_string; break;
    case K_FUNCTION: *result = napi_function; break;
    case K_EXTERNAL: *result = napi_external; break;
    default: *result = napi_object; break;  // objects, array buffers, typed arrays
  }
  return napi_ok;
}
napi_status napi_get_value_double(napi_env, napi_value v, double* r) {
  if (!v || v->kind != K_NUMBER) return napi_invalid_arg;
  *r = v->num;
  return napi_ok;
}
napi_status napi_get_value_int32(napi_env, napi_value v, int32_t* r) {
  if (!v || v->kind != K_NUMBER) return napi_invalid_arg;
  *r = (int32_t)(int64_t)v->num;  // ToInt32 for the finite values the tests use
  return napi_ok;
}
napi_status napi_get_value_int64(napi_env, napi_value v, int64_t* r) {
  if (!v || v->kind != K_NUMBER) return napi_invalid_arg;
  *r = (int64_t)v->num;
  return napi_ok;
}
napi_status napi_get_value_external(napi_env, napi_value v, void** r) {
  if (!v || v->kind != K_EXTERNAL) return napi_invalid_arg;
  *r = v->data;
  return napi_ok;
}
napi_status napi_create_external(napi_env env, void* data, napi_finalize fin, void* hint, napi_value* result) {
  napi_value v = env->make(K_EXTERNAL);
  v->data = data, v->fin = fin, v->fin_hint = hint;
  *result = v;
  return napi_ok;
}
napi_status napi_create_double(napi_env env, double value, napi_value* result) {
  napi_value v = env->make(K_NUMBER);
  v->num = value;
  *result = v;
  return napi_ok;
}
napi_status napi_create_int32(napi_env env, int32_t value, napi_value* result) { return napi_create_double(env, (double)value, result); }
napi_status napi_get_undefined(napi_env env, napi_value* result) {
  *result = env->make(K_UNDEFINED);
  return napi_ok;
}
napi_status napi_is_typedarray(napi_env, napi_value v, bool* result) {
  *result = v && v->kind == K_TYPEDARRAY;
  return napi_ok;
}
napi_status napi_get_typedarray_info(napi_env, napi_value v, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset) {
  if (!v || v->kind != K_TYPEDARRAY) return napi_invalid_arg;
  if (type) *type = v->ta_type;
  if (length) *length = v->ta_length;
  if (data) *data = v->ta_buffer ? static_cast<char*>(v->ta_buffer->data) + v->ta_offset : v->data;
  if (arraybuffer) *arraybuffer = v->ta_buffer;
  if (byte_offset) *byte_offset = v->ta_offset;
  return napi_ok;
}
napi_status napi_create_external_arraybuffer(napi_env env, void* data, size_t len, napi_finalize fin, void* hint, napi_value* result) {
  napi_value v = env->make(K_ARRAYBUFFER);
  v->data = data, v->byte_length = len, v->fin = fin, v->fin_hint = hint;
  *result = v;
  return napi_ok;
}
napi_status napi_create_function(napi_env env, const char* name, size_t, napi_callback cb, void* data, napi_value* result) {
  napi_value v = env->make(K_FUNCTION);
  v->str = name ? name : "";
  v->cb = cb, v->cb_data = data;
  *result = v;
  return napi_ok;
}
napi_status napi_set_named_property(napi_env, napi_value obj, const char* name, napi_value value) {
  if (!obj || obj->kind != K_OBJECT) return napi_invalid_arg;
  obj->props[name] = value;
  return napi_ok;
}
napi_status napi_get_named_property(napi_env env, napi_value obj, const char* name, napi_value* result) {
  if (!obj || (obj->kind != K_OBJECT && obj->kind != K_TYPEDARRAY && obj->kind != K_ARRAYBUFFER)) return napi_invalid_arg;
  auto it = obj->props.find(name);
  *result = it == obj->props.end() ? env->make(K_UNDEFINED) : it->second;
  return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char*, const char* msg) {
  env->pending = true;
  env->message = msg ? msg : "";
  return napi_ok;
}

// ------------------------------------------------------------------------------------------ driver API (ctypes)
napi_env mock_env_create(void) { return new napi_env__(); }

// dlopen the addon and run its module registration; returns the exports object (NULL + mock_error on failure)
napi_value mock_load_addon(napi_env env, const char* path) {
  env->addon = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!env->addon) {
    env->pending = true;
    env->message = dlerror();
    return nullptr;
  }
  typedef napi_value (*init_fn)(napi_env, napi_value);
  init_fn init = reinterpret_cast<init_fn>(dlsym(env->addon, "napi_register_module_v1"));
  if (!init) {
    env->pending = true;
    env->message = "napi_register_module_v1 not exported";
    return nullptr;
  }
  return init(env, env->make(K_OBJECT));
}
napi_value mock_number(napi_env env, double x) {
  napi_value v;
  napi_create_double(env, x, &v);
  return v;
}
napi_value mock_null(napi_env env) { return env->make(K_NULL); }
napi_value mock_undefined(napi_env env) { return env->make(K_UNDEFINED); }
napi_value mock_object(napi_env env) { return env->make(K_OBJECT); }
void mock_object_set(napi_env, napi_value obj, const char* key, napi_value v) { obj->props[key] = v; }
// a typed array over memory the driver owns (a numpy array): new Float64Array(buffer, 0, length)
napi_value mock_typedarray(napi_env env, int type, void* data, size_t length) {
  napi_value v = env->make(K_TYPEDARRAY);
  v->ta_type = (napi_typedarray_type)type, v->ta_length = length, v->data = data;
  return v;
}
// new Float64Array(arraybuffer, byteOffset, length) over an (external) ArrayBuffer value
napi_value mock_typedarray_on(napi_env env, napi_value ab, int type, size_t byte_offset, size_t length) {
  if (!ab || ab->kind != K_ARRAYBUFFER || byte_offset + length * elem_size((napi_typedarray_type)type) > ab->byte_length) return nullptr;
  napi_value v = env->make(K_TYPEDARRAY);
  v->ta_type = (napi_typedarray_type)type, v->ta_length = length, v->ta_buffer = ab, v->ta_offset = byte_offset;
  return v;
}
int mock_kind(napi_value v) { return v ? (int)v->kind : -1; }
double mock_get_number(napi_value v) { return v && v->kind == K_NUMBER ? v->num : 0.0; }
void* mock_buffer_data(napi_value v) { return v ? v->data : nullptr; }
size_t mock_buffer_length(napi_value v) { return v ? v->byte_length : 0; }
int mock_has_function(napi_value exports, const char* name) {
  auto it = exports->props.find(name);
  return it != exports->props.end() && it->second->kind == K_FUNCTION;
}
// number of exported functions; names are written '\n'-separated into buf
int mock_export_names(napi_value exports, char* buf, size_t cap) {
  std::string s;
  int n = 0;
  for (auto& kv : exports->props)
    if (kv.second->kind == K_FUNCTION) {
      s += kv.first + "\n";
      ++n;
    }
  snprintf(buf, cap, "%s", s.c_str());
  return n;
}
// exports[name](...argv); NULL when the call left an exception pending (mock_error returns and clears it)
napi_value mock_call(napi_env env, napi_value exports, const char* name, size_t argc, napi_value* argv) {
  auto it = exports->props.find(name);
  if (it == exports->props.end() || it->second->kind != K_FUNCTION) {
    env->pending = true;
    env->message = std::string("TypeError: exports.") + name + " is not a function";
    return nullptr;
  }
  napi_callback_info__ info{argc, argv, exports, it->second->cb_data};
  napi_value r = it->second->cb(env, &info);
  if (env->pending) return nullptr;
  return r ? r : env->make(K_UNDEFINED);
}
const char* mock_error(napi_env env) {
  static thread_local std::string last;
  if (!env->pending) return nullptr;
  last = env->message;
  env->pending = false;
  env->message.clear();
  return last.c_str();
}
// the garbage collector found `v` unreachable: run its finaliser (once)
int mock_collect(napi_env env, napi_value v) {
  if (!v || v->finalized || !v->fin) return 0;
  v->finalized = true;
  v->fin(env, v->data, v->fin_hint);
  env->finalizers_run++;
  return 1;
}
long mock_finalizers_run(napi_env env) { return env->finalizers_run; }
// environment teardown: every remaining finaliser, in creation order (order = 0) or reverse (order = 1)
void mock_env_destroy(napi_env env, int order) {
  const size_t n = env->heap.size();
  for (size_t i = 0; i < n; ++i) mock_collect(env, env->heap[order ? n - 1 - i : i].get());
  // the addon stays loaded (as in Node, addons are not unloaded at teardown)
  delete env;
}
}  // extern "C"
