/*
 * node_api_min.h - minimal declarations of the Node-API (N-API) C ABI used by pragma_napi.cc.
 *
 * This image has no Node.js and no node_api.h; Node-API is a stable C ABI, so the handful of
 * opaque types, enum values and function prototypes the shim needs are declared here by hand
 * (names, signatures and enum values as documented for Node-API version 8).  When building
 * against a real Node installation, include <node_api.h> instead (-DPDSP_HAVE_NODE_API_H).
 * Every napi_* symbol is left undefined in the addon and resolved from the host process at
 * dlopen time, which is how Node addons are linked.
 */
#ifndef PDSP_NODE_API_MIN_H
#define PDSP_NODE_API_MIN_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_ref__* napi_ref;
typedef struct napi_callback_info__* napi_callback_info;

typedef enum { napi_ok = 0, napi_invalid_arg = 1, napi_pending_exception = 10 } napi_status;
typedef enum {
  napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function,
  napi_external, napi_bigint
} napi_valuetype;
typedef enum {
  napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
  napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array,
  napi_biguint64_array
} napi_typedarray_type;

typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);

napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv,
                             napi_value* this_arg, void** data);
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype* result);
napi_status napi_get_value_double(napi_env env, napi_value value, double* result);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t* result);
napi_status napi_get_value_int64(napi_env env, napi_value value, int64_t* result);
napi_status napi_get_value_external(napi_env env, napi_value value, void** result);
napi_status napi_create_external(napi_env env, void* data, napi_finalize finalize_cb, void* finalize_hint,
                                 napi_value* result);
napi_status napi_create_double(napi_env env, double value, napi_value* result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value* result);
napi_status napi_get_undefined(napi_env env, napi_value* result);
napi_status napi_is_typedarray(napi_env env, napi_value value, bool* result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length,
                                     void** data, napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_create_external_arraybuffer(napi_env env, void* external_data, size_t byte_length,
                                             napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_create_function(napi_env env, const char* utf8name, size_t length, napi_callback cb, void* data,
                                 napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_get_named_property(napi_env env, napi_value object, const char* utf8name, napi_value* result);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);

#ifdef __cplusplus
}
#endif
#endif
