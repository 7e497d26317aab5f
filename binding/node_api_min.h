/* Minimal hand-written subset of Node's stable N-API (node_api.h is not present in this image).
 * Types and prototypes follow the published ABI-stable N-API v8 headers; when building inside a real
 * Node.js toolchain, drop this file and `#include <node_api.h>` instead. */
#ifndef EB200_NODE_API_MIN_H
#define EB200_NODE_API_MIN_H
#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok = 0 } napi_status;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef enum { napi_uint8_array = 1 } napi_typedarray_type;
typedef struct {
  int nm_version; unsigned int nm_flags; const char* nm_filename;
  napi_value (*nm_register_func)(napi_env env, napi_value exports);
  const char* nm_modname; void* nm_priv; void* reserved[4];
} napi_module;
#ifdef __cplusplus
extern "C" {
#endif
void napi_module_register(napi_module* mod);
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data, napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t* result);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t* result);
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void** data, napi_value* result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset, napi_value* result);
napi_status napi_create_function(napi_env env, const char* utf8name, size_t length, napi_callback cb, void* data, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
napi_status napi_get_undefined(napi_env env, napi_value* result);
#ifdef __cplusplus
}
#endif
#endif
