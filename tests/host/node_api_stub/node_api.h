/* Minimal stand-in for Node's <node_api.h>: just the declarations compressjs_b200/napi/addon.cc uses, with the
 * signatures of N-API version 6, so that the CPU test-suite can at least type-check the addon (`g++ -fsyntax-only`)
 * in an image without node.  NOT a replacement for building against the real header. */
#ifndef NODE_API_STUB_H
#define NODE_API_STUB_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok, napi_invalid_arg, napi_generic_failure } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function, napi_external, napi_bigint } napi_valuetype;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array, napi_uint32_array,
               napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array } napi_typedarray_type;
typedef enum { napi_default = 0 } napi_property_attributes;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
typedef struct {
  const char* utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter; napi_value value;
  napi_property_attributes attributes; void* data;
} napi_property_descriptor;
#define NAPI_AUTO_LENGTH ((size_t)-1)
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_get_buffer_info(napi_env, napi_value, void** data, size_t* length);
napi_status napi_get_typedarray_info(napi_env, napi_value, napi_typedarray_type*, size_t* length, void** data, napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_typeof(napi_env, napi_value, napi_valuetype*);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t*);
napi_status napi_get_value_double(napi_env, napi_value, double*);
napi_status napi_get_value_bool(napi_env, napi_value, bool*);
napi_status napi_create_int32(napi_env, int32_t, napi_value*);
napi_status napi_create_uint32(napi_env, uint32_t, napi_value*);
napi_status napi_create_double(napi_env, double, napi_value*);
napi_status napi_create_string_utf8(napi_env, const char*, size_t, napi_value*);
napi_status napi_create_error(napi_env, napi_value code, napi_value msg, napi_value* result);
napi_status napi_create_type_error(napi_env, napi_value code, napi_value msg, napi_value* result);
napi_status napi_create_array_with_length(napi_env, size_t, napi_value*);
napi_status napi_create_external_buffer(napi_env, size_t length, void* data, napi_finalize, void* hint, napi_value*);
napi_status napi_set_element(napi_env, napi_value object, uint32_t index, napi_value value);
napi_status napi_set_named_property(napi_env, napi_value object, const char* name, napi_value value);
napi_status napi_define_properties(napi_env, napi_value object, size_t count, const napi_property_descriptor*);
napi_status napi_throw(napi_env, napi_value error);
napi_status napi_throw_error(napi_env, const char* code, const char* msg);
#define NAPI_MODULE(name, init) napi_value napi_stub_register_##name(napi_env env, napi_value exports) { return init(env, exports); }
#define NODE_GYP_MODULE_NAME b2bz
#ifdef __cplusplus
}
#endif
#endif
