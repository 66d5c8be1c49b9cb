/* DECLARATION-ONLY STAND-IN for Node's <node_api.h>, limited to what js/addon.cc uses.
 *
 * Node.js (and its headers) are absent from this image, so the addon cannot be built here.  This file exists so that
 * `g++ -fsyntax-only -I tools/napi_stub js/addon.cc` (tests/test_capi_cpu.py) at least type-checks the addon against
 * the C ABI header and against the Node-API signatures as documented (nodejs.org/api/n-api.html, N-API version 8).
 * It is not a Node-API implementation and is never linked or shipped. */
#ifndef MB_NAPI_STUB_H
#define MB_NAPI_STUB_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_ref__ *napi_ref;
typedef struct napi_callback_info__ *napi_callback_info;
typedef struct napi_deferred__ *napi_deferred;
typedef struct napi_async_work__ *napi_async_work;
typedef enum { napi_ok, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_name_expected,
               napi_function_expected, napi_number_expected, napi_boolean_expected, napi_array_expected,
               napi_generic_failure, napi_pending_exception, napi_cancelled } napi_status;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
               napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array,
               napi_biguint64_array } napi_typedarray_type;
typedef enum { napi_default = 0, napi_writable = 1 << 0, napi_enumerable = 1 << 1, napi_configurable = 1 << 2 } napi_property_attributes;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);
typedef void (*napi_async_execute_callback)(napi_env env, void *data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void *data);
typedef struct {
    const char *utf8name;
    napi_value name;
    napi_callback method;
    napi_callback getter;
    napi_callback setter;
    napi_value value;
    napi_property_attributes attributes;
    void *data;
} napi_property_descriptor;
#define NAPI_AUTO_LENGTH SIZE_MAX

napi_status napi_throw_error(napi_env env, const char *code, const char *msg);
napi_status napi_throw_type_error(napi_env env, const char *code, const char *msg);
napi_status napi_has_named_property(napi_env env, napi_value object, const char *utf8name, bool *result);
napi_status napi_get_named_property(napi_env env, napi_value object, const char *utf8name, napi_value *result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char *utf8name, napi_value value);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t *result);
napi_status napi_get_value_double(napi_env env, napi_value value, double *result);
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t *argc, napi_value *argv, napi_value *this_arg, void **data);
napi_status napi_create_external(napi_env env, void *data, napi_finalize finalize_cb, void *finalize_hint, napi_value *result);
napi_status napi_get_value_external(napi_env env, napi_value value, void **result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type *type, size_t *length, void **data,
                                     napi_value *arraybuffer, size_t *byte_offset);
napi_status napi_create_object(napi_env env, napi_value *result);
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void **data, napi_value *result);
napi_status napi_create_external_arraybuffer(napi_env env, void *external_data, size_t byte_length, napi_finalize finalize_cb,
                                             void *finalize_hint, napi_value *result);
napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t *result);
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value *result);
napi_status napi_create_double(napi_env env, double value, napi_value *result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset,
                                   napi_value *result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value *result);
napi_status napi_create_int64(napi_env env, int64_t value, napi_value *result);
napi_status napi_create_string_utf8(napi_env env, const char *str, size_t length, napi_value *result);
napi_status napi_create_error(napi_env env, napi_value code, napi_value msg, napi_value *result);
napi_status napi_create_array_with_length(napi_env env, size_t length, napi_value *result);
napi_status napi_set_element(napi_env env, napi_value object, uint32_t index, napi_value value);
napi_status napi_create_reference(napi_env env, napi_value value, uint32_t initial_refcount, napi_ref *result);
napi_status napi_delete_reference(napi_env env, napi_ref ref);
napi_status napi_get_reference_value(napi_env env, napi_ref ref, napi_value *result);
napi_status napi_create_promise(napi_env env, napi_deferred *deferred, napi_value *promise);
napi_status napi_resolve_deferred(napi_env env, napi_deferred deferred, napi_value resolution);
napi_status napi_reject_deferred(napi_env env, napi_deferred deferred, napi_value rejection);
napi_status napi_create_async_work(napi_env env, napi_value async_resource, napi_value async_resource_name,
                                   napi_async_execute_callback execute, napi_async_complete_callback complete, void *data,
                                   napi_async_work *result);
napi_status napi_queue_async_work(napi_env env, napi_async_work work);
napi_status napi_delete_async_work(napi_env env, napi_async_work work);
napi_status napi_remove_wrap(napi_env env, napi_value js_object, void **result);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count, const napi_property_descriptor *properties);
#ifdef __cplusplus
}
#endif
#define NODE_GYP_MODULE_NAME meyda_b200
#define NAPI_MODULE(modname, regfunc) \
    extern "C" napi_value napi_register_module_v1(napi_env env, napi_value exports) { return regfunc(env, exports); }
#endif
