/* TEST STUB -- just enough of Node's <node_api.h> for `g++ -fsyntax-only` over bindings/node/tendrils_b200_napi.cc,
 * so that the shim cannot drift from include/tendrils_b200.h unnoticed (tests/test_host.py).  Node is not in this
 * image; declarations follow the documented N-API signatures. */
#ifndef TB_TEST_NODE_API_STUB_H
#define TB_TEST_NODE_API_STUB_H
#include <stddef.h>
#include <stdint.h>
typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_callback_info__ *napi_callback_info;
typedef enum { napi_ok = 0, napi_generic_failure = 9 } napi_status;
typedef enum {
    napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array,
    napi_uint32_array, napi_float32_array, napi_float64_array
} napi_typedarray_type;
typedef enum { napi_default = 0 } napi_property_attributes;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);
typedef struct {
    const char *utf8name;
    napi_value name;
    napi_callback method, getter, setter;
    napi_value value;
    napi_property_attributes attributes;
    void *data;
} napi_property_descriptor;
#ifdef __cplusplus
extern "C" {
#endif
napi_status napi_throw_error(napi_env env, const char *code, const char *msg);
napi_status napi_get_value_external(napi_env env, napi_value value, void **result);
napi_status napi_get_value_double(napi_env env, napi_value value, double *result);
napi_status napi_has_named_property(napi_env env, napi_value object, const char *utf8name, bool *result);
napi_status napi_get_named_property(napi_env env, napi_value object, const char *utf8name, napi_value *result);
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t *argc, napi_value *argv, napi_value *this_arg, void **data);
napi_status napi_create_external(napi_env env, void *data, napi_finalize finalize_cb, void *finalize_hint, napi_value *result);
napi_status napi_create_object(napi_env env, napi_value *result);
napi_status napi_create_double(napi_env env, double value, napi_value *result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char *utf8name, napi_value value);
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value *result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type *type, size_t *length, void **data,
                                     napi_value *arraybuffer, size_t *byte_offset);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count, const napi_property_descriptor *properties);
#ifdef __cplusplus
}
#endif
#define NAPI_MODULE(modname, regfunc) napi_value tb_test_napi_register(napi_env env, napi_value exports) { return regfunc(env, exports); }
#endif
