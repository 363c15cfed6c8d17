// tendrils_b200_napi.cc -- Node N-API addon over the C ABI of include/tendrils_b200.h.
//
// NOT BUILT IN THIS REPOSITORY'S IMAGE (no Node, no node_api.h).  It is the binding a maintainer of
// keeffEoghan/tendrils adds next to src/index.js; see INTEGRATION.md.  Build (where Node exists):
//   node-gyp configure build     (binding.gyp links -ltendrils_b200)
//
// Every function forwards to exactly one tb_* symbol; a non-zero status becomes a thrown Error
// carrying tb_last_error(), which is how the reference's stack.gl dependencies report failures.
#include <node_api.h>

#include <cstdint>
#include <cstring>

#include "../../include/tendrils_b200.h"

namespace {

#define NAPI_OK(call)                                                           \
    do {                                                                        \
        if ((call) != napi_ok) {                                                \
            napi_throw_error(env, nullptr, "tendrils-b200: N-API call failed"); \
            return nullptr;                                                     \
        }                                                                       \
    } while (0)

napi_value check(napi_env env, tb_ctx *ctx, int status) {
    if (status != TB_OK) napi_throw_error(env, nullptr, tb_last_error(ctx));
    return nullptr;
}

tb_ctx *unwrap(napi_env env, napi_value v) {
    void *p = nullptr;
    napi_get_value_external(env, v, &p);
    return static_cast<tb_ctx *>(p);
}

double num(napi_env env, napi_value v) {
    double d = 0;
    napi_get_value_double(env, v, &d);
    return d;
}

double prop(napi_env env, napi_value obj, const char *key, double fallback = 0) {
    napi_value v;
    bool has = false;
    napi_has_named_property(env, obj, key, &has);
    if (!has) return fallback;
    napi_get_named_property(env, obj, key, &v);
    return num(env, v);
}

void finalize(napi_env, void *data, void *) { tb_destroy(static_cast<tb_ctx *>(data)); }

// create({particlesW, particlesH, col0, col1, flowW, flowH, device}) -> external
napi_value Create(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_config cfg{};
    cfg.particles_w = static_cast<int32_t>(prop(env, argv[0], "particlesW"));
    cfg.particles_h = static_cast<int32_t>(prop(env, argv[0], "particlesH"));
    cfg.col0 = static_cast<int32_t>(prop(env, argv[0], "col0"));
    cfg.col1 = static_cast<int32_t>(prop(env, argv[0], "col1"));
    cfg.flow_w = static_cast<int32_t>(prop(env, argv[0], "flowW", 1));
    cfg.flow_h = static_cast<int32_t>(prop(env, argv[0], "flowH", 1));
    cfg.device = static_cast<int32_t>(prop(env, argv[0], "device"));
    tb_ctx *ctx = nullptr;
    if (int s = tb_create(&cfg, &ctx)) return check(env, nullptr, s);
    napi_value out;
    NAPI_OK(napi_create_external(env, ctx, finalize, nullptr, &out));
    return out;
}

// setState(ctx, state, viewSize): the state object of src/index.js:29-57, doubles -> float as gl.uniform1f does
napi_value SetState(napi_env env, napi_callback_info info) {
    size_t argc = 3;
    napi_value argv[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    tb_state s{};
#define F(name) s.name = static_cast<float>(prop(env, argv[1], #name))
    F(damping); F(speedLimit); F(forceWeight); F(varyForce); F(flowWeight); F(varyFlow); F(noiseWeight); F(varyNoise);
    F(flowDecay); F(flowWidth); F(noiseScale); F(varyNoiseScale); F(noiseSpeed); F(varyNoiseSpeed); F(target); F(varyTarget);
#undef F
    for (uint32_t i = 0; i < 2; ++i) {
        napi_value e;
        napi_get_element(env, argv[2], i, &e);
        s.viewSize[i] = static_cast<float>(num(env, e));
    }
    return check(env, ctx, tb_set_state(ctx, &s));
}

napi_value Step(napi_env env, napi_callback_info info) {          // step(ctx, time, dt)
    size_t argc = 3;
    napi_value argv[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    return check(env, ctx, tb_step(ctx, static_cast<float>(num(env, argv[1])), static_cast<float>(num(env, argv[2]))));
}

napi_value SplatFlow(napi_env env, napi_callback_info info) {     // splatFlow(ctx, time)
    size_t argc = 2;
    napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    return check(env, ctx, tb_splat_flow(ctx, static_cast<float>(num(env, argv[1]))));
}

napi_value ResizeFlow(napi_env env, napi_callback_info info) {    // resizeFlow(ctx, w, h)
    size_t argc = 3;
    napi_value argv[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    return check(env, ctx, tb_resize_flow(ctx, static_cast<int32_t>(num(env, argv[1])), static_cast<int32_t>(num(env, argv[2]))));
}

napi_value ClearFlow(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    return check(env, ctx, tb_clear_flow(ctx));
}

napi_value Reset(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    return check(env, ctx, tb_reset(ctx));
}

napi_value SpawnInit(napi_env env, napi_callback_info info) {     // spawnInit(ctx, target)
    size_t argc = 2;
    napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    return check(env, ctx, tb_spawn_init(ctx, static_cast<tb_target>(static_cast<int>(num(env, argv[1])))));
}

napi_value SpawnBall(napi_env env, napi_callback_info info) {     // spawnBall(ctx, radius, speed, target)
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    return check(env, ctx, tb_spawn_ball(ctx, static_cast<float>(num(env, argv[1])), static_cast<float>(num(env, argv[2])),
                                         static_cast<tb_target>(static_cast<int>(num(env, argv[3])))));
}

napi_value SetSpawnImage(napi_env env, napi_callback_info info) { // setSpawnImage(ctx, Float32Array rgba, w, h)
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    napi_typedarray_type type;
    size_t len = 0;
    void *data = nullptr;
    NAPI_OK(napi_get_typedarray_info(env, argv[1], &type, &len, &data, nullptr, nullptr));
    const int32_t w = static_cast<int32_t>(num(env, argv[2])), h = static_cast<int32_t>(num(env, argv[3]));
    if (type != napi_float32_array || len != static_cast<size_t>(w) * h * 4) {
        napi_throw_error(env, nullptr, "tendrils-b200: spawn image must be a Float32Array of w*h*4");
        return nullptr;
    }
    return check(env, ctx, tb_set_spawn_image(ctx, static_cast<const float *>(data), w, h));
}

// spawnPixels(ctx, {spawnSize, jitter, speed, bias, spawnMatrix}, variant, source, time, target)
napi_value SpawnPixels(napi_env env, napi_callback_info info) {
    size_t argc = 6;
    napi_value argv[6];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    tb_pixel_spawner p{};
    auto vec = [&](const char *key, float *dst, uint32_t n) {
        napi_value arr, e;
        napi_get_named_property(env, argv[1], key, &arr);
        for (uint32_t i = 0; i < n; ++i) {
            napi_get_element(env, arr, i, &e);
            dst[i] = static_cast<float>(num(env, e));
        }
    };
    vec("spawnSize", p.spawnSize, 2);
    vec("jitter", p.jitter, 2);
    vec("spawnMatrix", p.spawnMatrix, 9);
    p.speed = static_cast<float>(prop(env, argv[1], "speed", 1));
    p.bias = static_cast<float>(prop(env, argv[1], "bias", 1));
    return check(env, ctx, tb_spawn_pixels(ctx, &p, static_cast<tb_spawn_variant>(static_cast<int>(num(env, argv[2]))),
                                           static_cast<tb_spawn_source>(static_cast<int>(num(env, argv[3]))),
                                           static_cast<float>(num(env, argv[4])),
                                           static_cast<tb_target>(static_cast<int>(num(env, argv[5])))));
}

// upload(ctx, which, Float32Array) / download(ctx, which, Float32Array)
napi_value Transfer(napi_env env, napi_callback_info info, bool up) {
    size_t argc = 3;
    napi_value argv[3];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    napi_typedarray_type type;
    size_t len = 0;
    void *data = nullptr;
    NAPI_OK(napi_get_typedarray_info(env, argv[2], &type, &len, &data, nullptr, nullptr));
    if (type != napi_float32_array) {
        napi_throw_error(env, nullptr, "tendrils-b200: expected a Float32Array");
        return nullptr;
    }
    const tb_buffer which = static_cast<tb_buffer>(static_cast<int>(num(env, argv[1])));
    return check(env, ctx, up ? tb_upload(ctx, which, static_cast<const float *>(data), static_cast<int64_t>(len))
                              : tb_download(ctx, which, static_cast<float *>(data), static_cast<int64_t>(len)));
}
napi_value Upload(napi_env env, napi_callback_info info) { return Transfer(env, info, true); }
napi_value Download(napi_env env, napi_callback_info info) { return Transfer(env, info, false); }

// a typed array argument of the expected element type; returns its data pointer and length, or throws
void *typed(napi_env env, napi_value v, napi_typedarray_type want, size_t *len) {
    napi_typedarray_type type;
    void *data = nullptr;
    *len = 0;
    if (napi_get_typedarray_info(env, v, &type, len, &data, nullptr, nullptr) != napi_ok || type != want) {
        napi_throw_error(env, nullptr, "tendrils-b200: typed array of the wrong element type");
        return nullptr;
    }
    return data;
}

void vec2(napi_env env, napi_value obj, const char *key, float *dst) {
    napi_value arr, e;
    napi_get_named_property(env, obj, key, &arr);
    for (uint32_t i = 0; i < 2; ++i) {
        napi_get_element(env, arr, i, &e);
        dst[i] = static_cast<float>(num(env, e));
    }
}

// opticalFlow(ctx, {viewSize, scaleUV, offset, lambda, speed, speedLimit, time}, Uint8Array view, Uint8Array last, w, h)
// = OpticalFlow.update() + screen.render() with the flow FBO bound (src/demo.main.js:1131-1156)
napi_value OpticalFlow(napi_env env, napi_callback_info info) {
    size_t argc = 6;
    napi_value argv[6];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    tb_optical_flow_params p{};
    vec2(env, argv[1], "viewSize", p.viewSize);
    vec2(env, argv[1], "scaleUV", p.scaleUV);
    p.offset = static_cast<float>(prop(env, argv[1], "offset", 1));
    p.lambda = static_cast<float>(prop(env, argv[1], "lambda", 0.001));
    p.speed = static_cast<float>(prop(env, argv[1], "speed", 1));
    p.speedLimit = static_cast<float>(prop(env, argv[1], "speedLimit", 1));
    p.time = static_cast<float>(prop(env, argv[1], "time", 1));
    const int32_t w = static_cast<int32_t>(num(env, argv[4])), h = static_cast<int32_t>(num(env, argv[5]));
    size_t n_view = 0, n_last = 0;
    const void *view = typed(env, argv[2], napi_uint8_array, &n_view);
    const void *last = typed(env, argv[3], napi_uint8_array, &n_last);
    if (!view || !last) return nullptr;
    if (n_view != static_cast<size_t>(w) * h * 4 || n_last != n_view) {
        napi_throw_error(env, nullptr, "tendrils-b200: frames must be Uint8Arrays of w*h*4");
        return nullptr;
    }
    return check(env, ctx, tb_optical_flow(ctx, &p, static_cast<const uint8_t *>(view), static_cast<const uint8_t *>(last), w, h));
}

// flowLine(ctx, {viewSize, rad, speed, speedLimit, crestShape}, position, normal, miter, previous, time, dt)
// = FlowLine.draw() with the flow FBO bound (src/demo.main.js:1107-1121); the arrays are line.attributes.*.data
napi_value FlowLine(napi_env env, napi_callback_info info) {
    size_t argc = 8;
    napi_value argv[8];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    tb_flow_line_params p{};
    vec2(env, argv[1], "viewSize", p.viewSize);
    p.rad = static_cast<float>(prop(env, argv[1], "rad", 0.1));
    p.speed = static_cast<float>(prop(env, argv[1], "speed", 3));
    p.speedLimit = static_cast<float>(prop(env, argv[1], "speedLimit", 0.01));
    p.crestShape = static_cast<float>(prop(env, argv[1], "crestShape", 0.6));
    size_t len[6] = {};
    const float *a[6] = {};
    for (int i = 0; i < 6; ++i) {
        a[i] = static_cast<const float *>(typed(env, argv[2 + i], napi_float32_array, &len[i]));
        if (!a[i] && len[i]) return nullptr;
    }
    const size_t n = len[2];                                   // miter: one float per vertex
    if (len[0] != 2 * n || len[1] != 2 * n || len[3] != 2 * n || len[4] != n || len[5] != n) {
        napi_throw_error(env, nullptr, "tendrils-b200: flow line attribute arrays disagree in length");
        return nullptr;
    }
    return check(env, ctx, tb_flow_line(ctx, &p, static_cast<int32_t>(n), a[0], a[1], a[2], a[3], a[4], a[5]));
}

// blendIntoFlow(ctx, Float32Array rgba, w, h): any other layer the application draws into the flow FBO
napi_value BlendIntoFlow(napi_env env, napi_callback_info info) {
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    size_t len = 0;
    const void *data = typed(env, argv[1], napi_float32_array, &len);
    if (!data) return nullptr;
    const int32_t w = static_cast<int32_t>(num(env, argv[2])), h = static_cast<int32_t>(num(env, argv[3]));
    if (len != static_cast<size_t>(w) * h * 4) {
        napi_throw_error(env, nullptr, "tendrils-b200: layer must be a Float32Array of w*h*4");
        return nullptr;
    }
    return check(env, ctx, tb_blend_into_flow(ctx, static_cast<const float *>(data), w, h));
}

// stepStreamed(ctx, time, dt, Float32Array hostIn, Float32Array hostOut, chunks): tb_step for callers that keep the state on
// the host (Particles.pixels): upload, logic pass and download as one pipelined pass.  ASYNCHRONOUS -- call sync(ctx) before
// reading hostOut; the arrays must stay alive (and should be pinned / not moved by the GC: use external ArrayBuffers).
napi_value StepStreamed(napi_env env, napi_callback_info info) {
    size_t argc = 6;
    napi_value argv[6];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    size_t n_in = 0, n_out = 0;
    const float *in = static_cast<const float *>(typed(env, argv[3], napi_float32_array, &n_in));
    float *out = static_cast<float *>(typed(env, argv[4], napi_float32_array, &n_out));
    if (!in || !out) return nullptr;
    if (n_in != n_out) {
        napi_throw_error(env, nullptr, "tendrils-b200: stepStreamed needs two state arrays of the same length");
        return nullptr;
    }
    return check(env, ctx, tb_step_streamed(ctx, static_cast<float>(num(env, argv[1])), static_cast<float>(num(env, argv[2])), in, out,
                                            static_cast<int32_t>(num(env, argv[5]))));
}

napi_value Sync(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    return check(env, ctx, tb_sync(ctx));
}

napi_value SetOverlap(napi_env env, napi_callback_info info) {   // setOverlap(ctx, on)
    size_t argc = 2;
    napi_value argv[2];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    return check(env, ctx, tb_set_overlap(ctx, static_cast<int32_t>(num(env, argv[1]))));
}

// stats(ctx) -> { kernelLaunches, lastFragments }
napi_value Stats(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
    tb_ctx *ctx = unwrap(env, argv[0]);
    int64_t launches = 0, frags = 0;
    if (int st = tb_stats(ctx, &launches, &frags)) return check(env, ctx, st);
    napi_value out, a, b;
    NAPI_OK(napi_create_object(env, &out));
    NAPI_OK(napi_create_double(env, static_cast<double>(launches), &a));
    NAPI_OK(napi_create_double(env, static_cast<double>(frags), &b));
    NAPI_OK(napi_set_named_property(env, out, "kernelLaunches", a));
    NAPI_OK(napi_set_named_property(env, out, "lastFragments", b));
    return out;
}

napi_value Init(napi_env env, napi_value exports) {
    const napi_property_descriptor props[] = {
        {"create", nullptr, Create, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"setState", nullptr, SetState, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"step", nullptr, Step, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"splatFlow", nullptr, SplatFlow, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"resizeFlow", nullptr, ResizeFlow, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"clearFlow", nullptr, ClearFlow, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"reset", nullptr, Reset, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"spawnInit", nullptr, SpawnInit, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"spawnBall", nullptr, SpawnBall, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"setSpawnImage", nullptr, SetSpawnImage, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"spawnPixels", nullptr, SpawnPixels, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"upload", nullptr, Upload, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"download", nullptr, Download, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"opticalFlow", nullptr, OpticalFlow, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"flowLine", nullptr, FlowLine, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"blendIntoFlow", nullptr, BlendIntoFlow, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"stepStreamed", nullptr, StepStreamed, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"sync", nullptr, Sync, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"setOverlap", nullptr, SetOverlap, nullptr, nullptr, nullptr, napi_default, nullptr},
        {"stats", nullptr, Stats, nullptr, nullptr, nullptr, napi_default, nullptr},
    };
    napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
    return exports;
}

}  // namespace

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
