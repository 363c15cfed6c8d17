// tendrils-b200 -- pointer flow lines drawn into the flow grid.
//
// Reference: src/flow-line/index.vert:21-37 and src/flow-line/index.frag:10-17 run over the TRIANGLE_STRIP that
// src/geom/line/index.js:73-117 builds (two vertices per path point); drawn with the flow FBO bound and the
// over-blend on (src/demo.main.js:1107-1121, src/index.js:267-268).  Fixed function per spec/PARITY.md FL3-FL6:
// window positions snapped to 1/256 pixel, exact integer edge functions, top-left rule, affine interpolation from
// the edge functions, alpha-over blend in triangle order.
//
// The per-vertex and per-fragment functions below are __host__ __device__ and use plain binary32 operators:
// the library is built with -fmad=false (tendrils_b200/build.py), so nvcc contracts nothing, and `/`, sqrtf and
// the double division are IEEE by default (-prec-div / -prec-sqrt).  tests/test_flow_line.py also compiles this
// header with g++ -ffp-contract=off to cross-check the logic against the test suite's CPU restatement on machines
// without a GPU; the product itself only ever runs the CUDA kernels at the end of this file.
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define TB_HD __host__ __device__ __forceinline__
#else
#define TB_HD inline
#endif

namespace tb {
namespace fl {

struct Uniforms {                 // tb_flow_line_params
    float vsx, vsy;               // viewSize
    float rad, speed, speedLimit, crestShape;
};

struct Vertex {
    long long x, y;               // window position in 1/256 pixel (FL3)
    float v[7];                   // varyings: values.rgba, crest.xy, sdf
    int ok;                       // finite and inside the guard band; otherwise its triangles are culled
};

TB_HD float gl_min(float x, float y) { return (y < x) ? y : x; }
TB_HD float gl_max(float x, float y) { return (x < y) ? y : x; }
TB_HD float gl_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
TB_HD float gl_length(float x, float y) { return sqrtf(x * x + y * y); }
TB_HD float gl_mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }

TB_HD long long snap(float w) {
#if defined(__CUDA_ARCH__)
    return __float2ll_rn(w * 256.0f);
#else
    return llrintf(w * 256.0f);
#endif
}

// src/flow-line/index.vert:21-37, then the viewport transform and the snap
TB_HD void vertex_stage(const Uniforms &U, float px, float py, float nx, float ny, float miter, float prx, float pry,
                        float time, float dt, int W, int H, Vertex &out) {
    const float rate = U.speed / gl_max(dt, 1.0f);
    const float velx = (px - prx) * rate, vely = (py - pry) * rate;
    const float alpha = gl_min(gl_length(velx, vely) / U.speedLimit, 1.0f);     // flow(vel, speedLimit).a
    const float rad = U.rad * alpha;
    const float vx = px + (nx * rad) * miter, vy = py + (ny * rad) * miter;     // expand(position, normal, rad*values.a, miter)
    const float hw = static_cast<float>(W) / 2.0f, hh = static_cast<float>(H) / 2.0f;
    const float xw = (vx * U.vsx) * hw + hw, yw = (vy * U.vsy) * hh + hh;
    out.v[0] = velx; out.v[1] = vely; out.v[2] = time; out.v[3] = alpha;
    out.v[4] = nx * miter; out.v[5] = ny * miter;                               // crest
    out.v[6] = gl_sign(miter);                                                  // sdf
    // written so that NaN fails: a non-finite or far-away vertex culls its triangles
    out.ok = (fabsf(xw) < 262144.0f && fabsf(yw) < 262144.0f) ? 1 : 0;
    out.x = out.ok ? snap(xw) : 0;
    out.y = out.ok ? snap(yw) : 0;
}

TB_HD long long orient(long long ax, long long ay, long long bx, long long by, long long cx, long long cy) {
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
}
// a centre exactly on the edge a->b of a counter-clockwise triangle belongs to it iff the edge runs down, or is
// horizontal and runs towards -x (FL4)
TB_HD bool tie(long long ax, long long ay, long long bx, long long by) {
    const long long dx = bx - ax, dy = by - ay;
    return dy < 0 || (dy == 0 && dx < 0);
}

// One triangle at one pixel: coverage, varying interpolation, src/flow-line/index.frag:10-17.
// Returns false when the pixel centre is not covered.
TB_HD bool shade(const Vertex &a, const Vertex &b, const Vertex &c, int px, int py, float crestShape, float rgba[4]) {
    if (!(a.ok && b.ok && c.ok)) return false;
    const long long area = orient(a.x, a.y, b.x, b.y, c.x, c.y);
    if (area == 0) return false;
    const long long cx = static_cast<long long>(px) * 256 + 128, cy = static_cast<long long>(py) * 256 + 128;
    const long long e0 = orient(b.x, b.y, c.x, c.y, cx, cy), e1 = orient(c.x, c.y, a.x, a.y, cx, cy),
                    e2 = orient(a.x, a.y, b.x, b.y, cx, cy);
    const bool ccw = area > 0;
    const long long s0 = ccw ? e0 : -e0, s1 = ccw ? e1 : -e1, s2 = ccw ? e2 : -e2;
    if (s0 < 0 || s1 < 0 || s2 < 0) return false;
    // with area < 0 every edge of the counter-clockwise orientation is walked backwards
    if (s0 == 0 && !(ccw ? tie(b.x, b.y, c.x, c.y) : tie(c.x, c.y, b.x, b.y))) return false;
    if (s1 == 0 && !(ccw ? tie(c.x, c.y, a.x, a.y) : tie(a.x, a.y, c.x, c.y))) return false;
    if (s2 == 0 && !(ccw ? tie(a.x, a.y, b.x, b.y) : tie(b.x, b.y, a.x, a.y))) return false;
    const float b1 = static_cast<float>(static_cast<double>(e1) / static_cast<double>(area));      // FL5
    const float b2 = static_cast<float>(static_cast<double>(e2) / static_cast<double>(area));
    float in[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) in[k] = (a.v[k] + b1 * (b.v[k] - a.v[k])) + b2 * (c.v[k] - a.v[k]);
    const float d = fabsf(in[6]);
    const float speed = gl_length(in[0], in[1]) * (1.0f - d);
    const float t = d * crestShape;
    const float mx = gl_mix(in[0], in[4], t), my = gl_mix(in[1], in[5], t);
    const float len = gl_length(mx, my);
    rgba[0] = (mx / len) * speed;                                               // normalize(m)*speed
    rgba[1] = (my / len) * speed;
    rgba[2] = in[2];
    rgba[3] = in[3] - d;
    return true;
}

// blendFunc(SRC_ALPHA, ONE_MINUS_SRC_ALPHA) on all four channels (spec/PARITY.md B2), alpha unclamped (FL6)
TB_HD void blend(float dst[4], const float rgba[4]) {
    const float a = rgba[3], om = 1.0f - a;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float t1 = rgba[k] * a;
        const float t2 = dst[k] * om;
        dst[k] = t1 + t2;
    }
}

// every triangle of the strip, in order, at one pixel
TB_HD bool pixel(const Vertex *verts, int n_vertices, int px, int py, float crestShape, float dst[4]) {
    bool any = false;
    for (int t = 0; t + 2 < n_vertices; ++t) {
        const Vertex &a = verts[t], &b = verts[t + 1], &c = verts[t + 2];
        // cheap reject: the pixel centre against the triangle's bounding box
        const long long cx = static_cast<long long>(px) * 256 + 128, cy = static_cast<long long>(py) * 256 + 128;
        const long long lox = a.x < b.x ? (a.x < c.x ? a.x : c.x) : (b.x < c.x ? b.x : c.x);
        const long long hix = a.x > b.x ? (a.x > c.x ? a.x : c.x) : (b.x > c.x ? b.x : c.x);
        const long long loy = a.y < b.y ? (a.y < c.y ? a.y : c.y) : (b.y < c.y ? b.y : c.y);
        const long long hiy = a.y > b.y ? (a.y > c.y ? a.y : c.y) : (b.y > c.y ? b.y : c.y);
        if (cx < lox || cx > hix || cy < loy || cy > hiy) continue;
        float rgba[4];
        if (shade(a, b, c, px, py, crestShape, rgba)) {
            blend(dst, rgba);
            any = true;
        }
    }
    return any;
}

}  // namespace fl

#if defined(__CUDACC__)
// vertex stage + the pixel bounding box of the strip (bbox: min x, min y, max x, max y; initialised to an empty box)
__global__ void k_flow_line_vertices(const fl::Uniforms U, int n, const float *__restrict__ position, const float *__restrict__ normal,
                                     const float *__restrict__ miter, const float *__restrict__ previous,
                                     const float *__restrict__ time, const float *__restrict__ dt, int W, int H,
                                     fl::Vertex *__restrict__ verts, int *__restrict__ bbox) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fl::Vertex v;
    fl::vertex_stage(U, position[2 * i], position[2 * i + 1], normal[2 * i], normal[2 * i + 1], miter[i], previous[2 * i],
                     previous[2 * i + 1], time[i], dt[i], W, H, v);
    verts[i] = v;
    if (v.ok) {
        // pixels whose centre can lie inside: ceil / floor of (x - 128) / 256
        const long long x0 = (v.x - 128 + 255) >> 8, x1 = (v.x - 128) >> 8, y0 = (v.y - 128 + 255) >> 8, y1 = (v.y - 128) >> 8;
        atomicMin(bbox + 0, static_cast<int>(max(x1, -1ll)));
        atomicMin(bbox + 1, static_cast<int>(max(y1, -1ll)));
        atomicMax(bbox + 2, static_cast<int>(min(x0, static_cast<long long>(W))));
        atomicMax(bbox + 3, static_cast<int>(min(y0, static_cast<long long>(H))));
    }
}

// one thread per texel of the flow grid: all triangles in strip order (the blend is ordered)
__global__ void __launch_bounds__(256) k_flow_line_raster(const fl::Vertex *__restrict__ verts, int n, float crestShape,
                                                           const int *__restrict__ bbox, float4 *__restrict__ flow, int W, int H) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y;
    if (px >= W || py >= H) return;
    if (px < bbox[0] || py < bbox[1] || px > bbox[2] || py > bbox[3]) return;
    const size_t t = static_cast<size_t>(py) * W + px;
    const float4 f = flow[t];
    float d[4] = {f.x, f.y, f.z, f.w};
    if (fl::pixel(verts, n, px, py, crestShape, d)) flow[t] = make_float4(d[0], d[1], d[2], d[3]);
}
#endif

}  // namespace tb
