/*
 * tendrils_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See tendrils_oracle.h for the contract.  PARITY UNPINNED by the reference's own tests
 * (it has none); pinned against tests/golden/glsl_v1.npz (reference shader text run
 * through tools/glsl_interp.py) and spec/PARITY.md.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared
 * Every arithmetic statement below is ONE binary32 operation per GLSL operator, in
 * GLSL source order (left-to-right, component-wise); nothing may be contracted or
 * re-associated.  Do not add -ffast-math.
 */
#include "tendrils_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int or_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void or_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}


/* ------------------------------------------------------------------------------------
 * GLSL ES 1.00 built-ins (spec section 8), NaN behaviour as the spec's defining formulas.
 * ---------------------------------------------------------------------------------- */
static inline float g_min(float x, float y) { return (y < x) ? y : x; }   /* "y if y < x" */
static inline float g_max(float x, float y) { return (x < y) ? y : x; }   /* "y if x < y" */
static inline float g_step(float edge, float x) { return (x < edge) ? 0.0f : 1.0f; }
static inline float g_fract(float x) { return x - floorf(x); }
static inline float g_mod(float x, float y) { return x - y * floorf(x / y); }
static inline float g_mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
static inline float g_length2(float x, float y) { return sqrtf(x * x + y * y); }

/* sin/cos: spec/PARITY.md "TSIN-1".  Cody-Waite reduction by pi/2 in three steps and the
 * Cephes single-precision minimax polynomials; only + - * and a magic-number rint, so that
 * a CPU and a GPU evaluate it bit-identically.  Domain |x| <= 1e5, otherwise NaN. */
static inline void sincos_core(float x, float *s_out, float *c_out) {
    if (!(fabsf(x) <= 100000.0f)) { *s_out = NAN; *c_out = NAN; return; }
    float kf = x * 0.636619772f;
    kf = (kf + 12582912.0f) - 12582912.0f;            /* round to nearest even integer */
    float r = x - kf * 1.5703125f;
    r = r - kf * 4.837512969970703125e-4f;
    r = r - kf * 7.54978995489188216e-8f;
    int q = ((int)kf) & 3;
    float z = r * r;
    float ps = -1.9515295891e-4f * z;
    ps = ps + 8.3321608736e-3f;
    ps = ps * z;
    ps = ps - 1.6666654611e-1f;
    ps = ps * z;
    ps = ps * r;
    float s = ps + r;
    float pc = 2.443315711809948e-5f * z;
    pc = pc - 1.388731625493765e-3f;
    pc = pc * z;
    pc = pc + 4.166664568298827e-2f;
    pc = pc * z;
    pc = pc * z;
    float hz = 0.5f * z;
    pc = pc - hz;
    float c = pc + 1.0f;
    switch (q) {
        case 0: *s_out = s;  *c_out = c;  break;
        case 1: *s_out = c;  *c_out = -s; break;
        case 2: *s_out = -s; *c_out = -c; break;
        default: *s_out = -c; *c_out = s; break;
    }
}
float or_sin(float x) { float s, c; sincos_core(x, &s, &c); return s; }
float or_cos(float x) { float s, c; sincos_core(x, &s, &c); return c; }

/* glsl-random@0.0.5 (text in docs/js/demo.js.map, e.g. source ./src/spawn/ball/index.frag):
 *   dt = dot(co, vec2(12.9898, 78.233)); sn = mod(dt, 3.14); fract(sin(sn) * 43758.5453) */
float or_random(float cx, float cy) {
    float dt = cx * 12.9898f + cy * 78.233f;
    float sn = g_mod(dt, 3.14f);
    return g_fract(or_sin(sn) * 43758.5453f);
}

/* ------------------------------------------------------------------------------------
 * a4: glsl-noise@0.0.0 simplex/3d -- text: docs/js/index.js.map sourcesContent[79]
 * (the glslified src/logic.frag); call sites src/logic.frag:67-68.
 * ---------------------------------------------------------------------------------- */
static inline float n_mod289(float x) { return x - floorf(x * (1.0f / 289.0f)) * 289.0f; }
static inline float n_permute(float x) { return n_mod289(((x * 34.0f) + 1.0f) * x); }

float or_snoise3(float vx, float vy, float vz) {
    const float Cx = 1.0f / 6.0f, Cy = 1.0f / 3.0f;
    /* first corner */
    float d = (vx * Cy + vy * Cy) + vz * Cy;
    float ix = floorf(vx + d), iy = floorf(vy + d), iz = floorf(vz + d);
    float e = (ix * Cx + iy * Cx) + iz * Cx;
    float x0x = (vx - ix) + e, x0y = (vy - iy) + e, x0z = (vz - iz) + e;
    /* other corners */
    float gx = g_step(x0y, x0x), gy = g_step(x0z, x0y), gz = g_step(x0x, x0z);
    float lx = 1.0f - gx, ly = 1.0f - gy, lz = 1.0f - gz;
    float i1x = g_min(gx, lz), i1y = g_min(gy, lx), i1z = g_min(gz, ly);
    float i2x = g_max(gx, lz), i2y = g_max(gy, lx), i2z = g_max(gz, ly);
    float x1x = (x0x - i1x) + Cx, x1y = (x0y - i1y) + Cx, x1z = (x0z - i1z) + Cx;
    float x2x = (x0x - i2x) + Cy, x2y = (x0y - i2y) + Cy, x2z = (x0z - i2z) + Cy;
    float x3x = x0x - 0.5f, x3y = x0y - 0.5f, x3z = x0z - 0.5f;
    /* permutations */
    ix = n_mod289(ix); iy = n_mod289(iy); iz = n_mod289(iz);
    float p[4];
    {
        float az[4] = { iz + 0.0f, iz + i1z, iz + i2z, iz + 1.0f };
        float by[4] = { 0.0f, i1y, i2y, 1.0f };
        float bx[4] = { 0.0f, i1x, i2x, 1.0f };
        for (int k = 0; k < 4; ++k) {
            float t = n_permute(az[k]);
            t = n_permute((t + iy) + by[k]);
            p[k] = n_permute((t + ix) + bx[k]);
        }
    }
    /* gradients */
    const float n_ = 0.142857142857f;
    const float nsx = n_ * 2.0f - 0.0f, nsy = n_ * 0.5f - 1.0f, nsz = n_ * 1.0f - 0.0f;
    float gxk[4], gyk[4], hk[4];
    for (int k = 0; k < 4; ++k) {
        float j = p[k] - 49.0f * floorf((p[k] * nsz) * nsz);
        float x_ = floorf(j * nsz);
        float y_ = floorf(j - 7.0f * x_);
        float x = x_ * nsx + nsy;
        float y = y_ * nsx + nsy;
        float h = (1.0f - fabsf(x)) - fabsf(y);
        float sx = floorf(x) * 2.0f + 1.0f;
        float sy = floorf(y) * 2.0f + 1.0f;
        float sh = -g_step(h, 0.0f);
        gxk[k] = x + sx * sh;
        gyk[k] = y + sy * sh;
        hk[k] = h;
    }
    float xs[4][3] = { { x0x, x0y, x0z }, { x1x, x1y, x1z }, { x2x, x2y, x2z }, { x3x, x3y, x3z } };
    float mk[4], dk[4];
    for (int k = 0; k < 4; ++k) {
        float px = gxk[k], py = gyk[k], pz = hk[k];
        float nrm = 1.79284291400159f - 0.85373472095314f * ((px * px + py * py) + pz * pz);
        px *= nrm; py *= nrm; pz *= nrm;
        float xx = xs[k][0], xy = xs[k][1], xz = xs[k][2];
        float m = g_max(0.6f - ((xx * xx + xy * xy) + xz * xz), 0.0f);
        m = m * m;
        mk[k] = m * m;
        dk[k] = (px * xx + py * xy) + pz * xz;
    }
    return 42.0f * (((mk[0] * dk[0] + mk[1] * dk[1]) + mk[2] * dk[2]) + mk[3] * dk[3]);
}

/* ------------------------------------------------------------------------------------
 * Texture fetch, NEAREST + CLAMP_TO_EDGE on a float RGBA texture (gl-fbo 2.0.5 / gl-texture2d
 * 2.1.0 defaults; docs/js/index.js.map sourcesContent[41],[42]).  texel = floor(u*size),
 * clamped; a NaN coordinate selects texel 0 (spec/PARITY.md Q5).
 * ---------------------------------------------------------------------------------- */
static inline int texel_of(float u, int size) {
    float f = floorf(u * (float)size);
    if (!(f > 0.0f)) return 0;
    if (f > (float)(size - 1)) return size - 1;
    return (int)f;
}

static inline float vary(float base, float offset, float variance) {       /* logic.frag:41-43 */
    return base + (offset * variance * base);
}

/* a3: src/logic.frag:45-101 (+ flow/flow-at-screen-pos.glsl:13-27, flow/get.glsl:3-5,
 * map/pos-to-uv.glsl via glsl-map: outMin + (outMax-outMin)*(v-inMin)/(inMax-inMin)). */
void or_integrate(const or_params *P, int PW, int PH, int x0, int x1,
                  const float *state_in, float *state_out, const float *targets,
                  const float *flow, int W, int H, float time, float dt) {
    const float resx = (float)PW, resy = (float)PH;
#pragma omp parallel for schedule(static)
    for (int x = x0; x < x1; ++x) {
        for (int y = 0; y < PH; ++y) {
            size_t p = (size_t)x * PH + y;
            const float *st = state_in + 4 * p;
            float *out = state_out + 4 * p;
            float posx = st[0], posy = st[1], velx = st[2], vely = st[3];
            if (!(posx != -1000000.0f || posy != -1000000.0f)) {   /* pos == inert */
                out[0] = posx; out[1] = posy; out[2] = velx; out[3] = vely;
                continue;
            }
            float fcx = (float)x + 0.5f, fcy = (float)y + 0.5f;
            float uvx = fcx / resx, uvy = fcy / resy;
            float i = (fcx + (fcy * resx)) / (resx * resy);

            float ns = vary(P->noiseScale, i, P->varyNoiseScale);
            float npx = posx * ns, npy = posy * ns;
            float noiseTime = time * vary(P->noiseSpeed, i, P->varyNoiseSpeed);
            float wx = or_snoise3(npx, npy, uvx + noiseTime);
            float wy = or_snoise3(npx, npy, (uvy + noiseTime) + 1234.5678f);

            /* flowAtScreenPos(pos*viewSize, ...), levels = stride = 1 */
            float spx = posx * P->viewSize[0], spy = posy * P->viewSize[1];
            float fu = 0.0f + (1.0f - 0.0f) * (spx - -1.0f) / (1.0f - -1.0f);
            float fv = 0.0f + (1.0f - 0.0f) * (spy - -1.0f) / (1.0f - -1.0f);
            const float *fd = flow + 4 * ((size_t)texel_of(fv, H) * W + texel_of(fu, W));
            float fac = g_max(0.0f, 1.0f - ((time - fd[2]) * P->flowDecay));
            float ffx = 0.0f + (fd[0] * fac) * 1.0f, ffy = 0.0f + (fd[1] * fac) * 1.0f;
            ffx = ffx / 1.0f; ffy = ffy / 1.0f;                     /* flowForce/flowMax */

            float vforce = vary(P->forceWeight, i, P->varyForce);
            float vflow = vary(P->flowWeight, i, P->varyFlow);
            float vnoise = vary(P->noiseWeight, i, P->varyNoise);
            float nvx = ((velx * P->damping) * dt) + (vforce * (((ffx * dt) * vflow) + ((wx * dt) * vnoise)));
            float nvy = ((vely * P->damping) * dt) + (vforce * (((ffy * dt) * vflow) + ((wy * dt) * vnoise)));

            float vt = vary(P->target, i, P->varyTarget);
            nvx = nvx + (targets[4 * p + 0] - posx) * vt;
            nvy = nvy + (targets[4 * p + 1] - posy) * vt;

            float speed = g_length2(nvx, nvy);
            float sc = g_min(speed, P->speedLimit) / speed;
            nvx = nvx * sc; nvy = nvy * sc;
            out[0] = posx + nvx; out[1] = posy + nvy; out[2] = nvx; out[3] = nvy;
        }
    }
}

/* ------------------------------------------------------------------------------------
 * a8/a9: vertex LUT (src/particles.js:171-190) seen through stateAtFrame
 * (src/state/state-at-frame.glsl:12-22).  geomShape = (PW, 2*PH) (src/index.js:197).
 * ---------------------------------------------------------------------------------- */
void or_vertex_table(int PH, int *row_of_vertex, int *cur_of_vertex) {
    int h = 2 * PH; if (h < 2) h = 2;
    double invY = 1.0 / (double)(h - 1);
    for (int j = 0; j < 2 * PH; ++j) {
        float uvy = (float)((double)j * invY);            /* Float32Array store */
        float nearIndex = uvy * (float)PH;
        float fl = floorf(nearIndex);
        float offset = nearIndex - fl;                    /* fract */
        float lookupy = fl / (float)PH;
        row_of_vertex[j] = texel_of(lookupy, PH);
        cur_of_vertex[j] = (offset > 0.25f) ? 1 : 0;
    }
}
void or_column_table(int PW, int *col) {
    int w = PW; if (w < 2) w = 2;
    double invX = 1.0 / (double)(w - 1);
    for (int i = 0; i < PW; ++i) {
        float uvx = (float)((double)i * invX);
        col[i] = texel_of(uvx, PW);
    }
}

/* src/flow/apply/state.glsl:5-16 -- colour of a flow vertex. */
static inline void flow_colour(float vx, float vy, float time, float speedLimit, float c[4]) {
    c[0] = vx; c[1] = vy; c[2] = time;
    c[3] = g_min(g_length2(vx, vy) / speedLimit, 1.0f);
}

static inline int finite4(const float *s) {
    return isfinite(s[0]) && isfinite(s[1]) && isfinite(s[2]) && isfinite(s[3]);
}

/* The vertex stage alone (src/flow/vert/main.vert:10-17 + stateAtFrame + apply/state.glsl), exposed so that
 * tests can compare it with the reference's flow/index.vert run through tools/glsl_interp.py.
 * out6 = (gl_Position.xy, color.rgba); returns 1 when gl_Position is written (state not inert). */
int or_flow_vertex(const or_params *P, int PW, int PH, int i, int j, const float *cur, const float *prev,
                   float time, float *out6) {
    int *row = (int *)malloc(sizeof(int) * 2 * PH), *isc = (int *)malloc(sizeof(int) * 2 * PH);
    int *col = (int *)malloc(sizeof(int) * PW);
    or_vertex_table(PH, row, isc);
    or_column_table(PW, col);
    const float *st = (isc[j] ? cur : prev) + 4 * ((size_t)col[i] * PH + row[j]);
    free(row); free(isc); free(col);
    if (!(st[0] != -1000000.0f || st[1] != -1000000.0f)) return 0;
    float c[4];
    flow_colour(st[2], st[3], time, P->speedLimit, c);
    out6[0] = st[0] * P->viewSize[0]; out6[1] = st[1] * P->viewSize[1];
    out6[2] = c[0]; out6[3] = c[1]; out6[4] = c[2]; out6[5] = c[3];
    return 1;
}

/* One fragment, blendFunc(SRC_ALPHA, ONE_MINUS_SRC_ALPHA) on all four channels
 * (src/index.js:267-268): dst = src*a + dst*(1-a).  */
static inline void blend_over(float *dst, const float c[4]) {
    float a = c[3], om = 1.0f - a;
    for (int k = 0; k < 4; ++k) {
        float t1 = c[k] * a;
        float t2 = dst[k] * om;
        dst[k] = t1 + t2;
    }
}

/* GL_LINES, width 1, spec/PARITY.md "RASTER-1": centre-sampled major-axis rule (the
 * diamond-exit rule of OpenGL ES 2.0 section 3.4.1 up to its permitted deviations),
 * half-open towards the second vertex, scissored to the grid, attributes c0 + t*(c1-c0). */
typedef void (*frag_fn)(void *ctx, int gx, int gy, const float c[4]);

static long long raster_line(float xa, float ya, float xb, float yb, const float c0[4], const float c1[4],
                             int W, int H, frag_fn emit, void *ctx) {
    long long n = 0;
    float dx = xb - xa, dy = yb - ya;
    float adx = fabsf(dx), ady = fabsf(dy);
    if (adx >= ady) {
        if (!(adx > 0.0f)) return 0;
        float lo = g_min(xa, xb), hi = g_max(xa, xb);
        float flo = floorf(lo) - 1.0f, fhi = floorf(hi) + 1.0f;
        if (flo < 0.0f) flo = 0.0f;
        if (fhi > (float)(W - 1)) fhi = (float)(W - 1);
        if (!(flo <= fhi)) return 0;
        for (int i = (int)flo; i <= (int)fhi; ++i) {
            float ic = (float)i + 0.5f;
            int in = (dx > 0.0f) ? (xa <= ic && ic < xb) : (xb < ic && ic <= xa);
            if (!in) continue;
            float t = (ic - xa) / dx;
            float yy = ya + t * dy;
            float fj = floorf(yy);
            if (!(fj >= 0.0f && fj <= (float)(H - 1))) continue;
            float c[4];
            for (int k = 0; k < 4; ++k) c[k] = c0[k] + t * (c1[k] - c0[k]);
            emit(ctx, i, (int)fj, c);
            ++n;
        }
    } else {
        float lo = g_min(ya, yb), hi = g_max(ya, yb);
        float flo = floorf(lo) - 1.0f, fhi = floorf(hi) + 1.0f;
        if (flo < 0.0f) flo = 0.0f;
        if (fhi > (float)(H - 1)) fhi = (float)(H - 1);
        if (!(flo <= fhi)) return 0;
        for (int j = (int)flo; j <= (int)fhi; ++j) {
            float jc = (float)j + 0.5f;
            int in = (dy > 0.0f) ? (ya <= jc && jc < yb) : (yb < jc && jc <= ya);
            if (!in) continue;
            float t = (jc - ya) / dy;
            float xx = xa + t * dx;
            float fi = floorf(xx);
            if (!(fi >= 0.0f && fi <= (float)(W - 1))) continue;
            float c[4];
            for (int k = 0; k < 4; ++k) c[k] = c0[k] + t * (c1[k] - c0[k]);
            emit(ctx, (int)fi, j, c);
            ++n;
        }
    }
    return n;
}

typedef struct { float *flow; int W; } blend_ctx;
static void emit_blend(void *vctx, int gx, int gy, const float c[4]) {
    blend_ctx *b = (blend_ctx *)vctx;
    blend_over(b->flow + 4 * ((size_t)gy * b->W + gx), c);
}

/* a7-a10.  Serial by construction: the blend is order dependent (primitive order
 * p = x*PH + k, src/particles.js:182-186). */
long long or_splat(const or_params *P, int PW, int PH, int x0, int x1,
                   const float *cur, const float *prev, float *flow, int W, int H, float time) {
    int *row = (int *)malloc(sizeof(int) * 2 * PH), *isc = (int *)malloc(sizeof(int) * 2 * PH);
    int *col = (int *)malloc(sizeof(int) * PW);
    or_vertex_table(PH, row, isc);
    or_column_table(PW, col);
    blend_ctx bc = { flow, W };
    long long frags = 0;
    const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
    for (int x = x0; x < x1; ++x) {
        for (int k = 0; k < PH; ++k) {
            const float *sa = (isc[2 * k] ? cur : prev) + 4 * ((size_t)col[x] * PH + row[2 * k]);
            const float *sb = (isc[2 * k + 1] ? cur : prev) + 4 * ((size_t)col[x] * PH + row[2 * k + 1]);
            if (sa == sb) continue;                                    /* zero-length line */
            /* a vertex whose state is inert leaves gl_Position unwritten: culled (PARITY V1);
             * a primitive with a non-finite vertex is culled (PARITY V2). */
            if (!(sa[0] != -1000000.0f || sa[1] != -1000000.0f)) continue;
            if (!(sb[0] != -1000000.0f || sb[1] != -1000000.0f)) continue;
            if (!finite4(sa) || !finite4(sb)) continue;
            float ca[4], cb[4];
            flow_colour(sa[2], sa[3], time, P->speedLimit, ca);
            flow_colour(sb[2], sb[3], time, P->speedLimit, cb);
            float xa = (sa[0] * P->viewSize[0]) * hw + hw, ya = (sa[1] * P->viewSize[1]) * hh + hh;
            float xb = (sb[0] * P->viewSize[0]) * hw + hw, yb = (sb[1] * P->viewSize[1]) * hh + hh;
            frags += raster_line(xa, ya, xb, yb, ca, cb, W, H, emit_blend, &bc);
        }
    }
    free(row); free(isc); free(col);
    return frags;
}

/* Multi-threaded form of or_splat for the timed CPU baseline.  Same result bit for bit
 * (tests/test_oracle.py): fragments are generated in draw order, stably counting-sorted by
 * texel, and each texel's list is folded sequentially. */
typedef struct { int texel; float cx, cy, a; } o_frag;
typedef struct { o_frag *out; long long n; int W; } collect_ctx;
static void emit_collect(void *vctx, int gx, int gy, const float c[4]) {
    collect_ctx *cc = (collect_ctx *)vctx;
    if (cc->out) { o_frag f = { gy * cc->W + gx, c[0], c[1], c[3] }; cc->out[cc->n] = f; }
    cc->n++;
}

static long long splat_column(const or_params *P, int PH, int x, const int *row, const int *isc, const int *col,
                              const float *cur, const float *prev, int W, int H, float time, o_frag *out) {
    collect_ctx cc = { out, 0, W };
    const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
    for (int k = 0; k < PH; ++k) {
        const float *sa = (isc[2 * k] ? cur : prev) + 4 * ((size_t)col[x] * PH + row[2 * k]);
        const float *sb = (isc[2 * k + 1] ? cur : prev) + 4 * ((size_t)col[x] * PH + row[2 * k + 1]);
        if (sa == sb) continue;
        if (!(sa[0] != -1000000.0f || sa[1] != -1000000.0f)) continue;
        if (!(sb[0] != -1000000.0f || sb[1] != -1000000.0f)) continue;
        if (!finite4(sa) || !finite4(sb)) continue;
        float ca[4], cb[4];
        flow_colour(sa[2], sa[3], time, P->speedLimit, ca);
        flow_colour(sb[2], sb[3], time, P->speedLimit, cb);
        float xa = (sa[0] * P->viewSize[0]) * hw + hw, ya = (sa[1] * P->viewSize[1]) * hh + hh;
        float xb = (sb[0] * P->viewSize[0]) * hw + hw, yb = (sb[1] * P->viewSize[1]) * hh + hh;
        raster_line(xa, ya, xb, yb, ca, cb, W, H, emit_collect, &cc);
    }
    return cc.n;
}

long long or_splat_mt(const or_params *P, int PW, int PH, int x0, int x1,
                      const float *cur, const float *prev, float *flow, int W, int H, float time) {
    int *row = (int *)malloc(sizeof(int) * 2 * PH), *isc = (int *)malloc(sizeof(int) * 2 * PH);
    int *col = (int *)malloc(sizeof(int) * PW);
    or_vertex_table(PH, row, isc);
    or_column_table(PW, col);
    const int ncol = x1 - x0;
    const size_t G = (size_t)W * H;
    long long *cstart = (long long *)calloc((size_t)ncol + 1, sizeof(long long));
#pragma omp parallel for schedule(dynamic, 8)
    for (int x = x0; x < x1; ++x)
        cstart[x - x0 + 1] = splat_column(P, PH, x, row, isc, col, cur, prev, W, H, time, NULL);
    for (int i = 0; i < ncol; ++i) cstart[i + 1] += cstart[i];
    const long long F = cstart[ncol];
    o_frag *frags = (o_frag *)malloc(sizeof(o_frag) * (size_t)(F > 0 ? F : 1));
    o_frag *sorted = (o_frag *)malloc(sizeof(o_frag) * (size_t)(F > 0 ? F : 1));
#pragma omp parallel for schedule(dynamic, 8)
    for (int x = x0; x < x1; ++x)
        splat_column(P, PH, x, row, isc, col, cur, prev, W, H, time, frags + cstart[x - x0]);
    /* stable parallel counting sort by texel */
    int T = or_num_threads();
    if (T > 64) T = 64;
    long long *hist = (long long *)calloc((size_t)T * G, sizeof(long long));
    long long *tstart = (long long *)malloc(sizeof(long long) * (G + 1));
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        long long b = F * t / T, e = F * (t + 1) / T;
        long long *h = hist + (size_t)t * G;
        for (long long i = b; i < e; ++i) h[frags[i].texel]++;
    }
    {
        long long run = 0;
        for (size_t g = 0; g < G; ++g) {
            tstart[g] = run;
            for (int t = 0; t < T; ++t) { long long c = hist[(size_t)t * G + g]; hist[(size_t)t * G + g] = run; run += c; }
        }
        tstart[G] = run;
    }
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        long long b = F * t / T, e = F * (t + 1) / T;
        long long *h = hist + (size_t)t * G;
        for (long long i = b; i < e; ++i) sorted[h[frags[i].texel]++] = frags[i];
    }
#pragma omp parallel for schedule(dynamic, 1024)
    for (long long g = 0; g < (long long)G; ++g) {
        float *dst = flow + 4 * g;
        for (long long i = tstart[g]; i < tstart[g + 1]; ++i) {
            float c[4] = { sorted[i].cx, sorted[i].cy, time, sorted[i].a };
            blend_over(dst, c);
        }
    }
    free(hist); free(tstart); free(frags); free(sorted); free(cstart);
    free(row); free(isc); free(col);
    return F;
}

/* ------------------------------------------------------------------------------------
 * Spawners
 * ---------------------------------------------------------------------------------- */
void or_spawn_init(int PW, int PH, int x0, int x1, float *out) {      /* spawn/init/index.frag:5-10 */
    (void)PW;
    for (size_t p = (size_t)x0 * PH; p < (size_t)x1 * PH; ++p) {
        out[4 * p + 0] = -1000000.0f; out[4 * p + 1] = -1000000.0f;
        out[4 * p + 2] = 0.0f; out[4 * p + 3] = 0.0f;
    }
}

static const float TAU = 6.28318530717958647692f;

void or_spawn_ball(int PW, int PH, int x0, int x1, float radius, float speed, float *out) {
    (void)PW;                                                           /* spawn/ball/index.frag:11-19 */
#pragma omp parallel for schedule(static)
    for (int x = x0; x < x1; ++x)
        for (int y = 0; y < PH; ++y) {
            float fx = (float)x + 0.5f, fy = (float)y + 0.5f;
            float r0 = or_random(fx * 1.7654f + 2.3675f, fy * 1.7654f + 2.3675f);
            float r1 = or_random(fx * 1.23494f + 0.36434f, fy * 1.23494f + 0.36434f);
            float r2 = or_random(fx * 0.327789f + 3.498787f, fy * 0.327789f + 3.498787f);
            float r3 = or_random(fx * 9.0374f + 0.2773f, fy * 9.0374f + 0.2773f);
            float s, c;
            float *o = out + 4 * ((size_t)x * PH + y);
            sincos_core(r0 * TAU, &s, &c);
            o[0] = (c * r1) * radius; o[1] = (s * r1) * radius;
            sincos_core(r2 * TAU, &s, &c);
            o[2] = (c * r3) * speed; o[3] = (s * r3) * speed;
        }
}

/* spawn/pixels/frag/head.frag:28-34 */
static inline void spawn_to_pos(const or_spawn_pixels *S, float u, float v, float time, float *px, float *py) {
    float tt = time * 0.001f;
    float ox = g_mix(-S->jitter[0], S->jitter[0], or_random((u - 1.2345f) + tt, (v - 1.2345f) + tt));
    float oy = g_mix(-S->jitter[1], S->jitter[1], or_random((u + 1.2345f) + tt, (v + 1.2345f) + tt));
    float uu = u + ox, vv = v + oy;
    /* uvToPos: map(uv, 0, 1, -1, 1) = outMin + (outMax-outMin)*(v-inMin)/(inMax-inMin) */
    float qx = -1.0f + (1.0f - -1.0f) * (uu - 0.0f) / (1.0f - 0.0f);
    float qy = -1.0f + (1.0f - -1.0f) * (vv - 0.0f) / (1.0f - 0.0f);
    qx = (qx * 1.0f) * S->spawnSize[0];
    qy = (qy * -1.0f) * S->spawnSize[1];
    const float *m = S->spawnMatrix;            /* (m * vec3(q, 1)).xy, column-major */
    *px = (m[0] * qx + m[3] * qy) + m[6] * 1.0f;
    *py = (m[1] * qx + m[4] * qy) + m[7] * 1.0f;
}

/* filter/vignette.glsl:5-24 with vignette-head.glsl:4-6 (curve (.1,1,1), mid .5, limit .6) */
static inline float vignette(float u, float v) {
    float amount = g_min(1.0f - (g_length2(u - 0.5f, v - 0.5f) / 0.6f), 1.0f);
    float t = amount, ut = 1.0f - t;
    float bz = (0.1f * ut + 1.0f * t) * ut + (1.0f * ut + 1.0f * t) * t;     /* utils/bezier.glsl:9-13 */
    return g_max(0.0f, bz);
}

/* libs/glsl-hsv/rgb-hsv.glsl:4-11 */
static inline void rgb2hsv(float r, float g, float b, float hsv[3]) {
    const float kx = 0.0f, ky = -1.0f / 3.0f, kz = 2.0f / 3.0f, kw = -1.0f, e = 1.0e-10f;
    float p[4], q[4];
    if (g < b) { p[0] = b; p[1] = g; p[2] = kw; p[3] = kz; } else { p[0] = g; p[1] = b; p[2] = kx; p[3] = ky; }
    if (r < p[0]) { q[0] = p[0]; q[1] = p[1]; q[2] = p[3]; q[3] = r; } else { q[0] = r; q[1] = p[1]; q[2] = p[2]; q[3] = p[0]; }
    float d = q[0] - g_min(q[3], q[1]);
    hsv[0] = fabsf(q[2] + (q[3] - q[1]) / (6.0f * d + e));
    hsv[1] = d / (q[0] + e);
    hsv[2] = q[0];
}

static inline void fetch_image(const float *image, int IW, int IH, float u, float v, float px[4]) {
    const float *t = image + 4 * ((size_t)texel_of(v, IH) * IW + texel_of(u, IW));
    px[0] = t[0]; px[1] = t[1]; px[2] = t[2]; px[3] = t[3];
}

/* apply/<kind>.glsl composed with filter/pass/vignette.glsl when vignette != 0 */
static inline void apply_pixel(const or_spawn_pixels *S, int apply, int vig, float u, float v,
                               float posx, float posy, const float pxin[4], float time, float out[4]) {
    float px[4] = { pxin[0], pxin[1], pxin[2], pxin[3] };
    if (vig) {
        float w = vignette(u, v);
        for (int k = 0; k < 4; ++k) px[k] = px[k] * w;
    }
    if (apply == OR_APPLY_COLOR) {                         /* apply/color.glsl:13-17 */
        float hsv[3], s, c;
        rgb2hsv(px[0], px[1], px[2], hsv);
        sincos_core((hsv[0] + (time * 0.00003f)) * TAU, &s, &c);
        out[0] = posx; out[1] = posy;
        out[2] = ((c * hsv[1]) * hsv[2]) * px[3];
        out[3] = ((s * hsv[1]) * hsv[2]) * px[3];
    } else if (apply == OR_APPLY_BRIGHTEST) {              /* apply/brightest.glsl:12-16 */
        float dd = px[0] * px[2] + px[1] * px[3];          /* dot(pixel.rg, pixel.ba) */
        float ang = g_mod(or_random(u * dd, v * dd), 1.0f) * TAU;
        float luma = (px[0] * 0.299f + px[1] * 0.587f) + px[2] * 0.114f;   /* glsl-luma@1.0.1 */
        float s, c;
        sincos_core(ang, &s, &c);
        out[0] = posx; out[1] = posy;
        out[2] = (c * luma) * px[3];
        out[3] = (s * luma) * px[3];
    } else if (apply == OR_APPLY_IDENTITY) {               /* apply/identity.glsl */
        out[0] = px[0]; out[1] = px[1]; out[2] = px[2]; out[3] = px[3];
    } else {                                               /* apply/flow.glsl + flow/get.glsl:3-5 */
        float fac = g_max(0.0f, 1.0f - ((time - px[2]) * S->flowDecay));
        out[0] = posx; out[1] = posy; out[2] = px[0] * fac; out[3] = px[1] * fac;
    }
}

/* a14: spawn/pixels/index.frag -> frag/direct-main.frag:9-20 (apply = colour o vignette) */
void or_spawn_pixels_direct(const or_spawn_pixels *S, int PW, int PH, int x0, int x1,
                            const float *image, int IW, int IH, float time, float *out) {
#pragma omp parallel for schedule(static)
    for (int x = x0; x < x1; ++x)
        for (int y = 0; y < PH; ++y) {
            /* uv = (fc/dataRes)*(geomRes/dataRes), geomRes = (PW, 2*PH) */
            float u = (((float)x + 0.5f) / (float)PW) * ((float)PW / (float)PW);
            float v = (((float)y + 0.5f) / (float)PH) * ((float)(2 * PH) / (float)PH);
            float posx, posy, px[4], st[4];
            spawn_to_pos(S, u, v, time, &posx, &posy);
            fetch_image(image, IW, IH, u, v, px);
            apply_pixel(S, OR_APPLY_COLOR, 1, u, v, posx, posy, px, time, st);
            float *o = out + 4 * ((size_t)x * PH + y);
            o[0] = st[0]; o[1] = st[1]; o[2] = st[2] * S->speed; o[3] = st[3] * S->speed;
        }
}

/* a15: spawn/pixels/{best,bright,color,data,flow}-sample.frag -> frag/best-sample-main.frag:21-46,
 * test = test/particles.glsl:8-10 (length2 of .zw). */
void or_spawn_pixels_sample(const or_spawn_pixels *S, int apply, int vig, int samples,
                            int PW, int PH, int x0, int x1, const float *state_in,
                            const float *image, int IW, int IH, float time, float *out) {
#pragma omp parallel for schedule(static)
    for (int x = x0; x < x1; ++x)
        for (int y = 0; y < PH; ++y) {
            size_t p = (size_t)x * PH + y;
            float u = ((float)x + 0.5f) / (float)PW, v = ((float)y + 0.5f) / (float)PH;
            float st[4] = { state_in[4 * p], state_in[4 * p + 1], state_in[4 * p + 2], state_in[4 * p + 3] };
            float k0 = 1.2345f + (time * 0.001f);
            float base[4] = { (st[0] + u) + k0, (st[1] + v) + k0, (st[2] + u) + k0, (st[3] + v) + k0 };
            for (int n = 0; n < samples; ++n) {
                float fn = (float)n;
                float su = g_mod(or_random(base[0] + fn, base[1] + fn), 1.0f);
                float sv = g_mod(or_random(base[2] + fn, base[3] + fn), 1.0f);
                float posx, posy, px[4], o[4];
                spawn_to_pos(S, su, sv, time, &posx, &posy);
                fetch_image(image, IW, IH, su, sv, px);
                apply_pixel(S, apply, vig, su, sv, posx, posy, px, time, o);
                o[2] = o[2] * S->speed; o[3] = o[3] * S->speed;
                float tc = st[2] * st[2] + st[3] * st[3];
                float tn = o[2] * o[2] + o[3] * o[3];
                if (!(tc > S->bias * tn)) { st[0] = o[0]; st[1] = o[1]; st[2] = o[2]; st[3] = o[3]; }
            }
            out[4 * p] = st[0]; out[4 * p + 1] = st[1]; out[4 * p + 2] = st[2]; out[4 * p + 3] = st[3];
        }
}

/* ------------------------------------------------------------------------------------
 * f1: optical flow into the flow grid -- src/optical-flow/index.frag:55-81 drawn with the big triangle
 * (src/screen/index.vert:6-10) over the flow FBO under the alpha-over blend.  view / last are RGBA8
 * textures (gl-fbo without `float`), NEAREST + CLAMP_TO_EDGE; a byte b reads as b/255.
 * spec/PARITY.md OF1: the varying uv at fragment centre fc is ((fc/res)*2)-1.
 * ---------------------------------------------------------------------------------- */
static inline float of_gray(const unsigned char *tex, int w, int h, float u, float v) {
    const unsigned char *t = tex + 4 * ((size_t)texel_of(v, h) * w + texel_of(u, w));
    float r = (float)t[0] / 255.0f, g = (float)t[1] / 255.0f, b = (float)t[2] / 255.0f;
    return (r * 0.3f + g * 0.59f) + b * 0.11f;            /* utils/gray-scale.glsl */
}

void or_optical_flow(const float viewSize[2], const float scaleUV[2], float offset, float lambda, float speed,
                     float speedLimit, float time, const unsigned char *view, const unsigned char *last, int iw, int ih,
                     float *flow, int W, int H) {
#pragma omp parallel for schedule(static)
    for (int gy = 0; gy < H; ++gy)
        for (int gx = 0; gx < W; ++gx) {
            float uvx = (((float)gx + 0.5f) / (float)W) * 2.0f - 1.0f;
            float uvy = (((float)gy + 0.5f) / (float)H) * 2.0f - 1.0f;
            float px = (uvx * scaleUV[0]) / viewSize[0], py = (uvy * scaleUV[1]) / viewSize[1];
            float su = 0.0f + (1.0f - 0.0f) * (px - -1.0f) / (1.0f - -1.0f);
            float sv = 0.0f + (1.0f - 0.0f) * (py - -1.0f) / (1.0f - -1.0f);
            float gradX = (of_gray(view, iw, ih, su + offset, sv + 0.0f) - of_gray(view, iw, ih, su - offset, sv - 0.0f)) +
                          (of_gray(last, iw, ih, su + offset, sv + 0.0f) - of_gray(last, iw, ih, su - offset, sv - 0.0f));
            float gradY = (of_gray(view, iw, ih, su + 0.0f, sv + offset) - of_gray(view, iw, ih, su - 0.0f, sv - offset)) +
                          (of_gray(last, iw, ih, su + 0.0f, sv + offset) - of_gray(last, iw, ih, su - 0.0f, sv - offset));
            float gradMag = sqrtf((gradX * gradX + gradY * gradY) + lambda);
            float diff = of_gray(view, iw, ih, su, sv) - of_gray(last, iw, ih, su, sv);
            float vx = (diff * (gradX / gradMag)) * speed, vy = (diff * (gradY / gradMag)) * speed;
            float t = g_length2(vx, vy) / speedLimit, ut = 1.0f - t;
            float bz = (0.0f * ut + 0.0f * t) * ut + (0.0f * ut + 1.0f * t) * t;      /* bezier(vec3(0,0,1), t) */
            float ox = bz * vx, oy = bz * vy;
            float c[4] = { ox, oy, time, g_min(g_length2(ox, oy) / speedLimit, 1.0f) };
            blend_over(flow + 4 * ((size_t)gy * W + gx), c);
        }
}

/* ------------------------------------------------------------------------------------------------
 * f4: pointer flow lines drawn into the flow grid -- src/flow-line/index.vert, src/flow-line/index.frag over the
 * TRIANGLE_STRIP that src/geom/line/index.js builds (two vertices per path point, attributes position, normal,
 * miter, previous, time, dt).  Fixed function per spec/PARITY.md FL3-FL6: vertices snapped to 1/256 pixel,
 * exact integer edge functions, top-left rule, affine interpolation from the edge functions, alpha-over blend
 * in triangle order with the unclamped alpha values.a - d.
 * ------------------------------------------------------------------------------------------------ */
static inline float g_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

/* src/flow-line/index.vert:21-37.  out9 = (gl_Position.xy, values.rgba, crest.xy, sdf) */
void or_flow_line_vertex(const or_flow_line_uniforms *U, const float position[2], const float normal[2], float miter,
                         const float previous[2], float time, float dt, float *out9) {
    float sdf = g_sign(miter);
    float rate = U->speed / g_max(dt, 1.0f);
    float velx = (position[0] - previous[0]) * rate, vely = (position[1] - previous[1]) * rate;
    float values[4];
    flow_colour(velx, vely, time, U->speedLimit, values);        /* flow(vel, speedLimit), time = the attribute */
    float crx = normal[0] * miter, cry = normal[1] * miter;
    float rad = U->rad * values[3];
    float vx = position[0] + (normal[0] * rad) * miter, vy = position[1] + (normal[1] * rad) * miter;   /* expand() */
    out9[0] = vx * U->viewSize[0]; out9[1] = vy * U->viewSize[1];
    out9[2] = values[0]; out9[3] = values[1]; out9[4] = values[2]; out9[5] = values[3];
    out9[6] = crx; out9[7] = cry; out9[8] = sdf;
}

/* src/flow-line/index.frag:10-17.  in7 = (values.rgba, crest.xy, sdf) */
void or_flow_line_fragment(float crestShape, const float *in7, float *rgba) {
    float d = fabsf(in7[6]);
    float speed = g_length2(in7[0], in7[1]) * (1.0f - d);
    float t = d * crestShape;
    float mx = g_mix(in7[0], in7[4], t), my = g_mix(in7[1], in7[5], t);
    float len = g_length2(mx, my);
    rgba[0] = (mx / len) * speed; rgba[1] = (my / len) * speed;   /* normalize(m)*speed */
    rgba[2] = in7[2];
    rgba[3] = in7[3] - d;
}

static inline long long fl_orient(long long ax, long long ay, long long bx, long long by, long long cx, long long cy) {
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
}

/* a centre exactly on the edge a->b of a counter-clockwise triangle belongs to it iff the edge runs down, or is
 * horizontal and runs towards -x (FL4) */
static inline int fl_tie(long long ax, long long ay, long long bx, long long by) {
    long long dx = bx - ax, dy = by - ay;
    return dy < 0 || (dy == 0 && dx < 0);
}

long long or_flow_line(const or_flow_line_uniforms *U, int n_vertices, const float *position, const float *normal,
                       const float *miter, const float *previous, const float *time, const float *dt,
                       float *flow, int W, int H) {
    if (n_vertices < 3) return 0;
    float *vs = (float *)malloc(sizeof(float) * 9 * (size_t)n_vertices);
    long long *fx = (long long *)malloc(sizeof(long long) * 2 * (size_t)n_vertices);
    char *ok = (char *)malloc((size_t)n_vertices);
    const float hw = (float)W / 2.0f, hh = (float)H / 2.0f;
    for (int i = 0; i < n_vertices; ++i) {
        float *o = vs + 9 * (size_t)i;
        or_flow_line_vertex(U, position + 2 * i, normal + 2 * i, miter[i], previous + 2 * i, time[i], dt[i], o);
        float xw = o[0] * hw + hw, yw = o[1] * hh + hh;                       /* FL3 */
        ok[i] = isfinite(xw) && isfinite(yw) && fabsf(xw) < 262144.0f && fabsf(yw) < 262144.0f;
        fx[2 * i] = ok[i] ? llrintf(xw * 256.0f) : 0;
        fx[2 * i + 1] = ok[i] ? llrintf(yw * 256.0f) : 0;
    }
    long long frags = 0;
    for (int t = 0; t + 2 < n_vertices; ++t) {
        if (!(ok[t] && ok[t + 1] && ok[t + 2])) continue;
        const long long x0 = fx[2 * t], y0 = fx[2 * t + 1], x1 = fx[2 * t + 2], y1 = fx[2 * t + 3], x2 = fx[2 * t + 4], y2 = fx[2 * t + 5];
        const long long area = fl_orient(x0, y0, x1, y1, x2, y2);
        if (area == 0) continue;
        const long long sg = area > 0 ? 1 : -1;
        /* tie rule on the counter-clockwise orientation: with area < 0 every edge is walked backwards */
        const int tie0 = sg > 0 ? fl_tie(x1, y1, x2, y2) : fl_tie(x2, y2, x1, y1);
        const int tie1 = sg > 0 ? fl_tie(x2, y2, x0, y0) : fl_tie(x0, y0, x2, y2);
        const int tie2 = sg > 0 ? fl_tie(x0, y0, x1, y1) : fl_tie(x1, y1, x0, y0);
        long long minx = x0 < x1 ? x0 : x1, maxx = x0 > x1 ? x0 : x1, miny = y0 < y1 ? y0 : y1, maxy = y0 > y1 ? y0 : y1;
        if (x2 < minx) minx = x2;
        if (x2 > maxx) maxx = x2;
        if (y2 < miny) miny = y2;
        if (y2 > maxy) maxy = y2;
        long long px0 = (minx - 128 + 255) >> 8, px1 = (maxx - 128) >> 8;      /* ceil / floor of (v-128)/256 */
        long long py0 = (miny - 128 + 255) >> 8, py1 = (maxy - 128) >> 8;
        if (px0 < 0) px0 = 0;
        if (py0 < 0) py0 = 0;
        if (px1 > W - 1) px1 = W - 1;
        if (py1 > H - 1) py1 = H - 1;
        const float *v0 = vs + 9 * (size_t)t + 2, *v1 = v0 + 9, *v2 = v1 + 9;
        for (long long py = py0; py <= py1; ++py)
            for (long long px = px0; px <= px1; ++px) {
                const long long cx = px * 256 + 128, cy = py * 256 + 128;
                const long long e0 = fl_orient(x1, y1, x2, y2, cx, cy), e1 = fl_orient(x2, y2, x0, y0, cx, cy),
                                e2 = fl_orient(x0, y0, x1, y1, cx, cy);
                const long long s0 = e0 * sg, s1 = e1 * sg, s2 = e2 * sg;
                if (!((s0 > 0 || (s0 == 0 && tie0)) && (s1 > 0 || (s1 == 0 && tie1)) && (s2 > 0 || (s2 == 0 && tie2)))) continue;
                const float b1 = (float)((double)e1 / (double)area), b2 = (float)((double)e2 / (double)area);   /* FL5 */
                float in7[7], rgba[4];
                for (int k = 0; k < 7; ++k) in7[k] = (v0[k] + b1 * (v1[k] - v0[k])) + b2 * (v2[k] - v0[k]);
                or_flow_line_fragment(U->crestShape, in7, rgba);
                blend_over(flow + 4 * ((size_t)py * W + (size_t)px), rgba);
                ++frags;
            }
    }
    free(vs); free(fx); free(ok);
    return frags;
}

