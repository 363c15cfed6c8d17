/*
 * tendrils_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, float32 restatement of the reference's per-frame particle step:
 * the GLSL shaders src/logic.frag, src/flow/.., src/spawn/.. of keeffEoghan/tendrils
 * plus the WebGL fixed-function pieces they rely on (NEAREST texture fetch, GL_LINES
 * rasterisation, ordered SRC_ALPHA/ONE_MINUS_SRC_ALPHA blending).
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this
 * path and its shaders cannot be executed in this image (no
 * Node / headless-gl / Mesa).  This oracle is pinned instead against
 *   (1) tests/golden/glsl_v1.npz -- outputs of the reference's OWN shader text
 *       (docs/js/index.js.map, demo.js.map sourcesContent) executed by tools/glsl_interp.py, and
 *   (2) the rounding contract in spec/PARITY.md (every op is one IEEE-754 binary32
 *       operation, evaluated in GLSL source order, no FMA contraction).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (tendrils_b200/) never does.
 *
 * Layouts
 *   particles / targets : float[PW*PH*4], x-major ("draw order"): texel (x,y) at
 *                         index p = x*PH + y  -- the layout of Particles.pixels
 *                         (reference src/particles.js:76-78,94-113).
 *   flow grid           : float[W*H*4], row-major GL order: texel (gx,gy) at gy*W+gx.
 *   spawn image         : float[IW*IH*4], row-major, texel row 0 first.
 */
#ifndef TENDRILS_ORACLE_H
#define TENDRILS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Uniform block of logic.frag (reference src/logic.frag:9-34, defaults src/index.js:29-57). */
typedef struct or_params {
    float damping, speedLimit;
    float forceWeight, varyForce;
    float flowWeight, varyFlow;
    float noiseWeight, varyNoise;
    float flowDecay, flowWidth;
    float noiseScale, varyNoiseScale;
    float noiseSpeed, varyNoiseSpeed;
    float target, varyTarget;
    float viewSize[2];
} or_params;

/* Uniforms of the pixel spawners (reference src/spawn/pixels/index.js:49-58, frag/head.frag:3-15). */
typedef struct or_spawn_pixels {
    float spawnSize[2];
    float jitter[2];
    float speed;
    float bias;
    float spawnMatrix[9];   /* column-major mat3, as gl-matrix */
    float flowDecay;        /* only used by the flow-sample variant */
} or_spawn_pixels;

enum { OR_APPLY_COLOR = 0, OR_APPLY_BRIGHTEST = 1, OR_APPLY_IDENTITY = 2, OR_APPLY_FLOW = 3 };

/* scalar helpers exported for unit tests */
float or_sin(float x);
float or_cos(float x);
float or_random(float cx, float cy);
float or_snoise3(float x, float y, float z);

/* a3: logic.frag main() over columns [x0,x1) of a PW x PH particle texture. */
void or_integrate(const or_params *P, int PW, int PH, int x0, int x1,
                  const float *state_in, float *state_out, const float *targets,
                  const float *flow, int W, int H, float time, float dt);

/* a8/a9 (D6): vertex -> (texel row, current?) table of the 2*PH vertices of one column. */
void or_vertex_table(int PH, int *row_of_vertex, int *cur_of_vertex);
/* column index sampled by vertex column i (reference particles.js:171-190 + NEAREST fetch) */
void or_column_table(int PW, int *col_of_vertex_col);

/* the vertex stage of the flow draw alone: out6 = (gl_Position.xy, color.rgba); 0 when the vertex is inert */
int or_flow_vertex(const or_params *P, int PW, int PH, int i, int j, const float *cur, const float *prev,
                   float time, float *out6);

/* a7-a10: flow/index.vert + GL_LINES raster + ordered alpha-over blend, columns [x0,x1). */
/* returns the number of fragments blended. */
long long or_splat(const or_params *P, int PW, int PH, int x0, int x1,
                   const float *cur, const float *prev, float *flow, int W, int H, float time);

/* same result as or_splat, multi-threaded (used by the timed CPU baseline) */
long long or_splat_mt(const or_params *P, int PW, int PH, int x0, int x1,
                      const float *cur, const float *prev, float *flow, int W, int H, float time);

/* a12-a15 spawners; out may alias nothing. columns [x0,x1). */
void or_spawn_init(int PW, int PH, int x0, int x1, float *out);
void or_spawn_ball(int PW, int PH, int x0, int x1, float radius, float speed, float *out);
void or_spawn_pixels_direct(const or_spawn_pixels *S, int PW, int PH, int x0, int x1,
                            const float *image, int IW, int IH, float time, float *out);
void or_spawn_pixels_sample(const or_spawn_pixels *S, int apply, int vignette, int samples,
                            int PW, int PH, int x0, int x1, const float *state_in,
                            const float *image, int IW, int IH, float time, float *out);

/* f1: optical-flow pass blended into the flow grid (reference src/optical-flow/index.frag); view/last RGBA8 */
void or_optical_flow(const float viewSize[2], const float scaleUV[2], float offset, float lambda, float speed,
                     float speedLimit, float time, const unsigned char *view, const unsigned char *last, int iw, int ih,
                     float *flow, int W, int H);

/* f4: pointer flow lines (reference src/flow-line/index.{vert,frag} over the strip of src/geom/line/index.js).
 * Uniforms of the Line (src/flow-line/index.js:19-22, src/geom/line/index.js:16-20, state merged in at
 * src/demo.main.js:1118). */
typedef struct or_flow_line_uniforms {
    float viewSize[2];
    float rad, speed, speedLimit, crestShape;
} or_flow_line_uniforms;
/* the vertex stage alone: out9 = (gl_Position.xy, values.rgba, crest.xy, sdf) */
void or_flow_line_vertex(const or_flow_line_uniforms *U, const float position[2], const float normal[2], float miter,
                         const float previous[2], float time, float dt, float *out9);
/* the fragment stage alone: in7 = (values.rgba, crest.xy, sdf) -> gl_FragColor */
void or_flow_line_fragment(float crestShape, const float *in7, float *rgba);
/* the whole draw: vertex stage, TRIANGLE_STRIP raster (spec/PARITY.md FL3-FL5), fragment stage, ordered
 * alpha-over blend into flow [H][W][4]; attribute arrays as gl-geometry holds them.  Returns the fragment count. */
long long or_flow_line(const or_flow_line_uniforms *U, int n_vertices, const float *position, const float *normal,
                       const float *miter, const float *previous, const float *time, const float *dt,
                       float *flow, int W, int H);

int or_num_threads(void);
void or_set_threads(int n);     /* OpenMP threads of the parallel entry points (benchmarks: all host cores) */

#ifdef __cplusplus
}
#endif
#endif
