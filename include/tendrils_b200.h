/*
 * tendrils_b200.h -- C ABI of the B200-native Tendrils particle step.
 *
 * Drop-in boundary for ONE path of keeffEoghan/tendrils: what Tendrils.step(), the flow half
 * of Tendrils.draw(), Tendrils.spawn()/spawnShader() and the built-in spawners make the GPU
 * do (reference src/index.js:248-303,425-457, src/particles.js:94-158).  The reference has no
 * FFI of its own (it is JavaScript over WebGL); a Node N-API addon (bindings/node/) or any
 * other FFI binds exactly these symbols -- see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, a negative tb_status otherwise; the message is
 *     available from tb_last_error() (the reference throws JS Errors from stack.gl; the
 *     N-API shim turns a non-zero status into a thrown Error).
 *   - host pointers are borrowed for the duration of the call only.
 *   - calls are asynchronous on the context's CUDA stream unless they return data to the host.
 *   - a context is not re-entrant (the reference runs on the single JS thread).
 *   - there is NO CPU fallback: every entry point needs a CUDA device.
 *
 * Layouts (host side)
 *   particle state / targets : float[PW*PH*4] "x-major": texel (x,y) at ((x*PH)+y)*4 -- the
 *       layout of the reference's CPU mirror Particles.pixels, an ndarray of shape
 *       [PW,PH,4] (src/particles.js:76-78,94-113).  RGBA = (pos.x, pos.y, vel.x, vel.y).
 *       A context that owns columns [col0,col1) transfers only those columns.
 *   flow grid   : float[W*H*4], GL readPixels order: texel (gx,gy) at ((gy*W)+gx)*4;
 *       RGBA = (vel.x, vel.y, time stamp [ms], alpha)  (src/flow/apply/state.glsl:5-16).
 *   spawn image : float[IW*IH*4], texel row 0 first (what texImage2D stores).
 */
#ifndef TENDRILS_B200_H
#define TENDRILS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB_ABI_VERSION 2

typedef struct tb_ctx tb_ctx;

typedef enum tb_status {
    TB_OK = 0,
    TB_ERR_INVALID = -1,      /* bad argument                                   */
    TB_ERR_CUDA = -2,         /* CUDA runtime error (no device, OOM, launch)    */
    TB_ERR_UNSUPPORTED = -3,  /* behaviour outside the built-in shaders         */
    TB_ERR_OVERFLOW = -4      /* internal capacity exceeded and not recoverable */
} tb_status;

/* Replaces: new Tendrils(gl, options) + setupParticles(rootNum) + resize()
 * (src/index.js:84-147,186-210,393-408) and new Particles(gl,{shape,geomShape})
 * (src/particles.js:43-92).  geomShape is always (PW, 2*PH) (src/index.js:193-197). */
typedef struct tb_config {
    int32_t particles_w;      /* PW: columns of the particle texture (rootNum)          */
    int32_t particles_h;      /* PH: rows of the particle texture (rootNum)             */
    int32_t col0, col1;       /* columns owned by this context; col1 == 0 means PW      */
    int32_t flow_w, flow_h;   /* flow grid = viewRes (src/index.js:394-405)             */
    int32_t device;           /* CUDA device ordinal                                    */
    int32_t flags;            /* reserved, 0                                            */
} tb_config;

/* The step-relevant part of Tendrils.state (src/index.js:29-57) plus viewSize
 * (src/index.js:397-398).  Copied on tb_set_state, i.e. "read at step time" (src/index.js:255). */
typedef struct tb_state {
    float damping, speedLimit;
    float forceWeight, varyForce;
    float flowWeight, varyFlow;
    float noiseWeight, varyNoise;
    float flowDecay, flowWidth;
    float noiseScale, varyNoiseScale;
    float noiseSpeed, varyNoiseSpeed;
    float target, varyTarget;
    float viewSize[2];
} tb_state;

/* Uniforms of PixelSpawner.update (src/spawn/pixels/index.js:49-58). */
typedef struct tb_pixel_spawner {
    float spawnSize[2];
    float jitter[2];          /* jitterRad / viewRes                                    */
    float speed;
    float bias;
    float spawnMatrix[9];     /* column-major mat3 (gl-matrix)                          */
} tb_pixel_spawner;

/* Uniforms of the optical-flow pass (src/optical-flow/index.js:21-30, src/demo.main.js:1148-1153). */
typedef struct tb_optical_flow_params {
    float viewSize[2];
    float scaleUV[2];
    float offset, lambda, speed, speedLimit, time;
} tb_optical_flow_params;

/* Uniforms of a FlowLine's Line (src/flow-line/index.js:19-22, src/geom/line/index.js:16-20; the app merges
 * tendrils.state in before every draw, src/demo.main.js:1118, which is where speedLimit comes from). */
typedef struct tb_flow_line_params {
    float viewSize[2];
    float rad, speed, speedLimit, crestShape;
} tb_flow_line_params;

/* Where a spawn pass writes: spawnShader(shader, update) vs spawnShader(shader, update, tendrils.targets)
 * (src/index.js:432-457, src/particles.js:123-130). */
typedef enum tb_target { TB_TARGET_STATE = 0, TB_TARGET_TARGETS = 1 } tb_target;

/* Built-in pixel spawn shaders (src/spawn/pixels/ *.frag). */
typedef enum tb_spawn_variant {
    TB_SPAWN_DIRECT = 0,        /* index.frag -> frag/direct-main.frag            */
    TB_SPAWN_BEST_SAMPLE = 1,   /* best-sample.frag   (colour o vignette, 6)      */
    TB_SPAWN_BRIGHT_SAMPLE = 2, /* bright-sample.frag (brightest, 6)              */
    TB_SPAWN_COLOR_SAMPLE = 3,  /* color-sample.frag  (colour, 3)                 */
    TB_SPAWN_DATA_SAMPLE = 4,   /* data-sample.frag   (identity o vignette, 2)    */
    TB_SPAWN_FLOW_SAMPLE = 5    /* flow-sample.frag   (flow decode, 5)            */
} tb_spawn_variant;

/* What `spawnData` is bound to (src/spawn/pixels/index.js:51, src/demo.main.js:401-441). */
typedef enum tb_spawn_source {
    TB_SOURCE_IMAGE = 0,        /* the image last given to tb_set_spawn_image     */
    TB_SOURCE_FLOW = 1,         /* the flow grid (spawnFlow)                      */
    TB_SOURCE_PARTICLES = 2     /* the current particle texture (spawnFastest)    */
} tb_spawn_source;

typedef enum tb_buffer {
    TB_BUF_CURRENT = 0,         /* particles.buffers[0]                           */
    TB_BUF_PREVIOUS = 1,        /* particles.buffers[1]                           */
    TB_BUF_TARGETS = 2,         /* tendrils.targets                               */
    TB_BUF_FLOW = 3             /* tendrils.flow                                  */
} tb_buffer;

int tb_abi_version(void);
const char *tb_last_error(const tb_ctx *ctx);          /* ctx may be NULL: last create error */

/* lifetime -- replaces new Tendrils/Particles, Particles.setup(2), dispose() */
int tb_create(const tb_config *cfg, tb_ctx **out);
int tb_destroy(tb_ctx *ctx);

/* Tendrils.state + viewSize, read at step/draw time (src/index.js:255-263,284-293) */
int tb_set_state(tb_ctx *ctx, const tb_state *state);

/* Scheduling switch, no effect on results: with `on` (the default) tb_step evaluates the two simplex noises of logic.frag, which
 * do not read the flow grid, on a low-priority side stream under the previous flow splat and finishes the shader on the main
 * stream.  Off: one fused launch (what a roofline measurement of the kernel wants). */
int tb_set_overlap(tb_ctx *ctx, int32_t on);

/* Tendrils.resize(): flow.shape = viewRes, which reallocates and zeroes (src/index.js:393-408) */
int tb_resize_flow(tb_ctx *ctx, int32_t w, int32_t h);
/* Tendrils.clearFlow() (src/index.js:234-239) */
int tb_clear_flow(tb_ctx *ctx);

/* Tendrils.step() -> Particles.step(): rotate the ping-pong pair, run logic.frag
 * (src/index.js:248-272, src/particles.js:123-145, src/logic.frag:45-101).
 * time and dt in milliseconds (Timer, src/timer.js:24-60). */
int tb_step(tb_ctx *ctx, float time, float dt);

/* tb_upload(TB_BUF_CURRENT, host_in) + tb_step + tb_download(TB_BUF_CURRENT, host_out) as ONE pipelined pass for callers that
 * keep the particle state on the host (the reference's CPU mirror Particles.pixels, src/particles.js:76-78,94-117): the
 * columns go up in n_chunks blocks (1..64) on a copy stream, through the logic pass, and down on a second copy stream, so
 * that PCIe runs in both directions at once and under the kernels.  ASYNCHRONOUS: returns when everything is queued;
 * host_in must stay untouched and host_out is complete only after tb_sync (or the next call that reads state back).  Both
 * should be pinned.  host_in == host_out (the state round-trips through one buffer) is allowed: the upload of a chunk then
 * follows the previous step's download of that chunk. */
int tb_step_streamed(tb_ctx *ctx, float time, float dt, const float *host_in, float *host_out, int32_t n_chunks);

/* The flow half of Tendrils.draw(): particles.draw(LINES) with the flow shader into the flow
 * FBO under SRC_ALPHA/ONE_MINUS_SRC_ALPHA blending (src/index.js:278-303,267-268,
 * src/particles.js:147-158, src/flow/index.vert, src/flow/index.frag). */
int tb_splat_flow(tb_ctx *ctx, float time);

/* Split form of tb_splat_flow for column-sharded contexts: tb_splat_collect rasterises this
 * context's primitives into per-tile fragment bins (draw order inside a bin); tb_splat_fold blends
 * them onto the flow grid currently in the context.  Folding rank 0..P-1 in turn onto one grid
 * equals the single-context result bit for bit (primitive order = column order). */
int tb_splat_collect(tb_ctx *ctx, float time);
int tb_splat_fold(tb_ctx *ctx);

/* The flow half of draw() for a column-sharded run over peer memory (one process per GPU, one node, at most 16 ranks).
 * The fragment bins are owned round-robin by the ranks.  Every rank exports IPC handles of its bin array / flow grid /
 * totals table / flags (tb_owners_export; `reserve_fragments` fixes the capacity of the bin array, which the other ranks
 * map -- a draw that needs more fails with TB_ERR_OVERFLOW on every rank alike), the host layer all-gathers the blobs, and
 * each rank maps all of them (tb_owners_connect, `all_handles` = world x tb_owners_handle_bytes(), rank order).  Then
 * tb_splat_flow_owners replaces tb_splat_flow: count, exchange the per-bin totals, rasterise straight into the owners' bins
 * over NVLink (per bin the ranks side by side in rank order = column order = the reference's primitive order,
 * src/particles.js:182-186), fold the bins this rank owns, store the finished texels into every rank's grid; three
 * all-rank barriers per draw, all inside the CUDA stream.  Equal to the single-context draw bit for bit.  Must be redone
 * after tb_resize_flow. */
int64_t tb_owners_handle_bytes(void);
int tb_owners_export(tb_ctx *ctx, int64_t reserve_fragments, void *handles_out, int64_t n_bytes);
int tb_owners_connect(tb_ctx *ctx, int32_t rank, int32_t world, const void *all_handles, int64_t n_bytes);
int tb_splat_flow_owners(tb_ctx *ctx, float time);

/* Tendrils.spawn(cpuFn) with the default initSpawner: fills ALL buffers
 * (src/index.js:425-429, src/particles.js:94-117, src/spawn/init/cpu.js:3-8). */
int tb_reset(tb_ctx *ctx);

/* Tendrils.spawnShader(shader, update, buffer) with a built-in shader.  The caller ticks the
 * timer first (src/index.js:433) and passes the resulting time. */
int tb_spawn_init(tb_ctx *ctx, tb_target target);                               /* spawn/init/index.frag */
int tb_spawn_ball(tb_ctx *ctx, float radius, float speed, tb_target target);    /* spawn/ball/index.frag */
/* `rgba` may be a host pointer (borrowed for the call) or a pointer to memory of this device (unified
 * addressing; copied in stream order on tb_stream, the caller keeps it unchanged until then). */
int tb_set_spawn_image(tb_ctx *ctx, const float *rgba, int32_t w, int32_t h);   /* PixelSpawner.setPixels */
int tb_spawn_pixels(tb_ctx *ctx, const tb_pixel_spawner *params, tb_spawn_variant variant,
                    tb_spawn_source source, float time, tb_target target);

/* texture.setPixels / readPixels equivalents; also the checkpoint and parity taps.
 * For TB_BUF_FLOW n_floats = 4*W*H, otherwise 4*(col1-col0)*PH. */
int tb_upload(tb_ctx *ctx, tb_buffer which, const float *host, int64_t n_floats);
int tb_download(tb_ctx *ctx, tb_buffer which, float *host, int64_t n_floats);

/* Alpha-over blend of a caller-rendered RGBA layer into the flow grid (hook for the L4
 * inputs that draw into the flow FBO: optical flow, pointer flow-lines; src/demo.main.js:1107-1159). */
int tb_blend_into_flow(tb_ctx *ctx, const float *rgba, int32_t w, int32_t h);

/* diagnostic tap: the fragment bins left by the last splat.  offsets: tb_debug_max_bins() + 1 words (bin b holds
 * offsets[b+1] - offsets[b] fragments, b < *n_bins); info: tb_debug_max_bins() words, bin -> strip | sub << 16 |
 * log2(bins of the strip) << 24; the strip size in texels.  Used by tools/ to study the load distribution of the fold. */
int tb_debug_max_bins(void);
int tb_debug_bins(tb_ctx *ctx, uint32_t *offsets, uint32_t *info, int32_t *n_bins, int32_t *strip_w, int32_t *strip_h);
/* diagnostic tap: how many bins the last splat folded in segments (spec/PARITY.md B4), in how many segments, and how many
 * fragments the segments left on record for the join. */
int tb_debug_segments(tb_ctx *ctx, int32_t *bins, int32_t *segments, int64_t *records);

/* OpticalFlow.update() + screen.render() with the flow FBO bound (src/optical-flow/index.frag:55-81,
 * src/optical-flow/index.js:50-58, call site src/demo.main.js:1131-1156): the gradient optical flow of two RGBA8
 * frames (view = current, last = previous; w*h*4 bytes each, texel row 0 first), written in the flow encoding
 * and alpha-over blended into the flow grid.  The frames may be host pointers (borrowed for the call) or both
 * device pointers (e.g. decoded video frames; copied in stream order, no host synchronisation). */
/* FlowLine.draw() with the flow FBO bound (src/demo.main.js:1107-1121): the TRIANGLE_STRIP of a pointer path,
 * two vertices per path point, run through src/flow-line/index.vert and src/flow-line/index.frag and alpha-over
 * blended into the flow grid in triangle order.  The six attribute arrays are the ones Line.update() fills
 * (src/geom/line/index.js:73-117, src/flow-line/index.js:54-69): position, normal, previous hold 2 floats per
 * vertex, miter, time, dt one.  Fewer than 3 vertices draw nothing. */
int tb_flow_line(tb_ctx *ctx, const tb_flow_line_params *params, int32_t n_vertices, const float *position,
                 const float *normal, const float *miter, const float *previous, const float *time, const float *dt);

int tb_optical_flow(tb_ctx *ctx, const tb_optical_flow_params *params, const uint8_t *view_rgba8,
                    const uint8_t *last_rgba8, int32_t w, int32_t h);

/* plumbing for the host layer (PyTorch / NCCL): raw device pointers, stream, sync, timing */
int tb_device_ptr(tb_ctx *ctx, tb_buffer which, void **ptr, int64_t *n_floats);
int tb_stream(tb_ctx *ctx, void **cuda_stream);
/* Inputs handed over as DEVICE pointers (tb_set_spawn_image, tb_optical_flow, tb_flow_line) are read in stream order on
 * tb_stream.  If another stream wrote them (a decoder, torch's current stream), call this first: the context's stream then
 * waits for everything queued so far on `producer_stream` (a cudaStream_t of the same device; NULL = the legacy default stream). */
int tb_wait_stream(tb_ctx *ctx, void *producer_stream);
int tb_sync(tb_ctx *ctx);
/* counters since creation: kernels launched by this library, fragments blended by the last splat */
int tb_stats(tb_ctx *ctx, int64_t *kernel_launches, int64_t *last_fragments);
/* CUDA-event timing: number of timed calls since the last reset (at most 512 are kept) and their
 * summed device time in milliseconds, for (1) the integrate launch on the main stream, (2) the flow
 * splat (collect start to fold end) and (3) the noise launch on the side stream when tb_step overlaps
 * it with the previous splat (its time then hides under (2)).  Synchronises.  reset != 0 restarts. */
int tb_timing(tb_ctx *ctx, int reset, int64_t *n_integrate, float *integrate_ms, int64_t *n_splat, float *splat_ms,
              int64_t *n_noise, float *noise_ms);

#ifdef __cplusplus
}
#endif
#endif
