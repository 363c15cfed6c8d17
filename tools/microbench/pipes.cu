// Microbenchmark of the sm_100a pipes the integrate kernel leans on: scalar FMUL/FADD vs packed
// FFMA2 issue cost, FRND (floor), FSETP+FSEL.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a),"l"(b),"l"(c)); return r; }

constexpr int ITER = 2048;
__global__ void k_scalar(float* out, float a, float b) {   // 16 FP ops / iter (8 FMUL + 8 FADD), 8 chains
    float x[8];
    for (int j = 0; j < 8; ++j) x[j] = threadIdx.x * 1e-3f + j;
    for (int i = 0; i < ITER; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = __fadd_rn(__fmul_rn(x[j], a), b);
    float s = 0; for (int j = 0; j < 8; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float a, float b, float one, float nz) {   // same 16 scalar-equivalents as 8 FFMA2
    u64 x[4];
    for (int j = 0; j < 4; ++j) x[j] = pk(threadIdx.x * 1e-3f + 2 * j, threadIdx.x * 1e-3f + 2 * j + 1);
    const u64 A = pk(a, a), B = pk(b, b), ONE = pk(one, one), NZ = pk(nz, nz);
    for (int i = 0; i < ITER; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = fma2(fma2(x[j], A, NZ), ONE, B);
    float s = 0; for (int j = 0; j < 4; ++j) { float p, q; up(x[j], p, q); s += p + q; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_frnd(float* out, float a, float b) {     // 8 FRND + 8 FADD per iter
    float x[8];
    for (int j = 0; j < 8; ++j) x[j] = threadIdx.x * 1e-3f + j;
    for (int i = 0; i < ITER; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = __fadd_rn(floorf(x[j]), b);
    float s = 0; for (int j = 0; j < 8; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_magicfloor(float* out, float a, float b) {   // floor by magic add: 3 FADD + FSETP + FADD(b)
    float x[8];
    for (int j = 0; j < 8; ++j) x[j] = threadIdx.x * 1e-3f + j;
    for (int i = 0; i < ITER; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float r = __fsub_rn(__fadd_rn(x[j], 12582912.0f), 12582912.0f);
            if (r > x[j]) r = __fsub_rn(r, 1.0f);
            x[j] = __fadd_rn(r, b);
        }
    float s = 0; for (int j = 0; j < 8; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_sel(float* out, float a, float b) {      // 8 x (FSETP + FSEL + FADD)
    float x[8];
    for (int j = 0; j < 8; ++j) x[j] = threadIdx.x * 1e-3f + j;
    for (int i = 0; i < ITER; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = __fadd_rn((x[j] < a) ? b : x[j], b);
    float s = 0; for (int j = 0; j < 8; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mixed(float* out, float a, float b, float one, float nz) {  // packed FP + scalar FRND interleaved
    u64 x[4]; float y[4];
    for (int j = 0; j < 4; ++j) { x[j] = pk(threadIdx.x * 1e-3f + 2 * j, threadIdx.x * 1e-3f + 2 * j + 1); y[j] = j + threadIdx.x * 1e-3f; }
    const u64 A = pk(a, a), B = pk(b, b), ONE = pk(one, one), NZ = pk(nz, nz);
    for (int i = 0; i < ITER; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { x[j] = fma2(fma2(x[j], A, NZ), ONE, B); y[j] = __fadd_rn(floorf(y[j]), b); }
    float s = 0; for (int j = 0; j < 4; ++j) { float p, q; up(x[j], p, q); s += p + q + y[j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int r = 0; r < 5; ++r) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8, threads = 256;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    const double warps = double(blocks) * threads / 32;
    auto rep = [&](const char* name, float ms, double warp_instr_per_iter_note, double scalar_ops_per_iter) {
        double clkcycles = ms * 1e-3 * clk * 1e3;
        printf("%-12s %8.3f ms  cycles/iter/warp-per-SMSP %.2f  scalar-op-equiv/clk/SM %.1f\n", name, ms,
               clkcycles / ITER / (warps / sms / 4), scalar_ops_per_iter * 32 * warps * ITER / clkcycles / sms);
        (void)warp_instr_per_iter_note;
    };
    rep("scalar", timeit([&] { k_scalar<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16, 16);
    rep("packed", timeit([&] { k_packed<<<blocks, threads>>>(out, 1.0001f, 0.5f, 1.0f, -0.0f); }), 8, 16);
    rep("frnd", timeit([&] { k_frnd<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16, 16);
    rep("magicfloor", timeit([&] { k_magicfloor<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 40, 16);
    rep("sel", timeit([&] { k_sel<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 24, 24);
    rep("mixed", timeit([&] { k_mixed<<<blocks, threads>>>(out, 1.0001f, 0.5f, 1.0f, -0.0f); }), 16, 24);
    printf("SMs %d clock %d kHz\n", sms, clk);
    return 0;
}
