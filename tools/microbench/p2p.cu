// p2p.cu -- NVLink peer bandwidth of device-initiated LOADS vs STORES between GPU 0 and GPU 1, by access size
// and bytes in flight per thread.  Round 1 measured (inside the band fold) ~0.1 TB/s for a gather that pulls
// fragments with remote loads and ~0.6 TB/s for the same bytes pushed with remote stores; this standalone program
// is for charting that in isolation before tuning k_bands_push further.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o p2p tools/microbench/p2p.cu && ./p2p      (needs 2 GPUs)
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <int UNROLL, bool PULL>
__global__ void copy_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, size_t n) {
    // PULL: src is remote (loads cross NVLink); otherwise dst is remote (stores cross NVLink)
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = src[i + u * stride];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) dst[i + u * stride] = v[u];
    }
}

template <int UNROLL, bool PULL>
static int run(const char *name, const float4 *src, float4 *dst, size_t n, int blocks) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int w = 0; w < 2; ++w) copy_kernel<UNROLL, PULL><<<blocks, 256>>>(src, dst, n);
    CK(cudaEventRecord(a));
    const int reps = 10;
    for (int r = 0; r < reps; ++r) copy_kernel<UNROLL, PULL><<<blocks, 256>>>(src, dst, n);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    std::printf("%-34s unroll %d  blocks %5d : %8.1f GB/s\n", name, UNROLL, blocks, n * 16.0 * reps / (ms * 1e-3) / 1e9);
    return 0;
}

int main() {
    int n_dev = 0;
    CK(cudaGetDeviceCount(&n_dev));
    if (n_dev < 2) { std::printf("needs 2 GPUs\n"); return 0; }
    const size_t n = (256u << 20) / 16;                      // 256 MiB
    float4 *local = nullptr, *remote = nullptr;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&remote, n * 16)); CK(cudaMemset(remote, 1, n * 16));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&local, n * 16)); CK(cudaMemset(local, 2, n * 16));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    for (int blocks : {148, 148 * 4, 148 * 16}) {
        if (run<1, true>("pull (remote loads, local stores)", remote, local, n, blocks)) return 1;
        if (run<4, true>("pull (remote loads, local stores)", remote, local, n, blocks)) return 1;
        if (run<8, true>("pull (remote loads, local stores)", remote, local, n, blocks)) return 1;
        if (run<1, false>("push (local loads, remote stores)", local, remote, n, blocks)) return 1;
        if (run<4, false>("push (local loads, remote stores)", local, remote, n, blocks)) return 1;
    }
    return 0;
}
