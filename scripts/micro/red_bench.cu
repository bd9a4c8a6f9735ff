// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/red_bench scripts/micro/red_bench.cu
// Microbenchmark: throughput of red.global.add.f32 vs red.global.add.v4.f32 on L2-resident lines,
// 8 lanes per 128-byte line (the packed-gradient scatter pattern of the backward kernel).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void red4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <int MODE>
__global__ void k(float *buf, unsigned npix, int iters)
{
    const int lane = threadIdx.x & 31, grp = lane >> 3, chunk = lane & 7;
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) / 8 * 2654435761u + 12345u + grp;
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        unsigned pix = (MODE >= 2) ? ((blockIdx.x * 977u + (threadIdx.x >> 5) * 131u + i / 4 + grp) % npix) : (s >> 8) % npix;
        float *p = buf + (size_t)pix * 32 + chunk * 4;
        if (MODE == 0 || MODE == 2) red4(p, 1.f, 2.f, 3.f, 4.f);
        else { atomicAdd(p, 1.f); atomicAdd(p + 1, 2.f); atomicAdd(p + 2, 3.f); atomicAdd(p + 3, 4.f); }
    }
}
int main()
{
    const unsigned npix = 320000;   // 41 MB
    float *buf; cudaMalloc(&buf, (size_t)npix * 128); cudaMemset(buf, 0, (size_t)npix * 128);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 256, blocks = 148 * 4, threads = 512;
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, threads>>>(buf, npix, iters);
            if (mode == 1) k<1><<<blocks, threads>>>(buf, npix, iters);
            if (mode == 2) k<2><<<blocks, threads>>>(buf, npix, iters);
            if (mode == 3) k<3><<<blocks, threads>>>(buf, npix, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double lines = (double)blocks * threads / 8 * iters;
        printf("mode %d (%s, %s): %.3f ms, %.1f G lines(128B)/s, %.2f TB/s payload\n", mode, (mode & 1) ? "4x scalar red" : "red.v4",
               mode >= 2 ? "local walk" : "random", ms, lines / ms / 1e6, lines * 128 / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
