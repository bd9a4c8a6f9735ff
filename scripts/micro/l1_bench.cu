// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/l1_bench scripts/micro/l1_bench.cu
// How much L1 time does a 16-byte-per-lane gather cost when a warp's request touches 1, 2 or 4
// different 128-byte lines (all L1 hits) — and the same gather from shared memory?
// The loop body is loads only (addresses precomputed), 16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 ldg16(const char *p) { float4 r; asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ float2 ldg8(const char *p) { float2 r; asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p)); return r; }
__device__ __forceinline__ float ldg4(const char *p) { float r; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p)); return r; }
__device__ __forceinline__ float4 lds16(unsigned a) { float4 r; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a)); return r; }
// MODE 0: LDG.128 4 lines/request, 1: LDG.128 1 line, 2: LDG.64 2 lines, 3: LDS.128 4 rows, 4: LDG.32 1 line (128 B per request)
template <int MODE>
__global__ void __launch_bounds__(512) k(const char *buf, int npix, int iters, float *sink)
{
    extern __shared__ __align__(128) char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (MODE == 3) { for (int i = threadIdx.x; i < npix * 32; i += blockDim.x) reinterpret_cast<float *>(sm)[i] = (float)i; __syncthreads(); }
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    unsigned off[8];
    unsigned s = warp * 2654435761u + blockIdx.x * 97u;
    for (int u = 0; u < 8; ++u) {
        s = s * 1664525u + 1013904223u;
        const int grp = (MODE == 2) ? lane >> 4 : (MODE == 4 ? 0 : lane >> 3);
        const unsigned pix = (MODE == 1) ? (s >> 8) % npix : ((s >> 8) + grp * 37u) % npix;
        off[u] = pix * 128 + (MODE == 2 ? (lane & 15) * 8 : MODE == 4 ? lane * 4 : (lane & 7) * 16);
    }
    const unsigned smbase = (unsigned)__cvta_generic_to_shared(sm);
    const unsigned mask = (unsigned)npix * 128u - 1u;      // npix is a power of two
    for (int it = 0; it < iters; ++it) {
        const unsigned shift = (unsigned)it * 128u * 5u;       // every iteration reads other pixels (addresses depend on it)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned o = (off[u] + shift) & mask;
            if (MODE == 2) { const float2 t = ldg8(buf + o); a0 += t.x; a1 += t.y; }
            else if (MODE == 3) { const float4 t = lds16(smbase + o); a0 += t.x; a1 += t.y; a2 += t.z; a3 += t.w; }
            else if (MODE == 4) { a0 += ldg4(buf + o); }
            else { const float4 t = ldg16(buf + o); a0 += t.x; a1 += t.y; a2 += t.z; a3 += t.w; }
        }
    }
    if (a0 + a1 + a2 + a3 == 1.2345f) *sink = a0;
}
int main()
{
    const int npix = 512;   // 64 KB: L1 / shared resident
    char *buf; cudaMalloc(&buf, (size_t)npix * 128); cudaMemset(buf, 0, (size_t)npix * 128);
    float *sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000, blocks = 148, threads = 512;
    const char *names[] = {"LDG.128, 4 lines per request", "LDG.128, 1 line per request", "LDG.64, 2 lines per request", "LDS.128, 4 rows per request", "LDG.32, 1 line per request"};
    for (int mode = 0; mode < 5; ++mode) {
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, threads, 0>>>(buf, npix, iters, sink);
            if (mode == 1) k<1><<<blocks, threads, 0>>>(buf, npix, iters, sink);
            if (mode == 2) k<2><<<blocks, threads, 0>>>(buf, npix, iters, sink);
            if (mode == 3) { cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, npix * 128); k<3><<<blocks, threads, npix * 128>>>(buf, npix, iters, sink); }
            if (mode == 4) k<4><<<blocks, threads, 0>>>(buf, npix, iters, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const double req = (double)(threads / 32) * iters * 8;   // requests per SM
        printf("%-32s %.3f ms  %.2f cycles per warp request per SM\n", names[mode], ms, ms * 1e-3 * 1.965e9 / req);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
