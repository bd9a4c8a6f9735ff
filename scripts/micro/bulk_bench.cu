// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/bulk_bench scripts/micro/bulk_bench.cu
// Microbenchmark for the staged unprojection kernel: how fast can a CTA fill shared-memory
// texel patches with cp.async.bulk (UBLKCP) when every copy is ONE pixel (PB bytes, padded
// destination stride) versus one copy per patch ROW; and the LDS.128 gather rate from the
// padded patch with one voxel per lane.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// MODE 0: one copy per pixel (PB bytes -> stride PB+16); MODE 1: one copy per patch row (dense)
template <int MODE>
__global__ void __launch_bounds__(256) fill_kernel(const char *planes, int Wp, int Hp, int nplanes, int pw, int ph, int V, int PB,
                                                   int rounds, unsigned long long *sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar;
    __shared__ unsigned long long bars[8];
    const int PS = PB + 16;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int v = 0; v < 8; ++v) mbar_init(&bars[v], 1); }
    __syncthreads();
    unsigned acc = 0;
    const size_t plane_bytes = (size_t)Wp * Hp * PB;
    for (int r = 0; r < rounds; ++r) {
        const unsigned seed = (blockIdx.x * 7919u + r * 104729u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (MODE <= 1 && threadIdx.x == 0) mbar_expect(&bar, (unsigned)(V * pw * ph * PB));
        for (int v = 0; v < V; ++v) {
            const unsigned s = seed + v * 31u;
            const int x0 = s % (Wp - pw), y0 = (s / 97u) % (Hp - ph);
            const char *pl = planes + (size_t)((s / 7u) % nplanes) * plane_bytes;
            unsigned char *dstv = smem + (size_t)v * pw * ph * PS;
            if (MODE == 0) {
                for (int j = threadIdx.x; j < pw * ph; j += blockDim.x) {
                    const int rr = j / pw, cc = j - rr * pw;
                    bulk_g2s(dstv + (size_t)j * PS, pl + ((size_t)(y0 + rr) * Wp + x0 + cc) * PB, PB, &bar);
                }
            } else if (MODE == 4) {      // per row, one mbarrier per view, issuing threads spread over warps
                if (threadIdx.x == 32 * v) mbar_expect(&bars[v], (unsigned)(pw * ph * PB));
                if ((threadIdx.x >> 5) == v && (threadIdx.x & 31) < ph) {
                    const int rr = threadIdx.x & 31;
                    bulk_g2s(dstv + (size_t)rr * pw * PB, pl + ((size_t)(y0 + rr) * Wp + x0) * PB, pw * PB, &bars[v]);
                }
            } else if (MODE == 5) {      // per pixel, one mbarrier per view
                if (threadIdx.x == 0) mbar_expect(&bars[v], (unsigned)(pw * ph * PB));
                for (int j = threadIdx.x; j < pw * ph; j += blockDim.x) {
                    const int rr = j / pw, cc = j - rr * pw;
                    bulk_g2s(dstv + (size_t)j * PS, pl + ((size_t)(y0 + rr) * Wp + x0 + cc) * PB, PB, &bars[v]);
                }
            } else if (MODE == 1) {
                for (int rr = threadIdx.x; rr < ph; rr += blockDim.x)
                    bulk_g2s(dstv + (size_t)rr * pw * PB, pl + ((size_t)(y0 + rr) * Wp + x0) * PB, pw * PB, &bar);
            } else {
                // 16 bytes per lane: PB/16 lanes per pixel, padded destination
                const int lpp = PB / 16;
                for (int j = threadIdx.x; j < pw * ph * lpp; j += blockDim.x) {
                    const int pxl = j / lpp, ck = j - pxl * lpp;
                    const int rr = pxl / pw, cc = pxl - rr * pw;
                    const char *src = pl + ((size_t)(y0 + rr) * Wp + x0 + cc) * PB + ck * 16;
                    unsigned char *dst = dstv + (size_t)pxl * PS + ck * 16;
                    if (MODE == 2)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
                    else
                        *reinterpret_cast<uint4 *>(dst) = __ldg(reinterpret_cast<const uint4 *>(src));
                }
            }
        }
        if (MODE == 2) { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); __syncthreads(); }
        if (MODE == 3) __syncthreads();
        if (MODE <= 1) mbar_wait(&bar, r & 1);
        if (MODE >= 4) for (int v = 0; v < V; ++v) mbar_wait(&bars[v], r & 1);
        acc += *reinterpret_cast<unsigned *>(smem + ((threadIdx.x * 16) % (V * pw * ph * PB)));
        __syncthreads();
    }
    if (acc == 0xdeadbeef) *sink = acc;
}

// LDS.128 gather, one voxel per lane, 4 corners x NCH chunks, padded pixel stride
__global__ void __launch_bounds__(256) gather_kernel(int pw, int ph, float stepx, float stepy, int iters, float *sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int PS = 144;
    for (int i = threadIdx.x; i < pw * ph * PS / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int it = 0; it < iters; ++it) {
        const float fx = (lane & 7) * stepx + ((lane >> 3) & 3) * 0.1f + (it & 3) * 0.3f, fy = (lane & 7) * stepy + (lane >> 3) * 1.1f;
        const int x0 = min((int)fx, pw - 2), y0 = min((int)fy, ph - 2);
        const unsigned char *b0 = smem + (y0 * pw + x0) * PS, *b1 = b0 + pw * PS;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 t0 = *reinterpret_cast<const float4 *>(b0 + 16 * k), t1 = *reinterpret_cast<const float4 *>(b0 + 16 * k + PS);
            const float4 t2 = *reinterpret_cast<const float4 *>(b1 + 16 * k), t3 = *reinterpret_cast<const float4 *>(b1 + 16 * k + PS);
            a0 += t0.x + t1.x + t2.x + t3.x; a1 += t0.y + t1.y + t2.y + t3.y;
            a2 += t0.z + t1.z + t2.z + t3.z; a3 += t0.w + t1.w + t2.w + t3.w;
        }
    }
    if (a0 + a1 + a2 + a3 == 12345.678f) *sink = a0;
}

int main()
{
    const int Wp = 100, Hp = 100;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    unsigned long long *sink; cudaMalloc(&sink, 8);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double clk = 1.965e9;
    for (int nplanes : {8})
    for (int PB : {128}) {
        char *planes; cudaMalloc(&planes, (size_t)Wp * Hp * PB * nplanes); cudaMemset(planes, 1, (size_t)Wp * Hp * PB * nplanes);
        for (int mode : {1, 4, 5})
            for (int ctas : {1, 2, 3}) {
                const int V = 8, pw = 12, ph = 5, rounds = 2000;
                const size_t smem = (size_t)V * pw * ph * (PB + 16);
                auto kern = mode == 0 ? fill_kernel<0> : mode == 1 ? fill_kernel<1> : mode == 2 ? fill_kernel<2> : mode == 3 ? fill_kernel<3> : mode == 4 ? fill_kernel<4> : fill_kernel<5>;
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                float ms = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0);
                    kern<<<sms * ctas, 256, smem>>>(planes, Wp, Hp, nplanes, pw, ph, V, PB, rounds, sink);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    cudaEventElapsedTime(&ms, e0, e1);
                }
                const double copies = (mode == 1 ? (double)V * ph : (double)V * pw * ph) * rounds * ctas;   // per SM
                const double bytes = (double)V * pw * ph * PB * rounds * ctas;
                printf("planes %3d PB %3d %s, %d CTA/SM: %.3f ms  -> %.1f cycles per round per CTA, %.2f cycles/copy/SM, %.1f B/clk/SM  (%s)\n", nplanes, PB,
                       mode == 0 ? "bulk per pixel" : mode == 1 ? "bulk per row  " : mode == 2 ? "cp.async 16B  " : mode == 3 ? "ldg+sts 16B   " : mode == 4 ? "bulk row 8 bar" : "bulk pix 8 bar", ctas, ms, ms * 1e-3 * clk / rounds, ms * 1e-3 * clk / copies,
                       bytes / (ms * 1e-3 * clk), cudaGetErrorString(cudaGetLastError()));
            }
        cudaFree(planes);
    }
    float *fs; cudaMalloc(&fs, 4);
    for (float sx : {0.0f})
        for (float sy : {0.0f, 1.2f}) {
            const int pw = 17, ph = 13, iters = 4000;
            const size_t smem = (size_t)pw * ph * 144;
            cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                gather_kernel<<<sms * 3, 256, smem>>>(pw, ph, sx, sy, iters, fs);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
            }
            const double lds = 3.0 * 8 * iters * 32;     // LDS.128 warp-instructions per SM
            printf("gather step (%.2f, %.2f) px/lane: %.3f ms, %.2f cycles per LDS.128 warp-instr per SM (4.0 = conflict-free)\n", sx, sy, ms,
                   ms * 1e-3 * clk / lds);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
