// Fused unproject + aggregate (models/aggregation.py:20-87), shared-memory-staged design.
//
// The bound of this op is not HBM but the 128 B/clk path from L1 / shared memory into the
// register file: every voxel-channel-view needs four texels (16 B fp32) there and ~1 B from
// HBM.  The L1-gather kernel (unproject.cu) spends one lane GROUP per voxel so that a corner
// is one 128-byte line; that costs voxel records in shared memory, an output transposition
// and dependent global loads.  Here a CTA owns a BRICK of voxels (8 z x 4 y x 8 x by
// default), and
//   * phase A: one voxel per thread is projected through every view (same IEEE sequence as
//     everywhere else: make_cell); weights and the bilinear cell stay in REGISTERS;
//   * the bounding box of the brick's cells in every view is reduced (redux.sync + shared
//     atomics) and that texel patch is copied from the packed planes into shared memory with
//     one cp.async.bulk (TMA engine, UBLKCP) per patch ROW, completion on an mbarrier per
//     view.  The planes have a pixel stride of (pixel bytes + 16), so a row is one contiguous
//     run in global AND lands in shared memory with a stride that spreads neighbouring
//     pixels over all banks;
//   * phase B: one voxel per LANE, every lane reads its own four corner texels with LDS.128
//     (shared memory gathers at 16-byte granularity: a warp's 32 different pixels cost the
//     same four wavefronts as one line would), blends and fuses the views in registers
//     (Fuse2, same order as the reference) and stores 32 B-sector-coalesced runs along z
//     straight from registers: no records, no transposition tile.
// 2-3 CTAs per SM overlap one CTA's patch copies with the others' arithmetic.
// A brick whose patch does not fit (cameras very close to the grid) reads its texels
// straight from the planes in global memory with the same code — slow, but exact; the
// number of threads per voxel (1, 2, 4 or 8 — each takes a share of the channel vectors) is
// chosen on the device from the projected size of the central brick so that this stays rare.
#include <cuda_bf16.h>
#include <cstdlib>
#include "mvhmr_common.cuh"
#include "unproject_device.cuh"

namespace mvhmr {

constexpr int kSThreads = 256;
constexpr int kSWarps = kSThreads / 32;
constexpr int kSBrickZ = 8;                // voxels along z per brick: 32-byte store sectors

struct StagedParams {
    UnprojParams u;
    int cap;               // pixels per view patch
    int view_bytes;        // (2 zero pixels + cap) * pixel stride
    int off_patch;         // byte offset of the first patch in dynamic shared memory
    int force_t;           // threads per voxel (0: choose on the device)
};

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
// global -> shared bulk copy through the TMA engine (UBLKCP); completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// voxel (vx,vy,vz) of sample b: from the coord volume, or generated with coord_volume_kernel's arithmetic
__device__ __forceinline__ void voxel_xyz(const UnprojParams &p, int b, int vx, int vy, int vz, long long n,
                                          float &X, float &Y, float &Z)
{
    if (p.coord) {
        const float *xyz = p.coord + ((size_t)b * p.n_extent + (n - p.n_origin)) * 3;
        X = __ldg(xyz); Y = __ldg(xyz + 1); Z = __ldg(xyz + 2);
    } else {
        const float *c = p.centers + 3 * b, *R = p.rot + 9 * b;
        const float c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2);
        const float d0 = __fsub_rn(__fadd_rn(p.gpos[0], __fmul_rn(p.gstep[0], (float)vx)), c0);
        const float d1 = __fsub_rn(__fadd_rn(p.gpos[1], __fmul_rn(p.gstep[1], (float)vy)), c1);
        const float d2 = __fsub_rn(__fadd_rn(p.gpos[2], __fmul_rn(p.gstep[2], (float)vz)), c2);
        X = __fadd_rn(rot_row(__ldg(R), __ldg(R + 1), __ldg(R + 2), d0, d1, d2), c0);
        Y = __fadd_rn(rot_row(__ldg(R + 3), __ldg(R + 4), __ldg(R + 5), d0, d1, d2), c1);
        Z = __fadd_rn(rot_row(__ldg(R + 6), __ldg(R + 7), __ldg(R + 8), d0, d1, d2), c2);
    }
}

// Plain (not bit-exact) pixel position of a point: only used to size the bricks.
__device__ __forceinline__ bool plain_pixel(const float *P, float X, float Y, float Z, float sxs, float sys, float &ix, float &iy)
{
    const float x = X * P[0] + Y * P[1] + Z * P[2] + P[3], y = X * P[4] + Y * P[5] + Z * P[6] + P[7];
    const float w = X * P[8] + Y * P[9] + Z * P[10] + P[11];
    if (!(w > 0.0f)) return false;
    ix = x / w * sxs; iy = y / w * sys;
    return true;
}

// VMAX views, all present (V == VMAX).  NCH: 16-byte channel vectors per pixel (4 or 8).
template <int VMAX, bool BF16, int METHOD, int NCH>
__global__ void __launch_bounds__(kSThreads, VMAX <= 4 ? 3 : 2)
unproject_staged_kernel(const StagedParams q)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const UnprojParams &p = q.u;
    constexpr int NP = BF16 ? 4 : 2;                 // channel pairs per 16-byte vector
    constexpr int CPV = BF16 ? 8 : 4;                // channels per vector
    constexpr int PS = NCH * 16 + 16;                // pixel stride in the planes and in the patches
    constexpr int LN = NCH == 8 ? 3 : 2;             // log2(NCH)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem);             // [VMAX]
    int *bbox = reinterpret_cast<int *>(smem + 64);                                      // [2][VMAX][4] xmin ymin -xmax -ymax (all reduced with min)
    int *plan = reinterpret_cast<int *>(smem + 64 + 2 * VMAX * 16);                      // [VMAX] threads per voxel each view asks for
    unsigned char *patch = smem + q.off_patch;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float Vf = (float)VMAX;

    // ---- one-time set-up: barriers, zero pixels, bounding boxes ----
    if (tid < VMAX) mbar_init(&bars[tid], 1);
    for (int i = tid; i < 2 * VMAX * 4; i += kSThreads) bbox[i] = 0x7fffffff;
    for (int i = tid; i < VMAX * (2 * PS / 16); i += kSThreads) {
        const int v = i / (2 * PS / 16), k = i - v * (2 * PS / 16);
        reinterpret_cast<uint4 *>(patch + (size_t)v * q.view_bytes)[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    // ---- plan: threads per voxel T such that the central brick's patches fit with some margin ----
    // brick shapes (x,y,z): T=1 8x4x8, T=2 4x4x8, T=4 4x2x8, T=8 2x2x8
    if (tid < VMAX) {
        int need = 1;
        if (q.force_t) need = q.force_t;
        else {
            const int b = p.b0 + p.nb / 2;
            const int cx = p.x_lo + p.nx / 2, cy = p.gy / 2, cz = p.gz / 2;
            const long long yz = (long long)p.gy * p.gz;
            float P[12];
            for (int i = 0; i < 12; ++i) P[i] = __ldg(p.proj + ((size_t)b * VMAX + tid) * 12 + i);
            const float sxs = (p.Wf - 1.0f) / p.Hf, sys = (p.Hf - 1.0f) / p.Wf;     // pixel -> sampling position (a9/a10)
            float X, Y, Z, ix0, iy0;
            const long long n = ((long long)cx * p.gy + cy) * p.gz + cz;
            float dx[3] = {0, 0, 0}, dy[3] = {0, 0, 0};
            bool ok = false;
            if (n >= p.n0 && n < p.n1) {
                voxel_xyz(p, b, cx, cy, cz, n, X, Y, Z);
                ok = plain_pixel(P, X, Y, Z, sxs, sys, ix0, iy0);
            }
            if (ok) {
                const int ext[3] = {p.nx, p.gy, p.gz};
                const long long str[3] = {yz, p.gz, 1};
                for (int a = 0; a < 3; ++a) {
                    if (ext[a] < 2) continue;
                    int o[3] = {cx, cy, cz};
                    long long n2 = n + str[a];
                    o[a] += 1;
                    if (n2 >= p.n1 || (a == 0 && o[0] >= p.x_lo + p.nx) || (a == 1 && o[1] >= p.gy) || (a == 2 && o[2] >= p.gz)) {
                        n2 = n - str[a]; o[a] -= 2;
                    }
                    float ix, iy;
                    if (n2 >= p.n0 && n2 < p.n1) {
                        voxel_xyz(p, b, o[0], o[1], o[2], n2, X, Y, Z);
                        if (plain_pixel(P, X, Y, Z, sxs, sys, ix, iy)) { dx[a] = fabsf(ix - ix0); dy[a] = fabsf(iy - iy0); }
                    }
                }
                need = 8;
                const int shapes[4][3] = {{8, 4, 8}, {4, 4, 8}, {4, 2, 8}, {2, 2, 8}};
                for (int s = 3; s >= 0; --s) {
                    // extent of the brick's cells in the image (+2 texels per cell, +1 for the fractional start), 15 % margin
                    const float ex = 1.15f * ((shapes[s][0] - 1) * dx[0] + (shapes[s][1] - 1) * dx[1] + (shapes[s][2] - 1) * dx[2]) + 3.0f;
                    const float ey = 1.15f * ((shapes[s][0] - 1) * dy[0] + (shapes[s][1] - 1) * dy[1] + (shapes[s][2] - 1) * dy[2]) + 3.0f;
                    if (ex * ey <= (float)q.cap) need = 1 << s;
                }
            }
        }
        plan[tid] = need < NCH ? need : NCH;                        // at least one channel vector per thread
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    int T = 1;
#pragma unroll
    for (int v = 0; v < VMAX; ++v) T = max(T, plan[v]);
    const int tl = 31 - __clz(T);                                   // log2(T)
    // brick shape: x shrinks first, then y
    const int bxl = tl == 0 ? 3 : (tl == 3 ? 1 : 2), byl = tl <= 1 ? 2 : 1;
    const int bx = 1 << bxl, by = 1 << byl;
    const int nvl = 3 + byl + bxl;                                  // log2(voxels per brick) = 8 - tl
    const int lz = tid & 7, ly = (tid >> 3) & (by - 1), lx = (tid >> (3 + byl)) & (bx - 1);
    const int sub = tid >> nvl;                                     // which share of the channel vectors (warp-uniform)
    const int nbx = (p.nx + bx - 1) >> bxl, nby = (p.gy + by - 1) >> byl, nbz = (p.gz + kSBrickZ - 1) / kSBrickZ;
    const unsigned per_sample = (unsigned)nbx * nby * nbz;
    const unsigned nbricks = per_sample * (unsigned)p.nb;
    const unsigned prow = (unsigned)p.Wp * PS;                      // bytes of one plane row
    unsigned phase = 0;                                             // parity bit of every view's barrier

    unsigned it = 0;
    for (unsigned brick = blockIdx.x; brick < nbricks; brick += gridDim.x, ++it) {
        // brick -> (sample, x, y, z), z fastest
        unsigned t = brick;
        const int b = p.b0 + (int)(t / per_sample); t -= (unsigned)(b - p.b0) * per_sample;
        const int ibx = (int)(t / (unsigned)(nby * nbz)); t -= (unsigned)ibx * (unsigned)(nby * nbz);
        const int iby = (int)(t / (unsigned)nbz), ibz = (int)(t - (unsigned)iby * (unsigned)nbz);
        const int vx = p.x_lo + (ibx << bxl) + lx, vy = (iby << byl) + ly, vz = ibz * kSBrickZ + lz;
        const long long n = ((long long)vx * p.gy + vy) * p.gz + vz;
        const bool live = vx < p.x_lo + p.nx && vy < p.gy && vz < p.gz && n >= p.n0 && n < p.n1;
        int *bb = bbox + (it & 1) * (VMAX * 4);

        // ---- phase A: this thread's voxel through every view ----
        float4 w[VMAX];
        unsigned cell[VMAX];                                        // py << 16 | px of the nw corner; 0xffffffff: no texels needed
        {
            float X = 0.0f, Y = 0.0f, Z = 0.0f;
            if (live) voxel_xyz(p, b, vx, vy, vz, n, X, Y, Z);
            const float4 *Pb = reinterpret_cast<const float4 *>(p.proj + (size_t)b * VMAX * 12);
#pragma unroll
            for (int v = 0; v < VMAX; ++v) {
                const float4 P0 = __ldg(Pb + 3 * v), P1 = __ldg(Pb + 3 * v + 1), P2 = __ldg(Pb + 3 * v + 2);
                const ViewCell c = make_cell(P0, P1, P2, X, Y, Z, p, -1);
                const bool use = live && c.px >= 0;
                w[v] = use ? make_float4(c.w00, c.w01, c.w10, c.w11) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                cell[v] = use ? ((unsigned)c.py << 16 | (unsigned)c.px) : 0xffffffffu;
                const int big = 0x7fffffff;
                const int xmin = __reduce_min_sync(0xffffffffu, use ? c.px : big), ymin = __reduce_min_sync(0xffffffffu, use ? c.py : big);
                const int xmax = __reduce_min_sync(0xffffffffu, use ? -c.px : big), ymax = __reduce_min_sync(0xffffffffu, use ? -c.py : big);
                if (lane == 0 && xmin != big) {
                    atomicMin(&bb[4 * v + 0], xmin); atomicMin(&bb[4 * v + 1], ymin);
                    atomicMin(&bb[4 * v + 2], xmax); atomicMin(&bb[4 * v + 3], ymax);
                }
            }
        }
        __syncthreads();                                            // (1) bounding boxes complete

        // ---- patch geometry (identical in every thread), texel addresses of this voxel ----
        bool fits = true;
        unsigned b0[VMAX], b1[VMAX];
        int armed = 0;
#pragma unroll
        for (int v = 0; v < VMAX; ++v) {
            const int4 g = *reinterpret_cast<const int4 *>(&bb[4 * v]);
            const bool any = g.x != 0x7fffffff;
            const int pw = any ? -g.z - g.x + 2 : 0, ph = any ? -g.w - g.y + 2 : 0;
            fits = fits && pw * ph <= q.cap;
            armed |= (any ? 1 : 0) << v;
            const unsigned vb = (unsigned)v * (unsigned)q.view_bytes;
            if (cell[v] == 0xffffffffu) { b0[v] = vb; b1[v] = vb; }
            else {
                const int cx = (int)(cell[v] & 0xffffu) - g.x, cy = (int)(cell[v] >> 16) - g.y;
                b0[v] = vb + (unsigned)(2 + cy * pw + cx) * PS;
                b1[v] = b0[v] + (unsigned)pw * PS;
            }
        }
        // the other parity's boxes are free again (last read before barrier (2) of the previous brick)
        if (tid < VMAX * 4) bbox[((it + 1) & 1) * (VMAX * 4) + tid] = 0x7fffffff;

        const char *planes = p.packed + (size_t)b * VMAX * p.plane_bytes;
        if (fits) {
            // ---- one bulk copy per patch row; view v is issued by warp v % 8 ----
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
            for (int v = 0; v < VMAX; ++v) {
                if ((v & (kSWarps - 1)) != warp || !((armed >> v) & 1)) continue;
                const int4 g = *reinterpret_cast<const int4 *>(&bb[4 * v]);
                const int pw = -g.z - g.x + 2, ph = -g.w - g.y + 2;
                if (lane == 0) mbar_expect(&bars[v], (unsigned)(pw * ph) * PS);
                __syncwarp();
                const char *src = planes + (size_t)v * p.plane_bytes + ((size_t)g.y * p.Wp + g.x) * PS;
                unsigned char *dst = patch + (size_t)v * q.view_bytes + 2 * PS;
                for (int r = lane; r < ph; r += 32)
                    bulk_g2s(dst + (size_t)r * pw * PS, src + (size_t)r * prow, (unsigned)pw * PS, &bars[v]);
            }
        }

        // ---- phase B: blend, fuse, store — one voxel per lane, NCH / T channel vectors per thread ----
        float *ob = p.out + (size_t)b * p.C * p.n_extent + (live ? n - p.n_origin : 0);
        const size_t cs = (size_t)p.n_extent;
        auto phase_b = [&](auto load) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                if ((k >> (LN - tl)) != sub) continue;                      // this thread's share of the vectors
                u64 s[VMAX][NP];
#pragma unroll
                for (int v = 0; v < VMAX; ++v) {
                    const uint4 t00 = load(v, b0[v] + 16 * k), t01 = load(v, b0[v] + 16 * k + PS);
                    const uint4 t10 = load(v, b1[v] + 16 * k), t11 = load(v, b1[v] + 16 * k + PS);
                    blend_texels<BF16>(s[v], t00, t01, t10, t11, w[v]);
                }
                Fuse2<METHOD, VMAX, true> fz[NP];
#pragma unroll
                for (int i = 0; i < NP; ++i) fz[i].absorb(&s[0][i], NP, VMAX, true);
                if (live) {
#pragma unroll
                    for (int i = 0; i < NP; ++i) {
                        const f2 r = fz[i].result(Vf);
                        const int c = k * CPV + 2 * i;
                        if (c < p.C) __stcs(ob + (size_t)c * cs, r.x);
                        if (c + 1 < p.C) __stcs(ob + (size_t)(c + 1) * cs, r.y);
                    }
                }
            }
        };
        if (fits) {
#pragma unroll
            for (int v = 0; v < VMAX; ++v)
                if ((armed >> v) & 1) mbar_wait(&bars[v], (phase >> v) & 1);
            phase ^= (unsigned)armed;
            phase_b([&](int, unsigned off) { return *reinterpret_cast<const uint4 *>(patch + off); });
        } else {
            // patch too large for shared memory: the same arithmetic straight from the planes
#pragma unroll
            for (int v = 0; v < VMAX; ++v) {
                const unsigned vb = (unsigned)v * (unsigned)p.plane_bytes;      // < 4 GiB per sample (checked by the host)
                if (cell[v] == 0xffffffffu) { b0[v] = vb; b1[v] = vb; }         // border texels of the plane: zeros
                else {
                    b0[v] = vb + ((cell[v] >> 16) * (unsigned)p.Wp + (cell[v] & 0xffffu)) * PS;
                    b1[v] = b0[v] + prow;
                }
            }
            phase_b([&](int, unsigned off) { return __ldg(reinterpret_cast<const uint4 *>(planes + off)); });
        }
        __syncthreads();                                            // (2) patches and boxes may be overwritten
    }
}

template <int VMAX, bool BF16, int NCH>
static int launch_staged_method(const StagedParams &q, int method, dim3 grid, size_t smem, cudaStream_t st)
{
#define MVHMR_SLAUNCH(M)                                                                                        \
    {                                                                                                           \
        auto kern = unproject_staged_kernel<VMAX, BF16, M, NCH>;                                                \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (e != cudaSuccess) return fail(MVHMR_ERR_CUDA, "unproject_staged_kernel: %s", cudaGetErrorString(e)); \
        kern<<<grid, kSThreads, smem, st>>>(q);                                                                 \
    }
    switch (method) {
    case MVHMR_SUM: MVHMR_SLAUNCH(MVHMR_SUM) break;
    case MVHMR_MEAN: MVHMR_SLAUNCH(MVHMR_MEAN) break;
    case MVHMR_MAX: MVHMR_SLAUNCH(MVHMR_MAX) break;
    default: MVHMR_SLAUNCH(MVHMR_SOFTMAX) break;
    }
#undef MVHMR_SLAUNCH
    return check_launch("unproject_staged_kernel");
}

int launch_unproject_staged(const UnprojParams &p, bool bf16, int method, void *stream)
{
    StagedParams q;
    q.u = p;
    const int V = p.V, nch = p.nchunks, ps = nch * 16 + 16;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int ctas = V <= 4 ? 3 : 2;                                       // resident CTAs per SM the patches are sized for
    if (const char *env = getenv("MVHMR_STAGED_CTAS")) { const int v = atoi(env); if (v >= 1 && v <= 4) ctas = v; }   // tuning knob
    const int budget = 232448 / ctas - 1024;                          // 227 KB per SM, 1 KB reserved per CTA
    q.off_patch = (64 + 2 * V * 16 + V * 4 + 127) & ~127;
    q.cap = (budget - q.off_patch) / (V * ps) - 2;
    if (q.cap > 4000) q.cap = 4000;
    q.view_bytes = (q.cap + 2) * ps;
    q.force_t = 0;
    if (const char *env = getenv("MVHMR_STAGED_T")) { const int v = atoi(env); if (v == 1 || v == 2 || v == 4 || v == 8) q.force_t = v; }
    const size_t smem = (size_t)q.off_patch + (size_t)V * q.view_bytes;
    const dim3 grid((unsigned)(sms * ctas));
    cudaStream_t st = (cudaStream_t)stream;
    if (V == 4) {
        if (bf16) return nch == 8 ? launch_staged_method<4, true, 8>(q, method, grid, smem, st) : launch_staged_method<4, true, 4>(q, method, grid, smem, st);
        return nch == 8 ? launch_staged_method<4, false, 8>(q, method, grid, smem, st) : launch_staged_method<4, false, 4>(q, method, grid, smem, st);
    }
    if (bf16) return nch == 8 ? launch_staged_method<8, true, 8>(q, method, grid, smem, st) : launch_staged_method<8, true, 4>(q, method, grid, smem, st);
    return nch == 8 ? launch_staged_method<8, false, 8>(q, method, grid, smem, st) : launch_staged_method<8, false, 4>(q, method, grid, smem, st);
}

}  // namespace mvhmr
