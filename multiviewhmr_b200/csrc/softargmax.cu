// 3-D soft-argmax over the aggregated volume:
//   p = softmax(vol[b,j,:]);  out[b,j,:] = sum_n p[n] * coord[b,n,:]
//
// HBM-bound single pass: every volume value is read exactly once with 16-byte
// streaming loads, every coordinate once per sample (not once per joint).  A warp owns
// 512 voxels of one sample: it keeps their coordinates in registers (16 voxels per lane)
// and walks a chunk of <= 8 joints, three joints' loads in flight; per joint the warp agrees on the
// max first, so the partial sums (sum e, sum e*x, sum e*y, sum e*z) merge with plain
// shuffle-adds, and writes one record per (b, joint, 512-voxel slice).  No shared memory,
// no block-level synchronisation.  A tiny second kernel merges the per-slice records —
// the same merge a slab-sharded multi-GPU run uses across ranks.
#include "mvhmr_common.cuh"

namespace mvhmr {

constexpr int kSaBlock = 128;
constexpr int kSaWarps = kSaBlock / 32;
constexpr int kItemVox = 512;                          // voxels of one record (one warp, 16 per lane)
constexpr int kSaJointChunk = 8;                       // at most this many joints per warp (grid.z splits J)
constexpr float kSaLog2e = 1.4426950408889634f;

__device__ __forceinline__ float sa_ex2(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct Rec { float m, S, X, Y, Z; };

// merge b into a; safe for empty records (m = -inf, sums = 0)
__device__ __forceinline__ void merge(Rec &a, const Rec &b)
{
    const float mn = fmaxf(a.m, b.m);
    const float sa = (a.m == mn) ? 1.0f : sa_ex2((a.m - mn) * kSaLog2e);
    const float sb = (b.m == mn) ? 1.0f : sa_ex2((b.m - mn) * kSaLog2e);
    a.S = a.S * sa + b.S * sb;
    a.X = a.X * sa + b.X * sb;
    a.Y = a.Y * sa + b.Y * sb;
    a.Z = a.Z * sa + b.Z * sb;
    a.m = mn;
}

__device__ __forceinline__ Rec shfl_xor(const Rec &r, int mask)
{
    Rec o;
    o.m = __shfl_xor_sync(0xffffffffu, r.m, mask);
    o.S = __shfl_xor_sync(0xffffffffu, r.S, mask);
    o.X = __shfl_xor_sync(0xffffffffu, r.X, mask);
    o.Y = __shfl_xor_sync(0xffffffffu, r.Y, mask);
    o.Z = __shfl_xor_sync(0xffffffffu, r.Z, mask);
    return o;
}

// Cuboid grid of models/aggregation.py:135-187 for the GRID variant (coordinates generated, never read)
struct SaGrid {
    const float *centers, *rot;
    float pos[3], step[3];
    int gy, gz;
};

// VEC: N % 4 == 0 and n0 % 4 == 0 -> 16-byte loads of vol
// GRID: the voxel coordinates come from the grid descriptor with coord_volume_kernel's arithmetic
//       (bit-identical to reading a built coord volume) instead of from `coord`
template <bool VEC, bool GRID>
__global__ void __launch_bounds__(kSaBlock)
soft_argmax_partials_kernel(const float *__restrict__ vol, const float *__restrict__ coord,
                            float *__restrict__ partials, int J, int jchunk, long long N, long long n0, long long n1, int S,
                            long long bstride, const SaGrid g)
{
    const int b = blockIdx.y;
    const int j0 = blockIdx.z * jchunk, j1 = min(J, j0 + jchunk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slice = blockIdx.x * kSaWarps + warp;             // 512-voxel slice of this warp
    if (slice >= S) return;
    const long long base = n0 + (long long)slice * kItemVox;
    const int qn = (int)min((long long)kItemVox, n1 - base);    // voxels in this slice

    // coordinates of this lane's 16 voxels (4 runs of 4 consecutive voxels), kept in registers
    float cx[16], cy[16], cz[16];
    if (GRID) {
        const float c0 = __ldg(g.centers + 3 * b), c1 = __ldg(g.centers + 3 * b + 1), c2 = __ldg(g.centers + 3 * b + 2);
        float R[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(g.rot + 9 * b + i);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const long long n = base + (h * 32 + lane) * 4;           // first voxel of this run of four
            int ix, iy, iz;
            if (N <= 0x7fffffffLL) {                                   // 32-bit index arithmetic
                const unsigned t = (unsigned)n / (unsigned)g.gz;
                iz = (int)((unsigned)n - t * (unsigned)g.gz);
                ix = (int)(t / (unsigned)g.gy);
                iy = (int)(t - (unsigned)ix * (unsigned)g.gy);
            } else {
                iz = (int)(n % g.gz);
                const long long t = n / g.gz;
                iy = (int)(t % g.gy); ix = (int)(t / g.gy);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float d0 = __fsub_rn(__fadd_rn(g.pos[0], __fmul_rn(g.step[0], (float)ix)), c0);
                const float d1 = __fsub_rn(__fadd_rn(g.pos[1], __fmul_rn(g.step[1], (float)iy)), c1);
                const float d2 = __fsub_rn(__fadd_rn(g.pos[2], __fmul_rn(g.step[2], (float)iz)), c2);
                const bool in = (h * 32 + lane) * 4 + i < qn;
                cx[4 * h + i] = in ? __fadd_rn(rot_row(R[0], R[1], R[2], d0, d1, d2), c0) : 0.0f;
                cy[4 * h + i] = in ? __fadd_rn(rot_row(R[3], R[4], R[5], d0, d1, d2), c1) : 0.0f;
                cz[4 * h + i] = in ? __fadd_rn(rot_row(R[6], R[7], R[8], d0, d1, d2), c2) : 0.0f;
                if (++iz == g.gz) { iz = 0; if (++iy == g.gy) { iy = 0; ++ix; } }
            }
        }
    } else {
        const float *cp = coord + ((size_t)b * N + base) * 3;
        const bool al = (((uintptr_t)cp) & 15) == 0;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int v0 = (h * 32 + lane) * 4;
            if (al && v0 + 4 <= qn) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(cp + v0 * 3));
                const float4 d = __ldg(reinterpret_cast<const float4 *>(cp + v0 * 3) + 1);
                const float4 e = __ldg(reinterpret_cast<const float4 *>(cp + v0 * 3) + 2);
                cx[4 * h] = a.x; cy[4 * h] = a.y; cz[4 * h] = a.z;
                cx[4 * h + 1] = a.w; cy[4 * h + 1] = d.x; cz[4 * h + 1] = d.y;
                cx[4 * h + 2] = d.z; cy[4 * h + 2] = d.w; cz[4 * h + 2] = e.x;
                cx[4 * h + 3] = e.y; cy[4 * h + 3] = e.z; cz[4 * h + 3] = e.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool in = v0 + i < qn;
                    cx[4 * h + i] = in ? __ldg(cp + (v0 + i) * 3) : 0.0f;
                    cy[4 * h + i] = in ? __ldg(cp + (v0 + i) * 3 + 1) : 0.0f;
                    cz[4 * h + i] = in ? __ldg(cp + (v0 + i) * 3 + 2) : 0.0f;
                }
            }
        }
    }

    auto load_joint = [&](int j, float *x) {
        const float *vp = vol + (size_t)b * bstride + (size_t)j * N + base;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int v0 = (h * 32 + lane) * 4;
            if (VEC && v0 + 4 <= qn) {
                const float4 a = __ldcs(reinterpret_cast<const float4 *>(vp + v0));
                x[4 * h] = a.x; x[4 * h + 1] = a.y; x[4 * h + 2] = a.z; x[4 * h + 3] = a.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) x[4 * h + i] = (v0 + i < qn) ? __ldcs(vp + v0 + i) : -INFINITY;
            }
        }
    };
    auto reduce_joint = [&](int j, const float *x) {
        Rec r;
        r.m = x[0];
#pragma unroll
        for (int i = 1; i < 16; ++i) r.m = fmaxf(r.m, x[i]);
        // the warp agrees on the slice's max first, so the sums merge with plain adds
#pragma unroll
        for (int mask = 16; mask >= 1; mask >>= 1) r.m = fmaxf(r.m, __shfl_xor_sync(0xffffffffu, r.m, mask));
        const float ms = (r.m == -INFINITY) ? 0.0f : r.m;
        const float nm = -ms * kSaLog2e;
        r.S = r.X = r.Y = r.Z = 0.0f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float e = sa_ex2(fmaf(x[i], kSaLog2e, nm));
            r.S += e;
            r.X = fmaf(e, cx[i], r.X);
            r.Y = fmaf(e, cy[i], r.Y);
            r.Z = fmaf(e, cz[i], r.Z);
        }
#pragma unroll
        for (int mask = 16; mask >= 1; mask >>= 1) {
            r.S += __shfl_xor_sync(0xffffffffu, r.S, mask);
            r.X += __shfl_xor_sync(0xffffffffu, r.X, mask);
            r.Y += __shfl_xor_sync(0xffffffffu, r.Y, mask);
            r.Z += __shfl_xor_sync(0xffffffffu, r.Z, mask);
        }
        if (lane == 0) {
            float *o = partials + (((size_t)b * J + j) * S + slice) * 5;
            o[0] = r.m; o[1] = r.S; o[2] = r.X; o[3] = r.Y; o[4] = r.Z;
        }
    };

    // three joints' loads in flight per warp: joint j+2 is requested before joint j's
    // exp / sum chain runs
    float xa[16], xb[16], xc[16];
    load_joint(j0, xa);
    if (j0 + 1 < j1) load_joint(j0 + 1, xb);
    for (int j = j0; j < j1; j += 3) {
        if (j + 2 < j1) load_joint(j + 2, xc);
        reduce_joint(j, xa);
        if (j + 1 >= j1) break;
        if (j + 3 < j1) load_joint(j + 3, xa);
        reduce_joint(j + 1, xb);
        if (j + 2 >= j1) break;
        if (j + 4 < j1) load_joint(j + 4, xb);
        reduce_joint(j + 2, xc);
    }
}

// one CTA per (b,j): merge S records, divide.  Four records' loads are in flight per thread (the merge
// is a serial exp chain; with one record at a time the kernel is pure load latency: 11.7 us for the
// 2368 records per joint of the fused kernel, 4.5 us for 512).
constexpr int kFinBlock = 256;
__global__ void __launch_bounds__(kFinBlock)
soft_argmax_finalize_kernel(const float *__restrict__ partials, float *__restrict__ out, int BJ, int S)
{
    __shared__ float sm[kFinBlock / 32][5];
    const int bj = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Rec r;
    r.m = -INFINITY; r.S = r.X = r.Y = r.Z = 0.0f;
    const float *base = partials + (size_t)bj * S * 5;
    for (int s0 = threadIdx.x; s0 < S; s0 += 4 * kFinBlock) {
        Rec c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int s = s0 + u * kFinBlock;
            const float *q = base + (size_t)(s < S ? s : s0) * 5;
            c[u].m = q[0]; c[u].S = (s < S) ? q[1] : 0.0f; c[u].X = q[2]; c[u].Y = q[3]; c[u].Z = q[4];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c[u].S != 0.0f) merge(r, c[u]);         // S == 0: empty slot (all-zero records of the fused kernel's idle warps)
    }
#pragma unroll
    for (int mask = 16; mask >= 1; mask >>= 1) merge(r, shfl_xor(r, mask));
    if (lane == 0) { sm[warp][0] = r.m; sm[warp][1] = r.S; sm[warp][2] = r.X; sm[warp][3] = r.Y; sm[warp][4] = r.Z; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < kFinBlock / 32; ++w) {
            Rec c;
            c.m = sm[w][0]; c.S = sm[w][1]; c.X = sm[w][2]; c.Y = sm[w][3]; c.Z = sm[w][4];
            merge(r, c);
        }
        out[3 * bj + 0] = __fdiv_rn(r.X, r.S);
        out[3 * bj + 1] = __fdiv_rn(r.Y, r.S);
        out[3 * bj + 2] = __fdiv_rn(r.Z, r.S);
    }
}

}  // namespace mvhmr

using namespace mvhmr;

extern "C" int mvhmr_soft_argmax3d_num_slices(long long N)
{
    if (N <= 0) return 0;
    return (int)((N + kItemVox - 1) / kItemVox);
}

extern "C" size_t mvhmr_soft_argmax3d_workspace_bytes(int B, int J, long long N)
{
    if (B < 0 || J < 0 || N < 0) return 0;
    return (size_t)B * J * mvhmr_soft_argmax3d_num_slices(N) * 5 * sizeof(float);
}

static int partials_impl(const float *vol, const float *coord, const SaGrid *grid, float *partials,
                         int B, int J, long long N, long long n0, long long n1, long long bstride, void *stream)
{
    if (B < 0 || J < 0 || N < 1 || n0 < 0 || n1 > N || n0 >= n1)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "soft_argmax3d: bad shape B=%d J=%d N=%lld window [%lld,%lld)", B, J, N, n0, n1);
    if (B == 0 || J == 0) return MVHMR_OK;
    if (B > 65535) return fail(MVHMR_ERR_INVALID_ARGUMENT, "soft_argmax3d: B=%d exceeds 65535", B);
    if (!vol || (!coord && !grid) || !partials) return fail(MVHMR_ERR_INVALID_ARGUMENT, "soft_argmax3d: null pointer");
    const int S = mvhmr_soft_argmax3d_num_slices(n1 - n0);
    const int nz = (J + kSaJointChunk - 1) / kSaJointChunk, jchunk = (J + nz - 1) / nz;
    dim3 grid_dim((unsigned)((S + kSaWarps - 1) / kSaWarps), B, nz);
    if (bstride < (long long)J * N)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "soft_argmax3d: sample stride %lld smaller than J*N", bstride);
    const bool vec = (N % 4 == 0) && (n0 % 4 == 0) && (bstride % 4 == 0) && (((uintptr_t)vol & 15) == 0);
    const SaGrid none = {};
    cudaStream_t st = (cudaStream_t)stream;
    if (grid) {
        if (vec) soft_argmax_partials_kernel<true, true><<<grid_dim, kSaBlock, 0, st>>>(vol, nullptr, partials, J, jchunk, N, n0, n1, S, bstride, *grid);
        else soft_argmax_partials_kernel<false, true><<<grid_dim, kSaBlock, 0, st>>>(vol, nullptr, partials, J, jchunk, N, n0, n1, S, bstride, *grid);
    } else {
        if (vec) soft_argmax_partials_kernel<true, false><<<grid_dim, kSaBlock, 0, st>>>(vol, coord, partials, J, jchunk, N, n0, n1, S, bstride, none);
        else soft_argmax_partials_kernel<false, false><<<grid_dim, kSaBlock, 0, st>>>(vol, coord, partials, J, jchunk, N, n0, n1, S, bstride, none);
    }
    return check_launch("soft_argmax_partials_kernel");
}

extern "C" int mvhmr_soft_argmax3d_partials(const float *vol, const float *coord, float *partials,
                                            int B, int J, long long N, long long n0, long long n1, void *stream)
{
    return partials_impl(vol, coord, nullptr, partials, B, J, N, n0, n1, (long long)J * N, stream);
}

extern "C" int mvhmr_soft_argmax3d_finalize(const float *partials, float *out, int B, int J, int S, void *stream)
{
    if (B < 0 || J < 0 || S < 1) return fail(MVHMR_ERR_INVALID_ARGUMENT, "soft_argmax3d_finalize: bad shape B=%d J=%d S=%d", B, J, S);
    if (B == 0 || J == 0) return MVHMR_OK;
    if (!partials || !out) return fail(MVHMR_ERR_INVALID_ARGUMENT, "soft_argmax3d_finalize: null pointer");
    const int BJ = B * J;
    soft_argmax_finalize_kernel<<<BJ, kFinBlock, 0, (cudaStream_t)stream>>>(partials, out, BJ, S);
    return check_launch("soft_argmax_finalize_kernel");
}

extern "C" int mvhmr_soft_argmax3d(const float *vol, const float *coord, float *out,
                                   int B, int J, long long N, void *ws, size_t ws_bytes, void *stream)
{
    return mvhmr_soft_argmax3d_strided(vol, coord, out, B, J, N, (long long)J * N, ws, ws_bytes, stream);
}

extern "C" int mvhmr_soft_argmax3d_strided(const float *vol, const float *coord, float *out,
                                           int B, int J, long long N, long long sample_stride,
                                           void *ws, size_t ws_bytes, void *stream)
{
    const size_t need = mvhmr_soft_argmax3d_workspace_bytes(B, J, N);
    if (B > 0 && J > 0 && (!ws || ws_bytes < need))
        return fail(MVHMR_ERR_WORKSPACE, "soft_argmax3d: workspace of %zu bytes required, got %zu", need, ws_bytes);
    int rc = partials_impl(vol, coord, nullptr, (float *)ws, B, J, N, 0, N, sample_stride, stream);
    if (rc != MVHMR_OK) return rc;
    return mvhmr_soft_argmax3d_finalize((const float *)ws, out, B, J, mvhmr_soft_argmax3d_num_slices(N), stream);
}

extern "C" int mvhmr_soft_argmax3d_grid(const float *vol, const mvhmr_grid_t *grid, float *out,
                                        int B, int J, int gx, int gy, int gz, long long sample_stride,
                                        void *ws, size_t ws_bytes, void *stream)
{
    if (!grid || !grid->centers || !grid->rot)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "soft_argmax3d_grid: null grid descriptor / centers / rot");
    if (gx < 1 || gy < 1 || gz < 1)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "soft_argmax3d_grid: bad grid (%d,%d,%d)", gx, gy, gz);
    const long long N = (long long)gx * gy * gz;
    const size_t need = mvhmr_soft_argmax3d_workspace_bytes(B, J, N);
    if (B > 0 && J > 0 && (!ws || ws_bytes < need))
        return fail(MVHMR_ERR_WORKSPACE, "soft_argmax3d: workspace of %zu bytes required, got %zu", need, ws_bytes);
    SaGrid g;
    g.centers = grid->centers; g.rot = grid->rot;
    for (int k = 0; k < 3; ++k) { g.pos[k] = grid->pos[k]; g.step[k] = grid->step[k]; }
    g.gy = gy; g.gz = gz;
    int rc = partials_impl(vol, nullptr, &g, (float *)ws, B, J, N, 0, N, sample_stride, stream);
    if (rc != MVHMR_OK) return rc;
    return mvhmr_soft_argmax3d_finalize((const float *)ws, out, B, J, mvhmr_soft_argmax3d_num_slices(N), stream);
}
