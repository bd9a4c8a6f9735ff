// 3-D soft-argmax over the aggregated volume:
//   p = softmax(vol[b,j,:]);  out[b,j,:] = sum_n p[n] * coord[b,n,:]
//
// HBM-bound single pass: every volume value is read exactly once with 16-byte
// streaming loads, every coordinate once per sample (not once per joint).  A CTA
// stages the coordinates of 2048 voxels of one sample in shared memory; its work
// items are (joint, 512-voxel quarter) pairs dealt round-robin to the 8 warps, so
// a warp reduces a whole item alone: per lane 16 values -> online-softmax record
// (max, sum e, sum e*x, sum e*y, sum e*z), ONE 5-step shuffle merge per item, no
// block-level synchronisation after the staging.  A tiny second kernel merges the
// per-quarter records — the same merge a slab-sharded multi-GPU run uses across
// ranks.
#include "mvhmr_common.cuh"

namespace mvhmr {

constexpr int kSaBlock = 256;
constexpr int kSaWarps = kSaBlock / 32;
constexpr int kSliceVox = 2048;                        // voxels staged per CTA
constexpr int kItemVox = 512;                          // voxels of one record (one warp, 16 per lane)
constexpr int kItemsPerSlice = kSliceVox / kItemVox;
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

// VEC: N % 4 == 0 and n0 % 4 == 0 -> 16-byte loads of vol
template <bool VEC>
__global__ void __launch_bounds__(kSaBlock, 3)
soft_argmax_partials_kernel(const float *__restrict__ vol, const float *__restrict__ coord,
                            float *__restrict__ partials, int J, long long N, long long n0, long long n1, int S)
{
    __shared__ __align__(16) float cs[kSliceVox * 3];          // xyz of the slice's voxels
    const int b = blockIdx.y, slice = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long base = n0 + (long long)slice * kSliceVox;
    const int nvox = (int)min((long long)kSliceVox, n1 - base);

    // stage the coordinates (contiguous nvox*3 floats)
    {
        const float *cp = coord + ((size_t)b * N + base) * 3;
        const int nfl = nvox * 3;
        if ((((uintptr_t)cp) & 15) == 0) {
            for (int i = threadIdx.x * 4; i < nfl; i += kSaBlock * 4) {
                if (i + 4 <= nfl) *reinterpret_cast<float4 *>(cs + i) = __ldg(reinterpret_cast<const float4 *>(cp + i));
                else for (int k = i; k < nfl; ++k) cs[k] = __ldg(cp + k);
            }
        } else {
            for (int i = threadIdx.x; i < nfl; i += kSaBlock) cs[i] = __ldg(cp + i);
        }
    }
    __syncthreads();

    const int nq = (nvox + kItemVox - 1) / kItemVox;             // quarters present in this slice
    const int nitems = J * nq;

    auto load_item = [&](int item, float *x) {
        const int j = (nq == kItemsPerSlice) ? item / kItemsPerSlice : item / nq, q = item - j * nq;
        const float *vp = vol + ((size_t)b * J + j) * N + base + q * kItemVox;
        const int qn = min(kItemVox, nvox - q * kItemVox);       // voxels in this quarter
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int v0 = (h * 32 + lane) * 4;                  // this lane's 4 consecutive voxels
            if (VEC && v0 + 4 <= qn) {
                const float4 a = __ldcs(reinterpret_cast<const float4 *>(vp + v0));
                x[4 * h] = a.x; x[4 * h + 1] = a.y; x[4 * h + 2] = a.z; x[4 * h + 3] = a.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) x[4 * h + i] = (v0 + i < qn) ? __ldcs(vp + v0 + i) : -INFINITY;
            }
        }
    };
    auto reduce_item = [&](int item, const float *x) {
        const int j = (nq == kItemsPerSlice) ? item / kItemsPerSlice : item / nq, q = item - j * nq;
        const int qn = min(kItemVox, nvox - q * kItemVox);
        const float *cq = cs + q * kItemVox * 3;
        Rec r;
        r.m = x[0];
#pragma unroll
        for (int i = 1; i < 16; ++i) r.m = fmaxf(r.m, x[i]);
        // the warp agrees on the item's max first, so the sums merge with plain adds
#pragma unroll
        for (int mask = 16; mask >= 1; mask >>= 1) r.m = fmaxf(r.m, __shfl_xor_sync(0xffffffffu, r.m, mask));
        const float ms = (r.m == -INFINITY) ? 0.0f : r.m;
        const float nm = -ms * kSaLog2e;
        r.S = r.X = r.Y = r.Z = 0.0f;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int v0 = (h * 32 + lane) * 4;
            float c[12];
            if (v0 + 4 <= qn) {
                const float4 a = *reinterpret_cast<const float4 *>(cq + v0 * 3);
                const float4 d = *reinterpret_cast<const float4 *>(cq + v0 * 3 + 4);
                const float4 e = *reinterpret_cast<const float4 *>(cq + v0 * 3 + 8);
                c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = d.x; c[5] = d.y;
                c[6] = d.z; c[7] = d.w; c[8] = e.x; c[9] = e.y; c[10] = e.z; c[11] = e.w;
            } else {
#pragma unroll
                for (int k = 0; k < 12; ++k) c[k] = (v0 * 3 + k < qn * 3) ? cq[v0 * 3 + k] : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float e = sa_ex2(fmaf(x[4 * h + i], kSaLog2e, nm));
                r.S += e;
                r.X = fmaf(e, c[3 * i], r.X);
                r.Y = fmaf(e, c[3 * i + 1], r.Y);
                r.Z = fmaf(e, c[3 * i + 2], r.Z);
            }
        }
#pragma unroll
        for (int mask = 16; mask >= 1; mask >>= 1) {
            r.S += __shfl_xor_sync(0xffffffffu, r.S, mask);
            r.X += __shfl_xor_sync(0xffffffffu, r.X, mask);
            r.Y += __shfl_xor_sync(0xffffffffu, r.Y, mask);
            r.Z += __shfl_xor_sync(0xffffffffu, r.Z, mask);
        }
        if (lane == 0) {
            float *o = partials + (((size_t)b * J + j) * S + (size_t)slice * kItemsPerSlice + q) * 5;
            o[0] = r.m; o[1] = r.S; o[2] = r.X; o[3] = r.Y; o[4] = r.Z;
        }
    };

    // two items in flight per warp: the next item's 16-byte loads are issued before the current
    // item's exp / sum chain
    float xa[16], xb[16];
    int item = warp;
    if (item < nitems) load_item(item, xa);
    while (item < nitems) {
        const int nxt = item + kSaWarps;
        if (nxt < nitems) load_item(nxt, xb);
        reduce_item(item, xa);
        if (nxt >= nitems) break;
        const int nx2 = nxt + kSaWarps;
        if (nx2 < nitems) load_item(nx2, xa);
        reduce_item(nxt, xb);
        item = nx2;
    }
    // quarters that do not exist in a ragged last slice still own a record slot: mark them empty
    for (int item = threadIdx.x; item < J * (kItemsPerSlice - nq); item += kSaBlock) {
        const int j = item / (kItemsPerSlice - nq), q = nq + item % (kItemsPerSlice - nq);
        if ((size_t)slice * kItemsPerSlice + q < (size_t)S) {
            float *o = partials + (((size_t)b * J + j) * S + (size_t)slice * kItemsPerSlice + q) * 5;
            o[0] = -INFINITY; o[1] = 0.0f; o[2] = 0.0f; o[3] = 0.0f; o[4] = 0.0f;
        }
    }
}

// one 128-thread CTA per (b,j): merge S records, divide
__global__ void __launch_bounds__(128)
soft_argmax_finalize_kernel(const float *__restrict__ partials, float *__restrict__ out, int BJ, int S)
{
    __shared__ float sm[4][5];
    const int bj = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Rec r;
    r.m = -INFINITY; r.S = r.X = r.Y = r.Z = 0.0f;
    for (int s = threadIdx.x; s < S; s += 128) {
        const float *q = partials + ((size_t)bj * S + s) * 5;
        Rec c;
        c.m = q[0]; c.S = q[1]; c.X = q[2]; c.Y = q[3]; c.Z = q[4];
        merge(r, c);
    }
#pragma unroll
    for (int mask = 16; mask >= 1; mask >>= 1) merge(r, shfl_xor(r, mask));
    if (lane == 0) { sm[warp][0] = r.m; sm[warp][1] = r.S; sm[warp][2] = r.X; sm[warp][3] = r.Y; sm[warp][4] = r.Z; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 4; ++w) {
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

extern "C" int mvhmr_soft_argmax3d_partials(const float *vol, const float *coord, float *partials,
                                            int B, int J, long long N, long long n0, long long n1, void *stream)
{
    if (B < 0 || J < 0 || N < 1 || n0 < 0 || n1 > N || n0 >= n1)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "soft_argmax3d: bad shape B=%d J=%d N=%lld window [%lld,%lld)", B, J, N, n0, n1);
    if (B == 0 || J == 0) return MVHMR_OK;
    if (B > 65535) return fail(MVHMR_ERR_INVALID_ARGUMENT, "soft_argmax3d: B=%d exceeds 65535", B);
    if (!vol || !coord || !partials) return fail(MVHMR_ERR_INVALID_ARGUMENT, "soft_argmax3d: null pointer");
    const int S = mvhmr_soft_argmax3d_num_slices(n1 - n0);
    dim3 grid((unsigned)((n1 - n0 + kSliceVox - 1) / kSliceVox), B);
    const bool vec = (N % 4 == 0) && (n0 % 4 == 0) && (((uintptr_t)vol & 15) == 0);
    if (vec)
        soft_argmax_partials_kernel<true><<<grid, kSaBlock, 0, (cudaStream_t)stream>>>(vol, coord, partials, J, N, n0, n1, S);
    else
        soft_argmax_partials_kernel<false><<<grid, kSaBlock, 0, (cudaStream_t)stream>>>(vol, coord, partials, J, N, n0, n1, S);
    return check_launch("soft_argmax_partials_kernel");
}

extern "C" int mvhmr_soft_argmax3d_finalize(const float *partials, float *out, int B, int J, int S, void *stream)
{
    if (B < 0 || J < 0 || S < 1) return fail(MVHMR_ERR_INVALID_ARGUMENT, "soft_argmax3d_finalize: bad shape B=%d J=%d S=%d", B, J, S);
    if (B == 0 || J == 0) return MVHMR_OK;
    if (!partials || !out) return fail(MVHMR_ERR_INVALID_ARGUMENT, "soft_argmax3d_finalize: null pointer");
    const int BJ = B * J;
    soft_argmax_finalize_kernel<<<BJ, 128, 0, (cudaStream_t)stream>>>(partials, out, BJ, S);
    return check_launch("soft_argmax_finalize_kernel");
}

extern "C" int mvhmr_soft_argmax3d(const float *vol, const float *coord, float *out,
                                   int B, int J, long long N, void *ws, size_t ws_bytes, void *stream)
{
    const size_t need = mvhmr_soft_argmax3d_workspace_bytes(B, J, N);
    if (B > 0 && J > 0 && (!ws || ws_bytes < need))
        return fail(MVHMR_ERR_WORKSPACE, "soft_argmax3d: workspace of %zu bytes required, got %zu", need, ws_bytes);
    int rc = mvhmr_soft_argmax3d_partials(vol, coord, (float *)ws, B, J, N, 0, N, stream);
    if (rc != MVHMR_OK) return rc;
    return mvhmr_soft_argmax3d_finalize((const float *)ws, out, B, J, mvhmr_soft_argmax3d_num_slices(N), stream);
}
