// 3-D soft-argmax over the aggregated volume:
//   p = softmax(vol[b,j,:]);  out[b,j,:] = sum_n p[n] * coord[b,n,:]
//
// HBM-bound single pass: every volume value is read exactly once with 16-byte
// loads, every coordinate once per sample (not once per joint).  A CTA owns a
// slice of kSliceVox voxels of one sample, keeps its coordinates in registers
// and walks all J joints; per joint each thread forms an online-softmax record
// (max, sum e, sum e*x, sum e*y, sum e*z), records are merged inside the warp
// with shuffles and across warps through shared memory with a single
// __syncthreads per 64 joints.  A tiny second kernel merges the per-slice
// records — the same merge a slab-sharded multi-GPU run uses across ranks.
#include "mvhmr_common.cuh"

namespace mvhmr {

constexpr int kSaBlock = 256;
constexpr int kSaWarps = kSaBlock / 32;
constexpr int kSaPerThread = 8;                        // two float4 per joint
constexpr int kSliceVox = kSaBlock * kSaPerThread;     // 2048 voxels per CTA
constexpr int kSaJGroup = 64;
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

// VEC: N % 4 == 0 and n0 % 4 == 0 -> 16-byte loads of vol and coord
template <bool VEC>
__global__ void __launch_bounds__(kSaBlock)
soft_argmax_partials_kernel(const float *__restrict__ vol, const float *__restrict__ coord,
                            float *__restrict__ partials, int J, long long N, long long n0, long long n1, int S)
{
    __shared__ float sm[kSaJGroup][kSaWarps][5];
    const int b = blockIdx.y, slice = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long base = n0 + (long long)slice * kSliceVox;

    // this thread's voxels: two runs of 4 consecutive voxels
    long long vn[2];
    float cx[kSaPerThread], cy[kSaPerThread], cz[kSaPerThread];
    bool ok[kSaPerThread];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        vn[h] = base + ((long long)h * kSaBlock + threadIdx.x) * 4;
        const float *cp = coord + ((size_t)b * N + vn[h]) * 3;
        if (VEC && vn[h] + 4 <= n1) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(cp));
            const float4 c = __ldg(reinterpret_cast<const float4 *>(cp) + 1);
            const float4 d = __ldg(reinterpret_cast<const float4 *>(cp) + 2);
            cx[4 * h + 0] = a.x; cy[4 * h + 0] = a.y; cz[4 * h + 0] = a.z;
            cx[4 * h + 1] = a.w; cy[4 * h + 1] = c.x; cz[4 * h + 1] = c.y;
            cx[4 * h + 2] = c.z; cy[4 * h + 2] = c.w; cz[4 * h + 2] = d.x;
            cx[4 * h + 3] = d.y; cy[4 * h + 3] = d.z; cz[4 * h + 3] = d.w;
#pragma unroll
            for (int i = 0; i < 4; ++i) ok[4 * h + i] = true;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool in = vn[h] + i < n1;
                ok[4 * h + i] = in;
                cx[4 * h + i] = in ? __ldg(cp + 3 * i) : 0.0f;
                cy[4 * h + i] = in ? __ldg(cp + 3 * i + 1) : 0.0f;
                cz[4 * h + i] = in ? __ldg(cp + 3 * i + 2) : 0.0f;
            }
        }
    }

    for (int j0 = 0; j0 < J; j0 += kSaJGroup) {
        const int jn = min(kSaJGroup, J - j0);
        for (int jj = 0; jj < jn; ++jj) {
            const float *vp = vol + ((size_t)b * J + (j0 + jj)) * N;
            float x[kSaPerThread];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (VEC && vn[h] + 4 <= n1) {
                    const float4 a = __ldcs(reinterpret_cast<const float4 *>(vp + vn[h]));
                    x[4 * h] = a.x; x[4 * h + 1] = a.y; x[4 * h + 2] = a.z; x[4 * h + 3] = a.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[4 * h + i] = ok[4 * h + i] ? __ldcs(vp + vn[h] + i) : -INFINITY;
                }
            }
            Rec r;
            r.m = x[0];
#pragma unroll
            for (int i = 1; i < kSaPerThread; ++i) r.m = fmaxf(r.m, x[i]);
            const float ms = (r.m == -INFINITY) ? 0.0f : r.m;
            r.S = r.X = r.Y = r.Z = 0.0f;
#pragma unroll
            for (int i = 0; i < kSaPerThread; ++i) {
                const float e = sa_ex2((x[i] - ms) * kSaLog2e);
                r.S += e;
                r.X = fmaf(e, cx[i], r.X);
                r.Y = fmaf(e, cy[i], r.Y);
                r.Z = fmaf(e, cz[i], r.Z);
            }
#pragma unroll
            for (int mask = 16; mask >= 1; mask >>= 1) merge(r, shfl_xor(r, mask));
            if (lane == 0) {
                sm[jj][warp][0] = r.m; sm[jj][warp][1] = r.S; sm[jj][warp][2] = r.X;
                sm[jj][warp][3] = r.Y; sm[jj][warp][4] = r.Z;
            }
        }
        __syncthreads();
        if (threadIdx.x < jn) {
            Rec a;
            a.m = sm[threadIdx.x][0][0]; a.S = sm[threadIdx.x][0][1]; a.X = sm[threadIdx.x][0][2];
            a.Y = sm[threadIdx.x][0][3]; a.Z = sm[threadIdx.x][0][4];
#pragma unroll
            for (int w = 1; w < kSaWarps; ++w) {
                Rec c;
                c.m = sm[threadIdx.x][w][0]; c.S = sm[threadIdx.x][w][1]; c.X = sm[threadIdx.x][w][2];
                c.Y = sm[threadIdx.x][w][3]; c.Z = sm[threadIdx.x][w][4];
                merge(a, c);
            }
            float *o = partials + (((size_t)b * J + (j0 + threadIdx.x)) * S + slice) * 5;
            o[0] = a.m; o[1] = a.S; o[2] = a.X; o[3] = a.Y; o[4] = a.Z;
        }
        __syncthreads();
    }
}

// one warp per (b,j): merge S records, divide
__global__ void __launch_bounds__(128)
soft_argmax_finalize_kernel(const float *__restrict__ partials, float *__restrict__ out, int BJ, int S)
{
    const int bj = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (bj >= BJ) return;
    Rec r;
    r.m = -INFINITY; r.S = r.X = r.Y = r.Z = 0.0f;
    for (int s = lane; s < S; s += 32) {
        const float *q = partials + ((size_t)bj * S + s) * 5;
        Rec c;
        c.m = q[0]; c.S = q[1]; c.X = q[2]; c.Y = q[3]; c.Z = q[4];
        merge(r, c);
    }
#pragma unroll
    for (int mask = 16; mask >= 1; mask >>= 1) merge(r, shfl_xor(r, mask));
    if (lane == 0) {
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
    return (int)((N + kSliceVox - 1) / kSliceVox);
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
    dim3 grid(S, B);
    const bool vec = (N % 4 == 0) && (n0 % 4 == 0) && (((uintptr_t)vol & 15) == 0) && (((uintptr_t)coord & 15) == 0);
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
    soft_argmax_finalize_kernel<<<(BJ + 3) / 4, 128, 0, (cudaStream_t)stream>>>(partials, out, BJ, S);
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
