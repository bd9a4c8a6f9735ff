// Backward of unproject + aggregate w.r.t. the feature maps.
//
// out[b,c,n] = fuse_v( s_v ),  s_v = sum_k w_vk * feat[b,v,c,corner_k]   (0 if depth <= 0)
//   d out / d s_v :  sum 1 | mean 1/V | max [v == argmax] | softmax p_v * (1 + s_v - out)
// (the reference's autograd through (x * softmax(x)).sum(0), models/aggregation.py:77-83).
// Views with depth <= 0 were overwritten with zeros in place (:62) and pass no gradient to
// the features, but they still take part in the fusion.  One thread per voxel recomputes the
// cells with the forward's exact arithmetic, re-samples the NCHW maps and scatters with
// red.global.add.f32 into an fp32 gradient buffer (zeroed by the caller).
#include "mvhmr_common.cuh"

namespace mvhmr {

constexpr int kBwdMaxViews = 64;

struct BwdCell {
    int x0, y0;            // nw corner (may be outside the map)
    float w[4];            // nw ne sw se
    int valid;             // depth > 0
};

__device__ __forceinline__ BwdCell bwd_cell(const float *P, float X, float Y, float Z, int H, int W)
{
    const float xw = proj_row(X, Y, Z, P[0], P[1], P[2], P[3]);
    const float yw = proj_row(X, Y, Z, P[4], P[5], P[6], P[7]);
    const float ww = proj_row(X, Y, Z, P[8], P[9], P[10], P[11]);
    BwdCell c;
    c.valid = !(ww <= 0.0f);
    const float wd = (ww == 0.0f) ? 1.0f : ww;
    const float x = __fdiv_rn(xw, wd), y = __fdiv_rn(yw, wd);
    const float gx = __fmul_rn(2.0f, __fsub_rn(__fdiv_rn(x, (float)H), 0.5f));
    const float gy = __fmul_rn(2.0f, __fsub_rn(__fdiv_rn(y, (float)W), 0.5f));
    const float ix = __fmul_rn(__fadd_rn(gx, 1.0f), (float)(W - 1) / 2.0f);
    const float iy = __fmul_rn(__fadd_rn(gy, 1.0f), (float)(H - 1) / 2.0f);
    const float xf = floorf(ix), yf = floorf(iy);
    const float fw = __fsub_rn(ix, xf), fe = __fsub_rn(1.0f, fw), fn = __fsub_rn(iy, yf), fs = __fsub_rn(1.0f, fn);
    c.w[0] = __fmul_rn(fs, fe); c.w[1] = __fmul_rn(fs, fw); c.w[2] = __fmul_rn(fn, fe); c.w[3] = __fmul_rn(fn, fw);
    c.x0 = (int)fminf(fmaxf(xf, -2.0f), (float)W);      // clamped: anything further out has no in-map corner
    c.y0 = (int)fminf(fmaxf(yf, -2.0f), (float)H);
    return c;
}

template <bool BF16>
__device__ __forceinline__ float load_feat(const void *f, size_t i)
{
    if (BF16) return __uint_as_float((unsigned)__ldg(static_cast<const unsigned short *>(f) + i) << 16);
    return __ldg(static_cast<const float *>(f) + i);
}

template <bool BF16>
__global__ void __launch_bounds__(128)
unproject_backward_kernel(const float *__restrict__ gout, const void *__restrict__ feats, const float *__restrict__ proj,
                          const float *__restrict__ coord, float *__restrict__ gfeat,
                          int V, int C, int H, int W, long long N, int method)
{
    const int b = blockIdx.y;
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float *xyz = coord + ((size_t)b * N + n) * 3;
    const float X = __ldg(xyz), Y = __ldg(xyz + 1), Z = __ldg(xyz + 2);
    BwdCell cell[kBwdMaxViews];
    for (int v = 0; v < V; ++v) {
        float P[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) P[i] = __ldg(proj + ((size_t)b * V + v) * 12 + i);
        cell[v] = bwd_cell(P, X, Y, Z, H, W);
    }
    const size_t hw = (size_t)H * W;
    for (int c = 0; c < C; ++c) {
        const float g = __ldg(gout + ((size_t)b * C + c) * N + n);
        float s[kBwdMaxViews];
        for (int v = 0; v < V; ++v) {
            const size_t plane = (((size_t)b * V + v) * C + c) * hw;
            float acc = 0.0f;
            if (cell[v].valid) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int xx = cell[v].x0 + (k & 1), yy = cell[v].y0 + (k >> 1);
                    const float t = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? load_feat<BF16>(feats, plane + (size_t)yy * W + xx) : 0.0f;
                    acc = k == 0 ? __fmul_rn(t, cell[v].w[0]) : __fmaf_rn(t, cell[v].w[k], acc);
                }
            }
            s[v] = acc;
        }
        // d out / d s_v
        float m = s[0];
        int arg = 0;
        for (int v = 1; v < V; ++v) if (s[v] > m || (s[v] != s[v] && m == m)) { m = s[v]; arg = v; }
        float S = 0.0f, A = 0.0f;
        if (method == MVHMR_SOFTMAX) {
            for (int v = 0; v < V; ++v) { const float e = expf(s[v] - m); S += e; A = fmaf(s[v], e, A); }
        }
        const float o = (method == MVHMR_SOFTMAX) ? A / S : 0.0f;
        for (int v = 0; v < V; ++v) {
            if (!cell[v].valid) continue;
            float d;
            if (method == MVHMR_SUM) d = 1.0f;
            else if (method == MVHMR_MEAN) d = 1.0f / (float)V;
            else if (method == MVHMR_MAX) d = (v == arg) ? 1.0f : 0.0f;
            else d = expf(s[v] - m) / S * (1.0f + s[v] - o);
            const float gv = g * d;
            if (gv == 0.0f) continue;
            const size_t plane = (((size_t)b * V + v) * C + c) * hw;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xx = cell[v].x0 + (k & 1), yy = cell[v].y0 + (k >> 1);
                if (xx >= 0 && xx < W && yy >= 0 && yy < H) atomicAdd(gfeat + plane + (size_t)yy * W + xx, gv * cell[v].w[k]);
            }
        }
    }
}

}  // namespace mvhmr

using namespace mvhmr;

extern "C" int mvhmr_unproject_aggregate_backward(const float *grad_out, const void *feats, int feat_dtype,
                                                  const float *proj, const float *coord, float *grad_feats,
                                                  int B, int V, int C, int H, int W, long long N, int method, void *stream)
{
    if (method < MVHMR_SUM || method > MVHMR_SOFTMAX)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "Unknown aggregation_method: %d", method);
    if (feat_dtype != MVHMR_F32 && feat_dtype != MVHMR_BF16)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: unknown feat_dtype %d", feat_dtype);
    if (B < 0 || V < 1 || C < 1 || H < 1 || W < 1 || N < 0)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: bad shape B=%d V=%d C=%d H=%d W=%d N=%lld", B, V, C, H, W, N);
    if (V > kBwdMaxViews)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: V=%d exceeds %d", V, kBwdMaxViews);
    if (B == 0 || N == 0) return MVHMR_OK;
    if (B > 65535) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: B=%d exceeds 65535", B);
    if (!grad_out || !feats || !proj || !coord || !grad_feats)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: null pointer");
    dim3 grid((unsigned)((N + 127) / 128), B);
    if (feat_dtype == MVHMR_BF16)
        unproject_backward_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(grad_out, feats, proj, coord, grad_feats, V, C, H, W, N, method);
    else
        unproject_backward_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(grad_out, feats, proj, coord, grad_feats, V, C, H, W, N, method);
    return check_launch("unproject_backward_kernel");
}
