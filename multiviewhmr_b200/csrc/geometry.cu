// Grid / point geometry kernels: coord-volume builder, point rotation, 3x4
// projection.  All three are pure streaming kernels (12 B in / 12 B out per
// point) bounded by HBM; every arithmetic step is an explicit IEEE intrinsic so
// that nvcc cannot contract or reorder it — the results are bit-identical to
// the reference's CPU torch path.
#include "mvhmr_common.cuh"

namespace mvhmr {

struct GridParams {
    float pos[3];
    float step[3];
    int Gx, Gy, Gz;
};

// Four consecutive voxels (along z, with carry into y and x) per thread: the index is decomposed
// once per thread with 32-bit divisions and the 48 bytes leave as three 16-byte stores when the
// sample's slab is 16-byte aligned (N % 4 == 0).  models/aggregation.py:140-161,184-187.
template <bool VEC>
__global__ void __launch_bounds__(256)
coord_volume_kernel(float *__restrict__ out, const float *__restrict__ centers,
                    const float *__restrict__ rot, GridParams g)
{
    const int b = blockIdx.y;
    const long long N = (long long)g.Gx * g.Gy * g.Gz;
    const float c0 = __ldg(centers + 3 * b), c1 = __ldg(centers + 3 * b + 1), c2 = __ldg(centers + 3 * b + 2);
    float R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = __ldg(rot + 9 * b + i);
    const long long nquads = (N + 3) >> 2;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nquads;
         q += (long long)gridDim.x * blockDim.x) {
        const long long n = q << 2;
        int ix, iy, iz;
        if (N <= 0x7fffffffLL) {
            const unsigned t = (unsigned)n / (unsigned)g.Gz;
            iz = (int)((unsigned)n - t * (unsigned)g.Gz);
            ix = (int)(t / (unsigned)g.Gy);
            iy = (int)(t - (unsigned)ix * (unsigned)g.Gy);
        } else {
            iz = (int)(n % g.Gz);
            const long long t = n / g.Gz;
            iy = (int)(t % g.Gy);
            ix = (int)(t / g.Gy);
        }
        float v[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // f32(pos) + f32(step) * f32(idx): product and sum round separately
            const float d0 = __fsub_rn(__fadd_rn(g.pos[0], __fmul_rn(g.step[0], (float)ix)), c0);
            const float d1 = __fsub_rn(__fadd_rn(g.pos[1], __fmul_rn(g.step[1], (float)iy)), c1);
            const float d2 = __fsub_rn(__fadd_rn(g.pos[2], __fmul_rn(g.step[2], (float)iz)), c2);
            v[3 * k] = __fadd_rn(rot_row(R[0], R[1], R[2], d0, d1, d2), c0);
            v[3 * k + 1] = __fadd_rn(rot_row(R[3], R[4], R[5], d0, d1, d2), c1);
            v[3 * k + 2] = __fadd_rn(rot_row(R[6], R[7], R[8], d0, d1, d2), c2);
            if (++iz == g.Gz) { iz = 0; if (++iy == g.Gy) { iy = 0; ++ix; } }
        }
        float *o = out + ((size_t)b * N + n) * 3;
        if (VEC) {                                   // N % 4 == 0: whole quads, 16-byte aligned
            float4 *o4 = reinterpret_cast<float4 *>(o);
            __stcs(o4, make_float4(v[0], v[1], v[2], v[3]));
            __stcs(o4 + 1, make_float4(v[4], v[5], v[6], v[7]));
            __stcs(o4 + 2, make_float4(v[8], v[9], v[10], v[11]));
        } else {
            const int cnt = (int)min(4LL, N - n) * 3;
#pragma unroll
            for (int i = 0; i < 12; ++i) if (i < cnt) o[i] = v[i];
        }
    }
}

struct Rot3 { float r[9]; };

// Point streams: four points (48 bytes) per thread as three 16-byte loads / stores when the
// buffers are 16-byte aligned; the last N % 4 points and unaligned buffers go one by one.
template <bool VEC>
__global__ void __launch_bounds__(256)
rotate_points_kernel(float *__restrict__ out, const float *__restrict__ pts, Rot3 R, size_t N)
{
    const size_t nquads = VEC ? N >> 2 : 0;
    if (VEC)
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += (size_t)gridDim.x * blockDim.x) {
        const float4 *in4 = reinterpret_cast<const float4 *>(pts) + 3 * q;
        const float4 a = __ldcs(in4), c = __ldcs(in4 + 1), e = __ldcs(in4 + 2);
        const float d[12] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w, e.x, e.y, e.z, e.w};
        float o[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[3 * k] = rot_row(R.r[0], R.r[1], R.r[2], d[3 * k], d[3 * k + 1], d[3 * k + 2]);
            o[3 * k + 1] = rot_row(R.r[3], R.r[4], R.r[5], d[3 * k], d[3 * k + 1], d[3 * k + 2]);
            o[3 * k + 2] = rot_row(R.r[6], R.r[7], R.r[8], d[3 * k], d[3 * k + 1], d[3 * k + 2]);
        }
        float4 *o4 = reinterpret_cast<float4 *>(out) + 3 * q;
        __stcs(o4, make_float4(o[0], o[1], o[2], o[3]));
        __stcs(o4 + 1, make_float4(o[4], o[5], o[6], o[7]));
        __stcs(o4 + 2, make_float4(o[8], o[9], o[10], o[11]));
    }
    for (size_t n = (nquads << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
        const float d0 = pts[3 * n], d1 = pts[3 * n + 1], d2 = pts[3 * n + 2];
        const float o0 = rot_row(R.r[0], R.r[1], R.r[2], d0, d1, d2);
        const float o1 = rot_row(R.r[3], R.r[4], R.r[5], d0, d1, d2);
        const float o2 = rot_row(R.r[6], R.r[7], R.r[8], d0, d1, d2);
        out[3 * n] = o0; out[3 * n + 1] = o1; out[3 * n + 2] = o2;
    }
}

template <bool VEC>
__global__ void __launch_bounds__(256)
project_points_kernel(float *__restrict__ out, const float *__restrict__ P, const float *__restrict__ pts,
                      size_t N, int euclid)
{
    float p[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) p[i] = __ldg(P + i);
    const size_t nquads = VEC ? N >> 2 : 0;
    if (VEC)
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += (size_t)gridDim.x * blockDim.x) {
        const float4 *in4 = reinterpret_cast<const float4 *>(pts) + 3 * q;
        const float4 a = __ldcs(in4), c = __ldcs(in4 + 1), e = __ldcs(in4 + 2);
        const float d[12] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w, e.x, e.y, e.z, e.w};
        float xw[4], yw[4], w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            xw[k] = proj_row(d[3 * k], d[3 * k + 1], d[3 * k + 2], p[0], p[1], p[2], p[3]);
            yw[k] = proj_row(d[3 * k], d[3 * k + 1], d[3 * k + 2], p[4], p[5], p[6], p[7]);
            w[k] = proj_row(d[3 * k], d[3 * k + 1], d[3 * k + 2], p[8], p[9], p[10], p[11]);
        }
        if (euclid) {
            float4 *o4 = reinterpret_cast<float4 *>(out) + 2 * q;
            __stcs(o4, make_float4(__fdiv_rn(xw[0], w[0]), __fdiv_rn(yw[0], w[0]), __fdiv_rn(xw[1], w[1]), __fdiv_rn(yw[1], w[1])));
            __stcs(o4 + 1, make_float4(__fdiv_rn(xw[2], w[2]), __fdiv_rn(yw[2], w[2]), __fdiv_rn(xw[3], w[3]), __fdiv_rn(yw[3], w[3])));
        } else {
            float4 *o4 = reinterpret_cast<float4 *>(out) + 3 * q;
            __stcs(o4, make_float4(xw[0], yw[0], w[0], xw[1]));
            __stcs(o4 + 1, make_float4(yw[1], w[1], xw[2], yw[2]));
            __stcs(o4 + 2, make_float4(w[2], xw[3], yw[3], w[3]));
        }
    }
    for (size_t n = (nquads << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
        const float X = pts[3 * n], Y = pts[3 * n + 1], Z = pts[3 * n + 2];
        const float xw = proj_row(X, Y, Z, p[0], p[1], p[2], p[3]);
        const float yw = proj_row(X, Y, Z, p[4], p[5], p[6], p[7]);
        const float w = proj_row(X, Y, Z, p[8], p[9], p[10], p[11]);
        if (euclid) {
            out[2 * n] = __fdiv_rn(xw, w);
            out[2 * n + 1] = __fdiv_rn(yw, w);
        } else {
            out[3 * n] = xw; out[3 * n + 1] = yw; out[3 * n + 2] = w;
        }
    }
}

static unsigned stream_grid(size_t n, int block) {
    size_t g = (n + block - 1) / block;
    const size_t cap = 148u * 16u;          // 16 CTAs of 256 threads per SM-equivalent
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace mvhmr

using namespace mvhmr;

extern "C" int mvhmr_build_coord_volumes(float *out, const float *centers, const float *rot,
                                         const float *pos_host, const float *step_host,
                                         int B, int Gx, int Gy, int Gz, void *stream)
{
    if (B < 0 || Gx < 1 || Gy < 1 || Gz < 1)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "build_coord_volumes: bad shape B=%d G=(%d,%d,%d)", B, Gx, Gy, Gz);
    if (B == 0) return MVHMR_OK;
    if (!out || !centers || !rot || !pos_host || !step_host)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "build_coord_volumes: null pointer");
    if (B > 65535) return fail(MVHMR_ERR_INVALID_ARGUMENT, "build_coord_volumes: B=%d exceeds 65535", B);
    GridParams g;
    for (int k = 0; k < 3; ++k) { g.pos[k] = pos_host[k]; g.step[k] = step_host[k]; }
    g.Gx = Gx; g.Gy = Gy; g.Gz = Gz;
    const size_t N = (size_t)Gx * Gy * Gz;
    dim3 grid(stream_grid((N + 3) / 4, 256), B);
    if (N % 4 == 0 && ((uintptr_t)out & 15) == 0)
        coord_volume_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(out, centers, rot, g);
    else
        coord_volume_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(out, centers, rot, g);
    return check_launch("coord_volume_kernel");
}

extern "C" int mvhmr_rotate_points(float *out, const float *pts, const float *rot_host, size_t N, void *stream)
{
    if (N == 0) return MVHMR_OK;
    if (!out || !pts || !rot_host) return fail(MVHMR_ERR_INVALID_ARGUMENT, "rotate_points: null pointer");
    Rot3 R;
    for (int i = 0; i < 9; ++i) R.r[i] = rot_host[i];
    if ((((uintptr_t)out | (uintptr_t)pts) & 15) == 0 && N >= 4)
        rotate_points_kernel<true><<<stream_grid((N + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(out, pts, R, N);
    else
        rotate_points_kernel<false><<<stream_grid(N, 256), 256, 0, (cudaStream_t)stream>>>(out, pts, R, N);
    return check_launch("rotate_points_kernel");
}

extern "C" int mvhmr_project_points(float *out, const float *P, const float *pts, size_t N, int euclid, void *stream)
{
    if (N == 0) return MVHMR_OK;
    if (!out || !P || !pts) return fail(MVHMR_ERR_INVALID_ARGUMENT, "project_points: null pointer");
    if ((((uintptr_t)out | (uintptr_t)pts) & 15) == 0 && N >= 4)
        project_points_kernel<true><<<stream_grid((N + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(out, P, pts, N, euclid ? 1 : 0);
    else
        project_points_kernel<false><<<stream_grid(N, 256), 256, 0, (cudaStream_t)stream>>>(out, P, pts, N, euclid ? 1 : 0);
    return check_launch("project_points_kernel");
}
