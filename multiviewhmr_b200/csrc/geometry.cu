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

// One thread per voxel; consecutive threads walk z, so a warp writes 384
// contiguous bytes.  models/aggregation.py:140-161,184-187.
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
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N;
         n += (long long)gridDim.x * blockDim.x) {
        const int iz = (int)(n % g.Gz);
        const long long t = n / g.Gz;
        const int iy = (int)(t % g.Gy);
        const int ix = (int)(t / g.Gy);
        // f32(pos) + f32(step) * f32(idx): product and sum round separately
        const float d0 = __fsub_rn(__fadd_rn(g.pos[0], __fmul_rn(g.step[0], (float)ix)), c0);
        const float d1 = __fsub_rn(__fadd_rn(g.pos[1], __fmul_rn(g.step[1], (float)iy)), c1);
        const float d2 = __fsub_rn(__fadd_rn(g.pos[2], __fmul_rn(g.step[2], (float)iz)), c2);
        float *o = out + ((size_t)b * N + n) * 3;
        o[0] = __fadd_rn(rot_row(R[0], R[1], R[2], d0, d1, d2), c0);
        o[1] = __fadd_rn(rot_row(R[3], R[4], R[5], d0, d1, d2), c1);
        o[2] = __fadd_rn(rot_row(R[6], R[7], R[8], d0, d1, d2), c2);
    }
}

struct Rot3 { float r[9]; };

__global__ void __launch_bounds__(256)
rotate_points_kernel(float *__restrict__ out, const float *__restrict__ pts, Rot3 R, size_t N)
{
    for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
        const float d0 = pts[3 * n], d1 = pts[3 * n + 1], d2 = pts[3 * n + 2];
        const float o0 = rot_row(R.r[0], R.r[1], R.r[2], d0, d1, d2);
        const float o1 = rot_row(R.r[3], R.r[4], R.r[5], d0, d1, d2);
        const float o2 = rot_row(R.r[6], R.r[7], R.r[8], d0, d1, d2);
        out[3 * n] = o0; out[3 * n + 1] = o1; out[3 * n + 2] = o2;
    }
}

__global__ void __launch_bounds__(256)
project_points_kernel(float *__restrict__ out, const float *__restrict__ P, const float *__restrict__ pts,
                      size_t N, int euclid)
{
    float p[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) p[i] = __ldg(P + i);
    for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
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
    dim3 grid(stream_grid(N, 256), B);
    coord_volume_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, centers, rot, g);
    return check_launch("coord_volume_kernel");
}

extern "C" int mvhmr_rotate_points(float *out, const float *pts, const float *rot_host, size_t N, void *stream)
{
    if (N == 0) return MVHMR_OK;
    if (!out || !pts || !rot_host) return fail(MVHMR_ERR_INVALID_ARGUMENT, "rotate_points: null pointer");
    Rot3 R;
    for (int i = 0; i < 9; ++i) R.r[i] = rot_host[i];
    rotate_points_kernel<<<stream_grid(N, 256), 256, 0, (cudaStream_t)stream>>>(out, pts, R, N);
    return check_launch("rotate_points_kernel");
}

extern "C" int mvhmr_project_points(float *out, const float *P, const float *pts, size_t N, int euclid, void *stream)
{
    if (N == 0) return MVHMR_OK;
    if (!out || !P || !pts) return fail(MVHMR_ERR_INVALID_ARGUMENT, "project_points: null pointer");
    project_points_kernel<<<stream_grid(N, 256), 256, 0, (cudaStream_t)stream>>>(out, P, pts, N, euclid ? 1 : 0);
    return check_launch("project_points_kernel");
}
