// Reduced-precision fast path of the fused unproject + aggregate (models/aggregation.py:20-87)
// for callers that accept BASELINE.json's bf16 tolerance (1e-2): the bilinear sample is taken by
// the TEXTURE UNITS.
//
// The exact kernels spend most of their time moving four corner texels per voxel-channel-view
// through the L1 data pipe and blending them on the FMA pipe.  The texture units do the address
// arithmetic, the four fetches and the blend in hardware and return one filtered value: a quarter
// of the bytes through the L1 data path and no blend instruction at all.  Their interpolation
// weights are 1.8 fixed point (hence "not usable" for the 1e-5 fp32 contract),
// and the texels are fp16: measured error against the reference's fp32 blend on N(0,1) feature
// maps 4e-3 relative (scripts/micro/tex_bench.cu) — inside the bf16 budget, far outside the fp32 one.
// So this path is OPT-IN (`unprojection(..., precision="fast")`), never the default.
//
//   * tex_pack_kernel: (B,V,C,H,W) bf16 / fp32 -> fp16 planes of 4-channel texels (half4), one
//     plane per (sample, view, channel quad), stacked vertically in a pitch-linear 2-D image with
//     one zero row after every plane (vertical zeros padding; border addressing gives the
//     horizontal one).  bf16 -> fp16 is exact for |x| in [6e-5, 65504]; larger magnitudes saturate.
//   * unproject_tex_kernel: one voxel per thread (32 consecutive z per warp: coalesced 128-byte
//     stores, no transposition), the sampling position computed with the same IEEE sequence as
//     everywhere else (sample_position), V * C/4 tex2D<float4> per voxel, view fusion in fp32
//     registers (Fuse2, the reference's order).  ~48 registers: 40+ resident warps hide the
//     texture latency.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include "mvhmr_common.cuh"
#include "unproject_device.cuh"

namespace mvhmr {

constexpr int kTexMaxGroups = 64;          // texture objects per launch
constexpr int kTexMaxRows = 8192;          // rows per texture: keeps the fp32 row coordinate exact to 2^-10
constexpr int kTexThreads = 256;

struct TexParams {
    UnprojParams u;
    cudaTextureObject_t tex[kTexMaxGroups];
    int samples_per_tex;   // samples stacked in one texture
    int nq;                // channel quads = ceil(C / 4)
    int g0;                // first sample of tex[0] (absolute sample index)
    unsigned nzseg, nxb;   // z segments of 32 voxels; x blocks of 8 planes
    unsigned nyb;          // y blocks
    unsigned lane_axes;    // lane -> voxel mapping: 2 bits per lane bit (0 = z, 1 = x, 2 = y), lane bit 0 first.
                           // The texture unit filters quads of 4 lanes: compact quads (2 z x 2 x) share texels
    int zspan, xspan, yspan;   // voxels of one warp along each axis (product 32)
};

// blockDim = (32, 8): 8 rows of one (map, channel quad) plane per block, lanes along x
template <bool BF16>
__global__ void __launch_bounds__(256)
tex_pack_kernel(const void *__restrict__ feats, __half *__restrict__ planes, int C, int H, int W, int nq, size_t pitch_h)
{
    const int rb = (H + 1 + 7) / 8;                  // row blocks per plane
    const int plane = blockIdx.x / rb;               // bv * nq + q
    const int y = (blockIdx.x - plane * rb) * 8 + threadIdx.y;
    if (y > H) return;
    const int q = plane % nq, bv = plane / nq;
    __half *dst = planes + ((size_t)plane * (H + 1) + y) * pitch_h;
    const size_t chan = (size_t)H * W;
    const size_t src0 = ((size_t)(bv * C + 4 * q) * H + y) * W;
    for (int x = threadIdx.x; x < W; x += 32) {
        float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (y < H) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (4 * q + i < C) {
                    const size_t src = src0 + i * chan + x;
                    v[i] = BF16 ? __bfloat162float(static_cast<const __nv_bfloat16 *>(feats)[src]) : static_cast<const float *>(feats)[src];
                }
        }
        __half2 lo = __floats2half2_rn(fminf(fmaxf(v[0], -65504.0f), 65504.0f), fminf(fmaxf(v[1], -65504.0f), 65504.0f));
        __half2 hi = __floats2half2_rn(fminf(fmaxf(v[2], -65504.0f), 65504.0f), fminf(fmaxf(v[3], -65504.0f), 65504.0f));
        uint2 w;
        w.x = *reinterpret_cast<unsigned *>(&lo); w.y = *reinterpret_cast<unsigned *>(&hi);
        *reinterpret_cast<uint2 *>(dst + 4 * x) = w;
    }
}

// Vector form for rows made of whole, 16-byte aligned groups of 8 (bf16) / 4 (fp32) pixels: a thread loads
// one 16-byte run of each of the quad's four channels and writes the 8 / 4 texels as one contiguous run.
template <bool BF16>
__global__ void __launch_bounds__(256)
tex_pack_vec_kernel(const void *__restrict__ feats, __half *__restrict__ planes, int C, int H, int W, int nq, size_t pitch_h,
                    int groups_per_row, long long items)
{
    constexpr int PX = BF16 ? 8 : 4;                 // pixels per 16-byte load
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(it % groups_per_row);
        const long long row = it / groups_per_row;   // (bv * nq + q) * (H + 1) + y
        const int y = (int)(row % (H + 1));
        const long long plane = row / (H + 1);
        const int q = (int)(plane % nq);
        const long long bv = plane / nq;
        float v[4][PX];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 u = make_uint4(0u, 0u, 0u, 0u);
            if (y < H && 4 * q + i < C) {
                const size_t src = ((size_t)(bv * C + 4 * q + i) * H + y) * W + (size_t)g * PX;
                u = __ldg(reinterpret_cast<const uint4 *>(BF16 ? (const void *)(static_cast<const unsigned short *>(feats) + src)
                                                               : (const void *)(static_cast<const float *>(feats) + src)));
            }
            if (BF16) {
                const unsigned w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) { v[i][2 * k] = __uint_as_float(w[k] << 16); v[i][2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u); }
            } else {
                v[i][0] = __uint_as_float(u.x); v[i][1] = __uint_as_float(u.y); v[i][2] = __uint_as_float(u.z); v[i][3] = __uint_as_float(u.w);
            }
        }
        uint2 *dst = reinterpret_cast<uint2 *>(planes + (size_t)row * pitch_h) + (size_t)g * PX;
#pragma unroll
        for (int k = 0; k < PX; k += 2) {            // two texels = one 16-byte store
            unsigned h[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                __half2 lo = __floats2half2_rn(fminf(fmaxf(v[0][k + j], -65504.0f), 65504.0f), fminf(fmaxf(v[1][k + j], -65504.0f), 65504.0f));
                __half2 hi = __floats2half2_rn(fminf(fmaxf(v[2][k + j], -65504.0f), 65504.0f), fminf(fmaxf(v[3][k + j], -65504.0f), 65504.0f));
                h[2 * j] = *reinterpret_cast<unsigned *>(&lo); h[2 * j + 1] = *reinterpret_cast<unsigned *>(&hi);
            }
            *reinterpret_cast<uint4 *>(dst + k) = make_uint4(h[0], h[1], h[2], h[3]);
        }
    }
}

// Plain (not bit-exact) sampling position for the reduced-precision path: the same formulas as
// sample_position evaluated with ordinary FMAs and one Newton-refined reciprocal.  Differs from the exact
// sequence by ~1e-3 px at most — two orders below the texture units' own 1/512 px weight quantisation.
__device__ __forceinline__ f2 sample_position_fast(const float4 &P0, const float4 &P1, const float4 &P2,
                                                    float X, float Y, float Z, const UnprojParams &p, bool &invalid)
{
    const float xw = fmaf(X, P0.x, fmaf(Y, P0.y, fmaf(Z, P0.z, P0.w)));
    const float yw = fmaf(X, P1.x, fmaf(Y, P1.y, fmaf(Z, P1.z, P1.w)));
    const float ww = fmaf(X, P2.x, fmaf(Y, P2.y, fmaf(Z, P2.z, P2.w)));
    invalid = ww <= 0.0f;
    const float wd = (ww == 0.0f) ? 1.0f : ww;
    float r = rcp_approx(wd);
    r = fmaf(r, fmaf(-wd, r, 1.0f), r);
    // ix = ((2 (x/H - 0.5) + 1) (W-1)/2 = x (W-1)/H ; iy = y (H-1)/W
    f2 i;
    i.x = xw * r * (2.0f * p.sx * p.rH);
    i.y = yw * r * (2.0f * p.sy * p.rW);
    return i;
}

// A / S with the reciprocal on the FMA pipe: the XU (MUFU) is this kernel's busiest pipe (157 -> 154 us at cfg3).  S is in [1, V]: magic-constant seed (12 % off), two Newton steps -> 2e-4 relative, far inside the path's budget.
__device__ __forceinline__ f2 softmax_result_fma(u64 A, u64 S)
{
    const f2 s = upk(S);
    u64 r = pk(__int_as_float(0x7EF311C3 - __float_as_int(s.x)), __int_as_float(0x7EF311C3 - __float_as_int(s.y)));
    const u64 two = pk(2.0f, 2.0f), ns = pk(-s.x, -s.y);
    r = mul2(r, fma2(ns, r, two));
    r = mul2(r, fma2(ns, r, two));
    return upk(mul2(A, r));
}

// EXACT: V == VMAX (no per-view guards); FULLC: C % 4 == 0 (no per-channel guards)
template <int VMAX, int METHOD, bool EXACT, bool FULLC>
__global__ void __launch_bounds__(kTexThreads, VMAX > 4 ? 4 : 5)   // V <= 4: five CTAs per SM (44 registers); V = 8: four (64 registers) beat five with spills, 430 vs 467 us at cfg5's shape
unproject_tex_kernel(const TexParams q)
{
    const UnprojParams &p = q.u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // block -> (sample, z segment, x block, y block); lane -> voxel inside the warp's zspan x xspan x yspan brick
    unsigned t = blockIdx.x;
    const unsigned yb = t % q.nyb; t /= q.nyb;
    const unsigned xb = t % q.nxb; t /= q.nxb;
    const int seg = (int)(t % q.nzseg);
    const int b = p.b0 + (int)(t / q.nzseg);
    int dz = 0, dx = 0, dy = 0;
    {
        int sz = 0, sx = 0, sy = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const unsigned ax = (q.lane_axes >> (2 * i)) & 3u, bit = (lane >> i) & 1u;
            if (ax == 0) dz |= bit << sz++;
            else if (ax == 1) dx |= bit << sx++;
            else dy |= bit << sy++;
        }
    }
    const int xi = ((int)xb * (kTexThreads / 32) + warp) * q.xspan + dx;
    const int vy = (int)yb * q.yspan + dy, vz = seg * q.zspan + dz;
    const int vx = p.x_lo + xi;
    const long long n = ((long long)vx * p.gy + vy) * p.gz + vz;
    const bool mine = xi < p.nx && vy < p.gy && vz < p.gz && n >= p.n0 && n < p.n1;
    if (!__any_sync(0xffffffffu, mine)) return;

    // ---- sampling positions of this voxel in every view (exact), as texture coordinates ----
    float X = 0.0f, Y = 0.0f, Z = 0.0f;
    if (mine) {
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
    const int bl = b - q.g0;                                   // sample index relative to tex[0]
    const cudaTextureObject_t tex = q.tex[bl / q.samples_per_tex];
    const int rows_per_view = q.nq * (p.H + 1);
    const float row_sample = (float)((bl % q.samples_per_tex) * p.V * rows_per_view);
    float xt[VMAX], yt[VMAX];
    bool nanpos = false;                                       // some view's position is not finite: the reference yields NaN
    const float4 *Pb = reinterpret_cast<const float4 *>(p.proj + (size_t)b * p.V * 12);
#pragma unroll
    for (int v = 0; v < VMAX; ++v) {
        xt[v] = 0.0f; yt[v] = -1e9f;
        if (EXACT || v < p.V) {
            const float4 P0 = __ldg(Pb + 3 * v), P1 = __ldg(Pb + 3 * v + 1), P2 = __ldg(Pb + 3 * v + 2);
            bool invalid;
            const f2 i = sample_position_fast(P0, P1, P2, X, Y, Z, p, invalid);
            const bool finite = fabsf(i.x) < INFINITY && fabsf(i.y) < INFINITY;
            nanpos = nanpos || (!finite && !invalid);
            // rows beyond the plane contribute nothing: clamp onto the zero rows around it.  depth <= 0
            // (an exact 0 in the reference) and non-finite positions sample far above the whole image:
            // border addressing returns zeros there, whatever row offset is added later
            const float iy = fminf(fmaxf(i.y, -1.0f), p.Hf);
            const float ix = fminf(fmaxf(i.x, -2.0f), p.Wf + 1.0f);
            xt[v] = ix + 0.5f;
            yt[v] = (invalid || !finite) ? -1e9f : iy + 0.5f + row_sample + (float)(v * rows_per_view);
        }
    }
    const bool anynan = __any_sync(0xffffffffu, nanpos);
    const float Vf = (float)p.V;
    const float plane_rows = (float)(p.H + 1);
    float *o = p.out + (size_t)b * p.C * p.n_extent + (mine ? n - p.n_origin : 0);
    const size_t cs = (size_t)p.n_extent;
    float yoff = 0.0f;
    for (int c4 = 0; c4 < q.nq; ++c4, yoff += plane_rows, o += 4 * cs) {
        u64 s[VMAX][2];
#pragma unroll
        for (int v = 0; v < VMAX; ++v) {
            if (EXACT || v < p.V) {
                const float4 tx = tex2D<float4>(tex, xt[v], yt[v] + yoff);
                s[v][0] = pk(tx.x, tx.y); s[v][1] = pk(tx.z, tx.w);
            }
        }
        Fuse2<METHOD, VMAX, EXACT> f0, f1;
        f0.absorb(&s[0][0], 2, p.V, true);
        f1.absorb(&s[0][1], 2, p.V, true);
        f2 r0, r1;
        if (METHOD == MVHMR_SOFTMAX) { r0 = softmax_result_fma(f0.a, f0.S); r1 = softmax_result_fma(f1.a, f1.S); }
        else { r0 = f0.result(Vf); r1 = f1.result(Vf); }
        if (anynan && nanpos) { r0.x = r0.y = r1.x = r1.y = __int_as_float(0x7fc00000); }   // NaN propagates through every fusion mode
        if (mine) {
            __stcs(o, r0.x);
            if (FULLC || 4 * c4 + 1 < p.C) __stcs(o + cs, r0.y);
            if (FULLC || 4 * c4 + 2 < p.C) __stcs(o + 2 * cs, r1.x);
            if (FULLC || 4 * c4 + 3 < p.C) __stcs(o + 3 * cs, r1.y);
        }
    }
}

// ---- texture objects: small cache keyed by (pointer, geometry); a texture object is only a
// descriptor of memory the caller owns, so a stale entry is harmless ----
struct TexKey { const void *ptr; int W, rows; size_t pitch; int dev; cudaTextureObject_t obj; };
static std::mutex g_tex_mutex;
static std::vector<TexKey> g_tex_cache;

static int get_texture(const void *ptr, int W, int rows, size_t pitch, cudaTextureObject_t *out)
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_tex_mutex);
    for (const TexKey &k : g_tex_cache)
        if (k.ptr == ptr && k.W == W && k.rows == rows && k.pitch == pitch && k.dev == dev) { *out = k.obj; return MVHMR_OK; }
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = const_cast<void *>(ptr);
    rd.res.pitch2D.desc = cudaCreateChannelDescHalf4();
    rd.res.pitch2D.width = (size_t)W;
    rd.res.pitch2D.height = (size_t)rows;
    rd.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;      // zeros padding
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t obj = 0;
    cudaError_t e = cudaCreateTextureObject(&obj, &rd, &td, nullptr);
    if (e != cudaSuccess) return fail(MVHMR_ERR_CUDA, "unproject_aggregate_tex: cudaCreateTextureObject: %s", cudaGetErrorString(e));
    if (g_tex_cache.size() >= 256) {                                      // oldest entry: long out of use
        cudaDestroyTextureObject(g_tex_cache.front().obj);
        g_tex_cache.erase(g_tex_cache.begin());
    }
    g_tex_cache.push_back({ptr, W, rows, pitch, dev, obj});
    *out = obj;
    return MVHMR_OK;
}

static size_t tex_pitch(int W) { return ((size_t)W * 8 + 511) & ~(size_t)511; }   // row pitch: 512 B keeps every texture base aligned

}  // namespace mvhmr

using namespace mvhmr;

extern "C" size_t mvhmr_unproject_tex_workspace_bytes(int B, int V, int C, int H, int W)
{
    if (B < 0 || V < 1 || C < 1 || H < 1 || W < 1) return 0;
    return (size_t)B * V * ((C + 3) / 4) * (H + 1) * tex_pitch(W) + 512;
}

extern "C" int mvhmr_unproject_aggregate_tex(const void *feats, int feat_dtype,
                                             const float *proj, const float *coord, const mvhmr_grid_t *grid, float *out,
                                             int B, int V, int C, int H, int W,
                                             int gx, int gy, int gz, int method,
                                             int b0, int b1, long long n0, long long n1,
                                             long long n_origin, long long n_extent,
                                             void *ws, size_t ws_bytes, void *stream)
{
    if (method < MVHMR_SUM || method > MVHMR_SOFTMAX)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "Unknown aggregation_method: %d", method);
    if (feat_dtype != MVHMR_F32 && feat_dtype != MVHMR_BF16)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_tex: unknown feat_dtype %d", feat_dtype);
    if (B < 0 || V < 1 || C < 1 || H < 1 || W < 1 || gx < 1 || gy < 1 || gz < 1)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_tex: bad shape B=%d V=%d C=%d H=%d W=%d G=(%d,%d,%d)", B, V, C, H, W, gx, gy, gz);
    if (V > 8) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_tex: the texture path takes up to 8 views, got %d (use the exact path)", V);
    const int nq = (C + 3) / 4;
    const long long rows_per_sample = (long long)V * nq * (H + 1);
    if (rows_per_sample > kTexMaxRows || W > 65536)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_tex: V*ceil(C/4)*(H+1) = %lld rows per sample exceed %d (use the exact path)", rows_per_sample, kTexMaxRows);
    const long long N = (long long)gx * gy * gz;
    if (b0 < 0 || b1 > B || b0 > b1 || n0 < 0 || n1 > N || n0 > n1)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_tex: shard window [%d,%d)x[%lld,%lld) outside B=%d N=%lld", b0, b1, n0, n1, B, N);
    if (n_origin < 0 || n_extent < 0 || n0 < n_origin || n1 > n_origin + n_extent)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_tex: voxels [%lld,%lld) not inside the buffers' range", n0, n1);
    if (b0 == b1 || n0 == n1) return MVHMR_OK;
    if (!feats || !proj || !out || (!coord && !grid)) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_tex: null pointer");
    if (!coord && (!grid->centers || !grid->rot)) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_tex: null centers / rot");
    if ((uintptr_t)proj & 15) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_tex: proj must be 16-byte aligned");
    const size_t need = mvhmr_unproject_tex_workspace_bytes(B, V, C, H, W);
    if (!ws || ws_bytes < need)
        return fail(MVHMR_ERR_WORKSPACE, "unproject_aggregate_tex: workspace of %zu bytes required, got %zu", need, ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    const bool bf = feat_dtype == MVHMR_BF16;
    char *planes = (char *)(((uintptr_t)ws + 511) & ~(uintptr_t)511);
    const size_t pitch = tex_pitch(W);
    const size_t sample_bytes = (size_t)rows_per_sample * pitch;

    // fp16 planes of the window's samples
    {
        const size_t in_sample = (size_t)V * C * H * W * (bf ? 2 : 4);
        const long long blocks = (long long)(b1 - b0) * V * nq * ((H + 1 + 7) / 8);
        if (blocks > 0x7fffffffLL) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_tex: too many rows");
        const void *src = (const char *)feats + in_sample * b0;
        __half *dst = (__half *)(planes + sample_bytes * b0);
        const int px = bf ? 8 : 4;
        if (W % px == 0 && ((uintptr_t)src & 15) == 0) {
            const int gpr = W / px;
            const long long items = (long long)(b1 - b0) * rows_per_sample * gpr;
            const unsigned grid = (unsigned)((items + 255) / 256 < 148LL * 16 ? (items + 255) / 256 : 148LL * 16);
            if (bf) tex_pack_vec_kernel<true><<<grid, 256, 0, st>>>(src, dst, C, H, W, nq, pitch / 2, gpr, items);
            else tex_pack_vec_kernel<false><<<grid, 256, 0, st>>>(src, dst, C, H, W, nq, pitch / 2, gpr, items);
        } else if (bf) tex_pack_kernel<true><<<(unsigned)blocks, dim3(32, 8), 0, st>>>(src, dst, C, H, W, nq, pitch / 2);
        else tex_pack_kernel<false><<<(unsigned)blocks, dim3(32, 8), 0, st>>>(src, dst, C, H, W, nq, pitch / 2);
        int rc = check_launch("tex_pack_kernel");
        if (rc != MVHMR_OK) return rc;
    }

    TexParams q;
    UnprojParams &p = q.u;
    p.proj = proj; p.coord = coord; p.out = out;
    p.centers = coord ? nullptr : grid->centers;
    p.rot = coord ? nullptr : grid->rot;
    for (int k = 0; k < 3; ++k) {
        p.gpos[k] = coord ? 0.0f : grid->pos[k];
        p.gstep[k] = coord ? 0.0f : grid->step[k];
    }
    p.n0 = n0; p.n1 = n1; p.n_origin = n_origin; p.n_extent = n_extent;
    p.V = V; p.C = C; p.W = W; p.H = H;
    p.gx = gx; p.gy = gy; p.gz = gz;
    const long long yz = (long long)gy * gz;
    p.x_lo = (int)(n0 / yz);
    p.nx = (int)((n1 - 1) / yz) - p.x_lo + 1;
    p.Hf = (float)H; p.Wf = (float)W;
    p.sx = (float)(W - 1) / 2.0f; p.sy = (float)(H - 1) / 2.0f;
    p.rH = 1.0f / (float)H; p.rW = 1.0f / (float)W;
    q.nq = nq;
    q.samples_per_tex = (int)(kTexMaxRows / rows_per_sample);
    {
        // lane bits -> axes, lane bit 0 first ('z' / 'x' / 'y').  Default zxzzz: a warp covers 16 z of two
        // x planes and every texture quad is 2 z x 2 x (MVHMR_TEX_LANES overrides: tuning knob)
        const char *spec = getenv("MVHMR_TEX_LANES");
        if (!spec || strlen(spec) != 5) spec = "zxzzz";
        q.lane_axes = 0; q.zspan = q.xspan = q.yspan = 1;
        for (int i = 0; i < 5; ++i) {
            const unsigned ax = spec[i] == 'x' ? 1u : spec[i] == 'y' ? 2u : 0u;
            q.lane_axes |= ax << (2 * i);
            (ax == 0 ? q.zspan : ax == 1 ? q.xspan : q.yspan) *= 2;
        }
    }
    q.nzseg = (unsigned)((gz + q.zspan - 1) / q.zspan);
    q.nxb = (unsigned)((p.nx + q.xspan * (kTexThreads / 32) - 1) / (q.xspan * (kTexThreads / 32)));
    q.nyb = (unsigned)((gy + q.yspan - 1) / q.yspan);
    const int per_launch = q.samples_per_tex * kTexMaxGroups;
    for (int g0 = b0; g0 < b1; g0 += per_launch) {
        const int g1 = g0 + per_launch < b1 ? g0 + per_launch : b1;
        q.g0 = g0; p.b0 = g0; p.nb = g1 - g0;
        for (int k = 0; k < kTexMaxGroups; ++k) q.tex[k] = 0;
        for (int s0 = g0, k = 0; s0 < g1; s0 += q.samples_per_tex, ++k) {
            const int ns = s0 + q.samples_per_tex < g1 ? q.samples_per_tex : g1 - s0;
            int rc = get_texture(planes + sample_bytes * s0, W, (int)(ns * rows_per_sample), pitch, &q.tex[k]);
            if (rc != MVHMR_OK) return rc;
        }
        const long long blocks = (long long)p.nb * q.nzseg * q.nxb * q.nyb;
        if (blocks > 0x7fffffffLL) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_tex: too many blocks in one call");
#define MVHMR_TLAUNCH(VM, M, EX, FC) unproject_tex_kernel<VM, M, EX, FC><<<(unsigned)blocks, kTexThreads, 0, st>>>(q)
#define MVHMR_TMETHOD(VM, EX, FC)                                        \
        switch (method) {                                                \
        case MVHMR_SUM: MVHMR_TLAUNCH(VM, MVHMR_SUM, EX, FC); break;     \
        case MVHMR_MEAN: MVHMR_TLAUNCH(VM, MVHMR_MEAN, EX, FC); break;   \
        case MVHMR_MAX: MVHMR_TLAUNCH(VM, MVHMR_MAX, EX, FC); break;     \
        default: MVHMR_TLAUNCH(VM, MVHMR_SOFTMAX, EX, FC); break;        \
        }
        const bool fullc = C % 4 == 0;
        if (V == 4 && fullc) { MVHMR_TMETHOD(4, true, true) }
        else if (V == 8 && fullc) { MVHMR_TMETHOD(8, true, true) }
        else if (V <= 4) { MVHMR_TMETHOD(4, false, false) }
        else { MVHMR_TMETHOD(8, false, false) }
#undef MVHMR_TMETHOD
#undef MVHMR_TLAUNCH
        int rc = check_launch("unproject_tex_kernel");
        if (rc != MVHMR_OK) return rc;
    }
    return MVHMR_OK;
}
