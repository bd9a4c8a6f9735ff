// Fused unproject + aggregate (models/aggregation.py:20-87) for sm_100a.
//
// Data layout in HBM
//   features arrive NCHW (B,V,C,H,W).  pack_kernel rewrites them once into the
//   gather layout  (B*V, chunks, H+4, W+4) x 16-byte texels, a texel holding 4
//   fp32 (or 8 bf16) consecutive channels of one pixel, each plane surrounded
//   by a 2-texel ZERO border.  With the border, grid_sample's zeros padding is
//   just a clamp of the cell index: no per-corner predicate, no divergent code.
//   Neighbouring pixels are neighbouring 16-byte words, so the 32 lanes of a
//   warp (32 consecutive voxels along z, which project to a short run of
//   pixels) read a handful of 128-byte lines per LDG.128.
//
// unproject_kernel: one thread per voxel, all views, all channels.
//   1. project the voxel with every view's matrix (IEEE ops in the reference's
//      order, so the sampling position is bit-identical to the CPU torch path),
//      keep (texel offset, 4 weights) per view in registers;
//   2. for each group of 8 channels: gather 4 corners x views (all loads are
//      independent -> deep MLP), blend with one mul + three FMAs (== ATen),
//      fuse over views in registers (sum / mean / max / softmax) and write the
//      8 results with one coalesced 128-byte line per channel per warp.
//   No per-view volume is ever materialised; HBM sees the features once (they
//   stay L2-resident, 126 MB), the coordinates once and the output once.
#include <cuda_bf16.h>
#include "mvhmr_common.cuh"

namespace mvhmr {

constexpr int kBorder = 2;
constexpr int kBlock = 256;
constexpr int kGroupCh = 8;            // channels fused per inner iteration
constexpr float kLog2e = 1.4426950408889634f;

struct UnprojParams {
    const uint4 *packed;   // (B*V, chunks, Hp, Wp) 16-byte texels
    const float *proj;     // (B, V, 3, 4)
    const float *coord;    // (B, n_extent, 3): voxels [n_origin, n_origin + n_extent)
    float *out;            // (B, C, n_extent)
    long long n0, n1;      // voxels computed by this launch
    long long n_origin, n_extent;
    long long plane;       // Hp * Wp texels
    int V, C, W, H, Wp, chunks;
    int b0;
    int gx, gy, gz;        // volume shape (gx*gy*gz == N)
    int ltx, lty, ltz;     // log2 of the CTA brick (TX*TY*TZ == kBlock)
    int x_lo;              // first x-plane touched by [n0,n1)
    int tiles_y, tiles_z;
    float Hf, Wf, sx, sy;  // (float)H, (float)W, (W-1)/2, (H-1)/2
};

struct ViewCell {
    int off;               // texel offset of the nw corner inside a padded plane
    float w00, w01, w10, w11;
};

// models/aggregation.py:38-51 + ATen grid_sampler unnormalize/compute_interp_params.
__device__ __forceinline__ ViewCell make_cell(const float *Ps, float X, float Y, float Z, const UnprojParams &p)
{
    const float xw = proj_row(X, Y, Z, Ps[0], Ps[1], Ps[2], Ps[3]);
    const float yw = proj_row(X, Y, Z, Ps[4], Ps[5], Ps[6], Ps[7]);
    const float ww = proj_row(X, Y, Z, Ps[8], Ps[9], Ps[10], Ps[11]);
    const bool invalid = ww <= 0.0f;                 // :42 depth must be > 0
    const float wd = (ww == 0.0f) ? 1.0f : ww;       // :44 not to divide by zero
    const float x = __fdiv_rn(xw, wd);
    const float y = __fdiv_rn(yw, wd);
    // :49-50  2*(x/feature_shape[0] - 0.5): x by H, y by W (reference behaviour)
    const float gx = __fmul_rn(2.0f, __fsub_rn(__fdiv_rn(x, p.Hf), 0.5f));
    const float gy = __fmul_rn(2.0f, __fsub_rn(__fdiv_rn(y, p.Wf), 0.5f));
    // align_corners=True: (g + 1) * ((size - 1) / 2)
    const float ix = __fmul_rn(__fadd_rn(gx, 1.0f), p.sx);
    const float iy = __fmul_rn(__fadd_rn(gy, 1.0f), p.sy);
    const float xf = floorf(ix), yf = floorf(iy);
    const float fw = __fsub_rn(ix, xf), fe = __fsub_rn(1.0f, fw);
    const float fn = __fsub_rn(iy, yf), fs = __fsub_rn(1.0f, fn);
    ViewCell c;
    c.w00 = __fmul_rn(fs, fe); c.w01 = __fmul_rn(fs, fw);
    c.w10 = __fmul_rn(fn, fe); c.w11 = __fmul_rn(fn, fw);
    // clamp the cell into the zero border; NaN/inf positions land in the border
    // too and keep their NaN weights, as 0*NaN does in the reference.
    const int x0 = (int)fminf(fmaxf(xf, -2.0f), p.Wf);
    const int y0 = (int)fminf(fmaxf(yf, -2.0f), p.Hf);
    c.off = (y0 + kBorder) * p.Wp + (x0 + kBorder);
    if (invalid) {                                   // :62 zero out non-valid points
        c.off = 0;                                   // four border texels: exact +0
        c.w00 = c.w01 = c.w10 = c.w11 = 0.0f;
    }
    return c;
}

__device__ __forceinline__ float blend(float t00, float t01, float t10, float t11, const ViewCell &c)
{
    float acc = __fmul_rn(t00, c.w00);
    acc = __fmaf_rn(t01, c.w01, acc);
    acc = __fmaf_rn(t10, c.w10, acc);
    return __fmaf_rn(t11, c.w11, acc);
}

__device__ __forceinline__ float bf_lo(unsigned u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(unsigned u) { return __uint_as_float(u & 0xffff0000u); }

// 8 channels of one view: s[0..7]
template <bool BF16>
__device__ __forceinline__ void sample8(float *s, const uint4 *base, long long plane, int Wp, const ViewCell &c)
{
    if (BF16) {
        const uint4 a = __ldg(base), b = __ldg(base + 1), d = __ldg(base + Wp), e = __ldg(base + Wp + 1);
        s[0] = blend(bf_lo(a.x), bf_lo(b.x), bf_lo(d.x), bf_lo(e.x), c);
        s[1] = blend(bf_hi(a.x), bf_hi(b.x), bf_hi(d.x), bf_hi(e.x), c);
        s[2] = blend(bf_lo(a.y), bf_lo(b.y), bf_lo(d.y), bf_lo(e.y), c);
        s[3] = blend(bf_hi(a.y), bf_hi(b.y), bf_hi(d.y), bf_hi(e.y), c);
        s[4] = blend(bf_lo(a.z), bf_lo(b.z), bf_lo(d.z), bf_lo(e.z), c);
        s[5] = blend(bf_hi(a.z), bf_hi(b.z), bf_hi(d.z), bf_hi(e.z), c);
        s[6] = blend(bf_lo(a.w), bf_lo(b.w), bf_lo(d.w), bf_lo(e.w), c);
        s[7] = blend(bf_hi(a.w), bf_hi(b.w), bf_hi(d.w), bf_hi(e.w), c);
    } else {
        const float4 *f = reinterpret_cast<const float4 *>(base);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float4 *q = f + k * plane;
            const float4 a = __ldg(q), b = __ldg(q + 1), d = __ldg(q + Wp), e = __ldg(q + Wp + 1);
            s[4 * k + 0] = blend(a.x, b.x, d.x, e.x, c);
            s[4 * k + 1] = blend(a.y, b.y, d.y, e.y, c);
            s[4 * k + 2] = blend(a.z, b.z, d.z, e.z, c);
            s[4 * k + 3] = blend(a.w, b.w, d.w, e.w, c);
        }
    }
}

// bare MUFU.EX2: arguments here are always <= 0, results in [0,1]; 2^-22 relative
__device__ __forceinline__ float ex2_approx(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float max_nan(float a, float b)
{
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));   // torch.max propagates NaN
    return r;
}

// Running fusion state of one channel across views (views arrive in order).
//   sum/mean: acc = ((s0 + s1) + s2) ...   (the reference's order)
//   max     : running max
//   softmax : (m, S = sum e^(s-m), A = sum s*e^(s-m)); result A / S
template <int METHOD>
struct Fuse {
    float a, m, S;
    __device__ __forceinline__ void init() { a = 0.0f; m = -INFINITY; S = 0.0f; }
    // one block of nv views; `first` = no state yet
    template <int VMAX>
    __device__ __forceinline__ void absorb(const float (*s)[kGroupCh], int ch, int nv, bool first)
    {
        if (METHOD == MVHMR_SUM || METHOD == MVHMR_MEAN) {
            float acc = first ? s[0][ch] : __fadd_rn(a, s[0][ch]);
#pragma unroll
            for (int v = 1; v < VMAX; ++v) if (v < nv) acc = __fadd_rn(acc, s[v][ch]);
            a = acc;
        } else if (METHOD == MVHMR_MAX) {
            float mm = first ? s[0][ch] : max_nan(m, s[0][ch]);
#pragma unroll
            for (int v = 1; v < VMAX; ++v) if (v < nv) mm = max_nan(mm, s[v][ch]);
            m = mm;
        } else {
            float mb = s[0][ch];
#pragma unroll
            for (int v = 1; v < VMAX; ++v) if (v < nv) mb = fmaxf(mb, s[v][ch]);
            float SS = 0.0f, AA = 0.0f;
            if (!first) {
                const float mn = fmaxf(m, mb);
                const float sc = ex2_approx(__fsub_rn(m, mn) * kLog2e);
                SS = S * sc; AA = a * sc; mb = mn;
            }
#pragma unroll
            for (int v = 0; v < VMAX; ++v) if (v < nv) {
                const float e = ex2_approx(__fsub_rn(s[v][ch], mb) * kLog2e);
                SS += e;
                AA = fmaf(s[v][ch], e, AA);
            }
            m = mb; S = SS; a = AA;
        }
    }
    __device__ __forceinline__ float result(float Vf) const
    {
        if (METHOD == MVHMR_SUM) return a;
        if (METHOD == MVHMR_MEAN) return __fdiv_rn(a, Vf);
        if (METHOD == MVHMR_MAX) return m;
        return __fdividef(a, S);
    }
};

template <int VMAX, bool BF16, int METHOD, bool MULTI>
__global__ void __launch_bounds__(kBlock)
unproject_kernel(const UnprojParams p)
{
    extern __shared__ float Psm[];                   // V x 12 projection entries of sample b
    const int b = p.b0 + blockIdx.y;
    for (int i = threadIdx.x; i < p.V * 12; i += kBlock) Psm[i] = __ldg(p.proj + (size_t)b * p.V * 12 + i);
    __syncthreads();

    // brick coordinates: z fastest inside the warp, then x, then y
    const int t = threadIdx.x;
    const int lz = t & ((1 << p.ltz) - 1);
    const int lx = (t >> p.ltz) & ((1 << p.ltx) - 1);
    const int ly = t >> (p.ltz + p.ltx);
    int tile = blockIdx.x;
    const int tz = tile % p.tiles_z; tile /= p.tiles_z;
    const int ty = tile % p.tiles_y; tile /= p.tiles_y;
    const int vx = p.x_lo + (tile << p.ltx) + lx;
    const int vy = (ty << p.lty) + ly;
    const int vz = (tz << p.ltz) + lz;
    if (vx >= p.gx || vy >= p.gy || vz >= p.gz) return;
    const long long n = ((long long)vx * p.gy + vy) * p.gz + vz;
    if (n < p.n0 || n >= p.n1) return;

    const long long nl = n - p.n_origin;
    const float *xyz = p.coord + ((size_t)b * p.n_extent + nl) * 3;
    const float X = __ldg(xyz), Y = __ldg(xyz + 1), Z = __ldg(xyz + 2);

    ViewCell cell[VMAX];
    if (!MULTI) {
#pragma unroll
        for (int v = 0; v < VMAX; ++v) if (v < p.V) cell[v] = make_cell(Psm + 12 * v, X, Y, Z, p);
    }

    const int stride_k = BF16 ? 1 : 2;               // texel chunks per 8 channels
    const float Vf = (float)p.V;
    float *outp = p.out + (size_t)b * p.C * p.n_extent + nl;
    const uint4 *fb = p.packed + (size_t)b * p.V * p.chunks * p.plane;

    for (int c0 = 0; c0 < p.C; c0 += kGroupCh) {
        const int k0 = (c0 / kGroupCh) * stride_k;
        Fuse<METHOD> fz[kGroupCh];
        for (int vb = 0; vb < p.V; vb += VMAX) {
            const int nv = min(VMAX, p.V - vb);
            if (MULTI) {
#pragma unroll
                for (int v = 0; v < VMAX; ++v) if (v < nv) cell[v] = make_cell(Psm + 12 * (vb + v), X, Y, Z, p);
            }
            float s[VMAX][kGroupCh];
#pragma unroll
            for (int v = 0; v < VMAX; ++v) if (v < nv) {
                const uint4 *base = fb + ((size_t)(vb + v) * p.chunks + k0) * p.plane + cell[v].off;
                if (!BF16 && k0 + 1 >= p.chunks) {   // C % 8 in (0,4]: second chunk absent
                    float4 const *q = reinterpret_cast<const float4 *>(base);
                    const float4 a = __ldg(q), bb = __ldg(q + 1), d = __ldg(q + p.Wp), e = __ldg(q + p.Wp + 1);
                    s[v][0] = blend(a.x, bb.x, d.x, e.x, cell[v]); s[v][1] = blend(a.y, bb.y, d.y, e.y, cell[v]);
                    s[v][2] = blend(a.z, bb.z, d.z, e.z, cell[v]); s[v][3] = blend(a.w, bb.w, d.w, e.w, cell[v]);
                    s[v][4] = s[v][5] = s[v][6] = s[v][7] = 0.0f;
                } else {
                    sample8<BF16>(s[v], base, p.plane, p.Wp, cell[v]);
                }
            }
#pragma unroll
            for (int ch = 0; ch < kGroupCh; ++ch) fz[ch].template absorb<VMAX>(s, ch, nv, vb == 0);
            if (!MULTI) break;
        }
#pragma unroll
        for (int ch = 0; ch < kGroupCh; ++ch)
            if (c0 + ch < p.C) __stcs(outp + (size_t)(c0 + ch) * p.n_extent, fz[ch].result(Vf));
    }
}

// NCHW -> padded 16-byte-texel planes.  One thread per padded texel.
template <bool BF16>
__global__ void __launch_bounds__(256)
pack_kernel(const void *__restrict__ feats, uint4 *__restrict__ packed, size_t total,
            int C, int H, int W, int chunks, int Hp, int Wp)
{
    constexpr int CPT = BF16 ? 8 : 4;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % Wp) - kBorder;
        size_t r = idx / Wp;
        const int y = (int)(r % Hp) - kBorder; r /= Hp;
        const int k = (int)(r % chunks);
        const size_t bv = r / chunks;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (x >= 0 && x < W && y >= 0 && y < H) {
            const size_t pix = (size_t)y * W + x, hw = (size_t)H * W;
            if (BF16) {
                const unsigned short *src = static_cast<const unsigned short *>(feats) + (bv * C + (size_t)k * CPT) * hw + pix;
                unsigned h[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) h[i] = (k * CPT + i < C) ? (unsigned)__ldg(src + i * hw) : 0u;
                v = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
            } else {
                const float *src = static_cast<const float *>(feats) + (bv * C + (size_t)k * CPT) * hw + pix;
                float f[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) f[i] = (k * CPT + i < C) ? __ldg(src + i * hw) : 0.0f;
                v = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
            }
        }
        packed[idx] = v;
    }
}

static int ilog2_exact(int v) { int l = 0; while ((1 << l) < v) ++l; return (1 << l) == v ? l : -1; }

static int chunks_of(int dtype, int C) { return dtype == MVHMR_BF16 ? (C + 7) / 8 : (C + 3) / 4; }

template <int VMAX, bool BF16, bool MULTI>
static void launch_method(int method, dim3 grid, size_t smem, cudaStream_t st, const UnprojParams &p)
{
    switch (method) {
    case MVHMR_SUM: unproject_kernel<VMAX, BF16, MVHMR_SUM, MULTI><<<grid, kBlock, smem, st>>>(p); break;
    case MVHMR_MEAN: unproject_kernel<VMAX, BF16, MVHMR_MEAN, MULTI><<<grid, kBlock, smem, st>>>(p); break;
    case MVHMR_MAX: unproject_kernel<VMAX, BF16, MVHMR_MAX, MULTI><<<grid, kBlock, smem, st>>>(p); break;
    default: unproject_kernel<VMAX, BF16, MVHMR_SOFTMAX, MULTI><<<grid, kBlock, smem, st>>>(p); break;
    }
}

}  // namespace mvhmr

using namespace mvhmr;

extern "C" size_t mvhmr_packed_bytes(int feat_dtype, int BV, int C, int H, int W)
{
    if ((feat_dtype != MVHMR_F32 && feat_dtype != MVHMR_BF16) || BV < 0 || C < 1 || H < 1 || W < 1) return 0;
    return (size_t)BV * chunks_of(feat_dtype, C) * (H + 2 * kBorder) * (W + 2 * kBorder) * sizeof(uint4);
}

extern "C" size_t mvhmr_unproject_workspace_bytes(int feat_dtype, int feat_layout, int B, int V, int C, int H, int W)
{
    if (feat_layout == MVHMR_LAYOUT_PACKED) return 0;
    if (B < 0 || V < 0) return 0;
    return mvhmr_packed_bytes(feat_dtype, B * V, C, H, W);
}

extern "C" int mvhmr_pack_features(const void *feats, int feat_dtype, void *packed,
                                   int BV, int C, int H, int W, void *stream)
{
    if (feat_dtype != MVHMR_F32 && feat_dtype != MVHMR_BF16)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "pack_features: unknown feat_dtype %d", feat_dtype);
    if (BV < 0 || C < 1 || H < 1 || W < 1)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "pack_features: bad shape BV=%d C=%d H=%d W=%d", BV, C, H, W);
    if (BV == 0) return MVHMR_OK;
    if (!feats || !packed) return fail(MVHMR_ERR_INVALID_ARGUMENT, "pack_features: null pointer");
    if ((uintptr_t)packed & 15) return fail(MVHMR_ERR_INVALID_ARGUMENT, "pack_features: packed buffer must be 16-byte aligned");
    const int chunks = chunks_of(feat_dtype, C), Hp = H + 2 * kBorder, Wp = W + 2 * kBorder;
    const size_t total = (size_t)BV * chunks * Hp * Wp;
    size_t g = (total + 255) / 256;
    const size_t cap = 148u * 32u;
    const unsigned grid = (unsigned)(g > cap ? cap : g);
    if (feat_dtype == MVHMR_BF16)
        pack_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(feats, (uint4 *)packed, total, C, H, W, chunks, Hp, Wp);
    else
        pack_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(feats, (uint4 *)packed, total, C, H, W, chunks, Hp, Wp);
    return check_launch("pack_kernel");
}

extern "C" int mvhmr_unproject_aggregate(const void *feats, int feat_dtype, int feat_layout,
                                         const float *proj, const float *coord, float *out,
                                         int B, int V, int C, int H, int W,
                                         int gx, int gy, int gz, int method,
                                         int b0, int b1, long long n0, long long n1,
                                         long long n_origin, long long n_extent,
                                         unsigned tile_hint, void *ws, size_t ws_bytes, void *stream)
{
    if (method < MVHMR_SUM || method > MVHMR_SOFTMAX)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "Unknown aggregation_method: %d", method);
    if (feat_dtype != MVHMR_F32 && feat_dtype != MVHMR_BF16)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate: unknown feat_dtype %d", feat_dtype);
    if (feat_layout != MVHMR_LAYOUT_NCHW && feat_layout != MVHMR_LAYOUT_PACKED)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate: unknown feat_layout %d", feat_layout);
    if (B < 0 || V < 1 || C < 1 || H < 1 || W < 1 || gx < 1 || gy < 1 || gz < 1)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate: bad shape B=%d V=%d C=%d H=%d W=%d G=(%d,%d,%d)",
                    B, V, C, H, W, gx, gy, gz);
    if (V > 1024) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate: V=%d exceeds 1024", V);
    const long long N = (long long)gx * gy * gz;
    if (b0 < 0 || b1 > B || b0 > b1 || n0 < 0 || n1 > N || n0 > n1)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate: shard window [%d,%d)x[%lld,%lld) outside B=%d N=%lld",
                    b0, b1, n0, n1, B, N);
    if (n_origin < 0 || n_extent < 0 || n0 < n_origin || n1 > n_origin + n_extent)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate: voxels [%lld,%lld) not inside the buffers' range [%lld,%lld)",
                    n0, n1, n_origin, n_origin + n_extent);
    if (b0 == b1 || n0 == n1) return MVHMR_OK;
    if (b1 - b0 > 65535) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate: more than 65535 samples per call");
    if (!feats || !proj || !coord || !out) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate: null pointer");
    if ((long long)H * W > (1LL << 30)) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate: feature map too large");

    // CTA brick: z fastest.  Default: the largest power-of-two z run <= 32 that
    // divides gz (full 128-byte store lines when gz % 32 == 0), the rest of the
    // 256 threads along x, so that one CTA covers an x-z slab at fixed y.
    int TX, TY, TZ;
    if (tile_hint) {
        TX = tile_hint & 0xff; TY = (tile_hint >> 8) & 0xff; TZ = (tile_hint >> 16) & 0xff;
        if (ilog2_exact(TX) < 0 || ilog2_exact(TY) < 0 || ilog2_exact(TZ) < 0 || TX * TY * TZ != kBlock)
            return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate: tile_hint %ux%ux%u must be powers of two with product %d",
                        TX, TY, TZ, kBlock);
    } else {
        TZ = 32;
        if (gz < 32) { TZ = 1; while (TZ < gz) TZ <<= 1; }
        else if (gz % 32 != 0) { if (gz % 16 == 0) TZ = 16; else if (gz % 8 == 0) TZ = 8; }
        const int rest = kBlock / TZ;
        TX = 1; while (TX < rest && TX < gx) TX <<= 1;
        TY = rest / TX;
    }

    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = chunks_of(feat_dtype, C);
    const uint4 *packed;
    if (feat_layout == MVHMR_LAYOUT_NCHW) {
        const size_t need = mvhmr_packed_bytes(feat_dtype, B * V, C, H, W);
        if (!ws || ws_bytes < need)
            return fail(MVHMR_ERR_WORKSPACE, "unproject_aggregate: workspace of %zu bytes required, got %zu", need, ws_bytes);
        if ((uintptr_t)ws & 15) return fail(MVHMR_ERR_WORKSPACE, "unproject_aggregate: workspace must be 16-byte aligned");
        // only the samples of the shard window are packed
        const size_t per_sample_in = (size_t)V * C * H * W * (feat_dtype == MVHMR_BF16 ? 2 : 4);
        const size_t per_sample_pk = need / (size_t)B;
        int rc = mvhmr_pack_features((const char *)feats + per_sample_in * b0, feat_dtype,
                                     (char *)ws + per_sample_pk * b0, (b1 - b0) * V, C, H, W, stream);
        if (rc != MVHMR_OK) return rc;
        packed = (const uint4 *)ws;
    } else {
        if ((uintptr_t)feats & 15) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate: packed features must be 16-byte aligned");
        packed = (const uint4 *)feats;
    }

    UnprojParams p;
    p.packed = packed; p.proj = proj; p.coord = coord; p.out = out;
    p.n0 = n0; p.n1 = n1; p.n_origin = n_origin; p.n_extent = n_extent;
    p.Wp = W + 2 * kBorder; p.plane = (long long)(H + 2 * kBorder) * p.Wp;
    p.V = V; p.C = C; p.W = W; p.H = H; p.chunks = chunks; p.b0 = b0;
    p.gx = gx; p.gy = gy; p.gz = gz;
    p.ltx = ilog2_exact(TX); p.lty = ilog2_exact(TY); p.ltz = ilog2_exact(TZ);
    const long long yz = (long long)gy * gz;
    p.x_lo = (int)(n0 / yz);
    const int x_hi = (int)((n1 - 1) / yz);
    const int tiles_x = (x_hi - p.x_lo + TX) / TX;
    p.tiles_y = (gy + TY - 1) / TY; p.tiles_z = (gz + TZ - 1) / TZ;
    p.Hf = (float)H; p.Wf = (float)W;
    p.sx = (float)(W - 1) / 2.0f; p.sy = (float)(H - 1) / 2.0f;
    const long long tiles = (long long)tiles_x * p.tiles_y * p.tiles_z;
    if (tiles > 0x7fffffffLL) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate: volume too large");
    dim3 grid((unsigned)tiles, (unsigned)(b1 - b0));
    const size_t smem = (size_t)V * 12 * sizeof(float);
    const bool bf = feat_dtype == MVHMR_BF16;
    if (V <= 4) {
        if (bf) launch_method<4, true, false>(method, grid, smem, st, p);
        else launch_method<4, false, false>(method, grid, smem, st, p);
    } else if (V <= 8) {
        if (bf) launch_method<8, true, false>(method, grid, smem, st, p);
        else launch_method<8, false, false>(method, grid, smem, st, p);
    } else {
        if (bf) launch_method<8, true, true>(method, grid, smem, st, p);
        else launch_method<8, false, true>(method, grid, smem, st, p);
    }
    return check_launch("unproject_kernel");
}
