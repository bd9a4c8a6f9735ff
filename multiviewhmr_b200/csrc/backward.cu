// Backward of unproject + aggregate w.r.t. the feature maps.
//
// out[b,c,n] = fuse_v( s_v ),  s_v = sum_k w_vk * feat[b,v,c,corner_k]   (0 if depth <= 0)
//   d out / d s_v :  sum 1 | mean 1/V | max [v == argmax] | softmax p_v * (1 + s_v - out)
// (the reference's autograd through (x * softmax(x)).sum(0), models/aggregation.py:77-83).
// Views with depth <= 0 were overwritten with zeros in place (:62) and pass no gradient to
// the features, but they still take part in the fusion.  One thread per voxel recomputes the
// cells with the forward's exact arithmetic, re-samples the NCHW maps and scatters with
// red.global.add.f32 into an fp32 gradient buffer (zeroed by the caller).
#include <cstdlib>
#include "mvhmr_common.cuh"
#include "unproject_device.cuh"

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


// ---------------------------------------------------------------------------------------------
// Fast path (mvhmr_unproject_aggregate_backward_ws): the forward's mapping, mirrored.
//   * A warp task = a run of <= 32 consecutive voxels.  Phase A: one voxel per lane, cells of all
//     views (make_cell, the forward's arithmetic) -> records (4 weights, pixel index) in shared
//     memory.  The incoming gradient tile (channels x run) is read with coalesced 128-byte rows and
//     transposed through a swizzled shared tile.
//   * Phase B: a voxel is served by a group of lanes, one 16-byte texel vector each.  For max /
//     softmax the sampled values s_v are rebuilt from the packed feature planes exactly as in the
//     forward (same blend order, so the arg-max agrees with the forward's); V <= VREG views stay
//     in registers, more views are re-sampled in the second pass.
//   * The gradient is scattered into pixel-major fp32 planes with the forward's zero border (the
//     border absorbs corners outside the map: no bounds tests) using red.global.add.v4.f32 — one
//     128-byte line per (voxel, view, corner) for 32 channels, 3.3x the throughput of scalar
//     reds on B200 (scripts/micro/red_bench.cu) — and un-packed to NCHW by a second kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kBwdWarps = 8;

struct BwdParams {
    UnprojParams u;            // packed feature planes (NULL for sum / mean), proj, coord, map geometry
    const float *gout;         // (B, C, N)
    char *gpacked;             // (B*V, Hp, Wp, CG) fp32, zeroed
    long long N;
    long long gplane_bytes;    // Hp * Wp * CG * 4
    int glpb;                  // log2(CG * 4)
    int lz;                    // voxels per warp task
    int warp_smem, rec_bytes, off_tile;
};

__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ACC (sum / mean, fp32 maps, V <= 4): contributions of consecutive voxels of a run that fall into the same cell of
// a view are summed in registers and leave as one set of reds when the cell changes.
#ifndef MVHMR_BWD_MINBLOCKS
#define MVHMR_BWD_MINBLOCKS 3
#endif
// max / softmax re-sample the maps: three CTAs per SM (<= 80 registers, a few spilled words) hide that
// latency better than two with everything in registers (cfg2: max 921 -> 841 us, softmax 1064 -> 1020 us)
template <bool BF16, int METHOD, bool ACC>
__global__ void __launch_bounds__(kBwdWarps * 32, (METHOD == MVHMR_MAX || METHOD == MVHMR_SOFTMAX) ? MVHMR_BWD_MINBLOCKS : 2)
unproject_backward_packed_kernel(const BwdParams q)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const UnprojParams &p = q.u;
    constexpr int NP = BF16 ? 4 : 2;                 // channel pairs per lane
    constexpr int NCH = 2 * NP;                      // channels per lane
    constexpr int VREG = BF16 ? 4 : 8;               // views whose samples stay in registers
    constexpr bool FWD = (METHOD == MVHMR_MAX || METHOD == MVHMR_SOFTMAX);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *recs = smem_raw + (size_t)warp * q.warp_smem;
    float4 *tile = reinterpret_cast<float4 *>(recs + q.off_tile);

    const int nch_pass = min(p.nchunks, kVecPass);
    const int lpv_log = 31 - __clz(nch_pass);
    const int ngroups = 32 >> lpv_log;
    const int grp = lane >> lpv_log, chunk = lane & (nch_pass - 1);
    const int nvec = BF16 ? 2 * nch_pass : nch_pass;
    const int V = p.V, wbytes = V * 16, rec_bytes = q.rec_bytes;
    const int b = blockIdx.y;
    const long long nrow = ((long long)blockIdx.x * kBwdWarps + warp) * q.lz;     // first voxel of this warp's run
    if (nrow >= q.N) return;
    const int zn = (int)min((long long)q.lz, q.N - nrow);
    const int steps = (zn + ngroups - 1) >> (5 - lpv_log);

    // ---- phase A: one voxel per lane ----
    if (lane < zn) {
        const float *xyz = p.coord + ((size_t)b * q.N + nrow + lane) * 3;
        const float X = __ldg(xyz), Y = __ldg(xyz + 1), Z = __ldg(xyz + 2);
        unsigned char *rec = recs + lane * rec_bytes + (lane / steps) * 16;
        const float4 *Pb = reinterpret_cast<const float4 *>(p.proj + (size_t)b * V * 12);
        for (int v = 0; v < V; ++v) {
            const float4 P0 = __ldg(Pb + 3 * v), P1 = __ldg(Pb + 3 * v + 1), P2 = __ldg(Pb + 3 * v + 2);
            const ViewCell c = make_cell(P0, P1, P2, X, Y, Z, p, 0);       // offset in pixels
            reinterpret_cast<float4 *>(rec)[v] = make_float4(c.w00, c.w01, c.w10, c.w11);
            reinterpret_cast<unsigned *>(rec + wbytes)[v] = c.off;
        }
    }
    __syncwarp();

    const float inv_v = 1.0f / (float)V;
    for (int cb = 0; cb < p.nchunks; cb += kVecPass) {
        const int c_base = NCH * cb;
        // ---- incoming gradient: lane <-> voxel, one coalesced row per channel, transposed into the tile ----
        {
            const int zr = lane < zn ? lane : 0;
            const float *g = q.gout + ((size_t)b * p.C + c_base) * q.N + nrow + zr;
            float4 *trow = tile + zr * nvec;
            const int sw = zr & (nvec - 1);
            for (int k = 0; k < nvec; ++k) {
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                const int c = c_base + 4 * k;
                if (lane < zn) {
                    if (c < p.C) r.x = __ldcs(g + (size_t)(4 * k) * q.N);
                    if (c + 1 < p.C) r.y = __ldcs(g + (size_t)(4 * k + 1) * q.N);
                    if (c + 2 < p.C) r.z = __ldcs(g + (size_t)(4 * k + 2) * q.N);
                    if (c + 3 < p.C) r.w = __ldcs(g + (size_t)(4 * k + 3) * q.N);
                    trow[k ^ sw] = r;
                }
            }
        }
        __syncwarp();

        const char *fbase = FWD ? p.packed + (size_t)b * V * p.plane_bytes + ((size_t)(cb + chunk) << 4) : nullptr;
        char *gbase = q.gpacked + (size_t)b * V * q.gplane_bytes + ((size_t)(cb + chunk) * (NCH * 4));
        const unsigned px = 1u << p.lpb, row = (unsigned)p.Wp << p.lpb;
        const unsigned gpx = 1u << q.glpb, grow = (unsigned)p.Wp << q.glpb;

        // sampled values of view v for this lane's channels (the forward's gather + blend)
        auto sample = [&](const unsigned char *r, int v, u64 *s) {
            const unsigned o = reinterpret_cast<const unsigned *>(r + wbytes)[v];
            const char *q0 = fbase + (size_t)v * p.plane_bytes + ((size_t)o << p.lpb);
            const uint4 t00 = __ldg(reinterpret_cast<const uint4 *>(q0));
            const uint4 t01 = __ldg(reinterpret_cast<const uint4 *>(q0 + px));
            const uint4 t10 = __ldg(reinterpret_cast<const uint4 *>(q0 + row));
            const uint4 t11 = __ldg(reinterpret_cast<const uint4 *>(q0 + row + px));
            blend_texels<BF16>(s, t00, t01, t10, t11, reinterpret_cast<const float4 *>(r)[v]);
        };

        constexpr int VA = ACC ? 4 : 1;
        float acc[VA][4][NCH];
        unsigned apix[VA];
#pragma unroll
        for (int v = 0; v < VA; ++v) apix[v] = 0xffffffffu;
        auto scatter = [&](int v, unsigned pix, const float (*val)[NCH]) {       // four corner lines of one cell
            char *g0 = gbase + (size_t)v * q.gplane_bytes + ((size_t)pix << q.glpb);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                char *dst = g0 + ((k & 1) ? gpx : 0u) + ((k & 2) ? grow : 0u);
#pragma unroll
                for (int h = 0; h < NCH / 4; ++h)
                    red_add_v4(reinterpret_cast<float *>(dst) + 4 * h, val[k][4 * h], val[k][4 * h + 1], val[k][4 * h + 2], val[k][4 * h + 3]);
            }
        };

        const unsigned char *rec = recs + (grp * steps) * rec_bytes + grp * 16;
        int zl = grp * steps;
        for (int st = 0; st < steps; ++st, ++zl, rec += rec_bytes) {
            if (zl >= zn) continue;
            float g[NCH];
#pragma unroll
            for (int h = 0; h < NP / 2; ++h) {
                const int vec = BF16 ? 2 * chunk + h : chunk;
                const float4 t = tile[zl * nvec + (vec ^ (zl & (nvec - 1)))];
                g[4 * h] = t.x; g[4 * h + 1] = t.y; g[4 * h + 2] = t.z; g[4 * h + 3] = t.w;
            }
            // pass 1: fusion statistics per channel
            float m[NCH], S[NCH], o[NCH];
            int arg[NCH];
            u64 sreg[VREG][NP];
            const bool keep = V <= VREG;
            if (FWD) {
#pragma unroll
                for (int i = 0; i < NCH; ++i) { m[i] = 0.0f; S[i] = 0.0f; o[i] = 0.0f; arg[i] = 0; }
                if (keep) {
#pragma unroll
                    for (int v = 0; v < VREG; ++v) if (v < V) sample(rec, v, sreg[v]);
                    if (METHOD == MVHMR_SOFTMAX) {           // all views at hand: the max first, then one exp per sample
#pragma unroll
                        for (int i = 0; i < NP; ++i) {
                            f2 mx = upk(sreg[0][i]);
#pragma unroll
                            for (int v = 1; v < VREG; ++v) if (v < V) {
                                const f2 x = upk(sreg[v][i]);
                                mx.x = fmaxf(mx.x, x.x); mx.y = fmaxf(mx.y, x.y);
                            }
                            m[2 * i] = mx.x; m[2 * i + 1] = mx.y;
                        }
                    }
                }
                auto stats = [&](const int v, const u64 *sv) {
#pragma unroll
                    for (int i = 0; i < NP; ++i) {
                        const f2 x = upk(sv[i]);
                        const float xs[2] = {x.x, x.y};
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int c = 2 * i + e;
                            const float sc = xs[e];
                            if (METHOD == MVHMR_MAX) {
                                if (v == 0 || sc > m[c] || (sc != sc && m[c] == m[c])) { m[c] = sc; arg[c] = v; }
                            } else if (keep) {
                                const float ev = ex2_approx((sc - m[c]) * kLog2e);
                                S[c] += ev;
                                o[c] = fmaf(sc, ev, o[c]);             // A = sum s * e
                            } else {                                   // online softmax statistics
                                const float mn = (v == 0) ? sc : fmaxf(m[c], sc);
                                const float scale = (v == 0) ? 0.0f : ex2_approx((m[c] - mn) * kLog2e);
                                const float ev = ex2_approx((sc - mn) * kLog2e);
                                S[c] = fmaf(S[c], scale, ev);
                                o[c] = fmaf(o[c], scale, sc * ev);
                                m[c] = mn;
                            }
                        }
                    }
                };
                if (keep) {
#pragma unroll
                    for (int v = 0; v < VREG; ++v) if (v < V) stats(v, sreg[v]);
                } else {
                    for (int v = 0; v < V; ++v) {
                        u64 sv[NP];
                        sample(rec, v, sv);
                        stats(v, sv);
                    }
                }
                if (METHOD == MVHMR_SOFTMAX) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        S[c] = rcp_approx(S[c]);                       // from here on S holds 1 / sum e
                        o[c] *= S[c];                                  // out = A / S
                    }
                }
            }
            // pass 2: d out / d s_v, scatter
            auto pass2 = [&](const int v, const u64 *kept) {
                const float4 w = reinterpret_cast<const float4 *>(rec)[v];
                if (w.x == 0.0f && w.y == 0.0f && w.z == 0.0f && w.w == 0.0f) return;    // depth <= 0: no gradient (:62)
                float gs[NCH];
                if (!FWD) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) gs[c] = (METHOD == MVHMR_MEAN) ? g[c] * inv_v : g[c];
                } else {
                    u64 sv[NP];
                    if (kept) {
#pragma unroll
                        for (int i = 0; i < NP; ++i) sv[i] = kept[i];
                    } else {
                        sample(rec, v, sv);
                    }
#pragma unroll
                    for (int i = 0; i < NP; ++i) {
                        const f2 x = upk(sv[i]);
                        const float xs[2] = {x.x, x.y};
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int c = 2 * i + e;
                            if (METHOD == MVHMR_MAX) gs[c] = (arg[c] == v) ? g[c] : 0.0f;
                            else {
                                const float pv = ex2_approx((xs[e] - m[c]) * kLog2e) * S[c];
                                gs[c] = g[c] * (pv * (1.0f + xs[e] - o[c]));
                            }
                        }
                    }
                }
                bool any = false;
#pragma unroll
                for (int c = 0; c < NCH; ++c) any = any || (gs[c] != 0.0f);
                if (!any) return;
                const unsigned o_pix = reinterpret_cast<const unsigned *>(rec + wbytes)[v];
                const float wk[4] = {w.x, w.y, w.z, w.w};
                if (ACC) {
                    const int va = ACC ? v : 0;
                    if (o_pix != apix[va]) {
                        if (apix[va] != 0xffffffffu) scatter(v, apix[va], acc[va]);
                        apix[va] = o_pix;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
#pragma unroll
                            for (int c = 0; c < NCH; ++c) acc[va][k][c] = gs[c] * wk[k];
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
#pragma unroll
                            for (int c = 0; c < NCH; ++c) acc[va][k][c] = fmaf(gs[c], wk[k], acc[va][k][c]);
                    }
                } else {
                    float val[4][NCH];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int c = 0; c < NCH; ++c) val[k][c] = gs[c] * wk[k];
                    scatter(v, o_pix, val);
                }
            };
            if (ACC) {
#pragma unroll
                for (int v = 0; v < VA; ++v) if (v < V) pass2(v, nullptr);
            } else if (FWD && keep) {
#pragma unroll
                for (int v = 0; v < VREG; ++v) if (v < V) pass2(v, sreg[v]);
            } else {
                for (int v = 0; v < V; ++v) pass2(v, nullptr);
            }
        }
        if (ACC) {                                   // cells still open at the end of the run
#pragma unroll
            for (int v = 0; v < VA; ++v) if (apix[v] != 0xffffffffu) scatter(v, apix[v], acc[v]);
        }
        __syncwarp();
    }
}

// pixel-major fp32 gradient planes (with border) -> (B*V, C, H, W).  One CTA per map row; the pixel
// vectors are read as 16-byte loads (CG is a multiple of 4), transposed through shared memory and
// written as coalesced rows, one warp per channel.  No integer division per element.
__global__ void __launch_bounds__(256)
unpack_grad_kernel(const float *__restrict__ gpacked, float *__restrict__ gfeat, int C, int H, int W, int CG, int Wp, int Hp)
{
    extern __shared__ float unpack_tile[];           // [min(CG,64)][XB+1]
    constexpr int XB = 128, CB = 64;
    const int bv = blockIdx.x / H, y = blockIdx.x % H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *src_row = gpacked + ((size_t)(bv * Hp + y + kBorder) * Wp + kBorder) * CG;
    for (int x0 = 0; x0 < W; x0 += XB) {
        const int xn = min(XB, W - x0);
        for (int cb = 0; cb < C; cb += CB) {
            const int cn = min(CB, CG - cb);             // a power of two >= 4
            const int qshift = 31 - __clz(cn >> 2);      // log2(16-byte vectors per pixel in this block)
            for (int e = threadIdx.x; e < (xn << qshift); e += blockDim.x) {
                const int j = e >> qshift, q = e & ((cn >> 2) - 1);
                const float4 v = __ldcs(reinterpret_cast<const float4 *>(src_row + (size_t)(x0 + j) * CG + cb) + q);
                float *t = unpack_tile + (4 * q) * (XB + 1) + j;
                t[0] = v.x; t[XB + 1] = v.y; t[2 * (XB + 1)] = v.z; t[3 * (XB + 1)] = v.w;
            }
            __syncthreads();
            const int cw = min(cn, C - cb);
            for (int c = warp; c < cw; c += 8) {
                float *dst = gfeat + ((size_t)(bv * C + cb + c) * H + y) * W + x0;
                for (int j = lane; j < xn; j += 32) dst[j] = unpack_tile[c * (XB + 1) + j];
            }
            __syncthreads();
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

static size_t bwd_gpacked_bytes(int feat_dtype, int B, int V, int C, int H, int W)
{
    const int CG = nchunks_of(feat_dtype, C) * (feat_dtype == MVHMR_BF16 ? 8 : 4);
    return (size_t)B * V * (H + 2 * kBorder) * (W + 2 * kBorder) * CG * sizeof(float);
}

extern "C" size_t mvhmr_unproject_backward_workspace_bytes(int feat_dtype, int B, int V, int C, int H, int W, int method)
{
    if ((feat_dtype != MVHMR_F32 && feat_dtype != MVHMR_BF16) || B < 0 || V < 1 || C < 1 || H < 1 || W < 1) return 0;
    size_t n = bwd_gpacked_bytes(feat_dtype, B, V, C, H, W);
    if (method == MVHMR_MAX || method == MVHMR_SOFTMAX) n += packed_bytes_layout(feat_dtype, B * V, C, H, W, nchunks_of(feat_dtype, C));
    return n;
}

template <bool BF16, bool ACC>
static void launch_bwd(int method, dim3 grid, size_t smem, cudaStream_t st, const BwdParams &q)
{
#define MVHMR_BWD(M)                                                                                          \
    {                                                                                                         \
        auto kern = unproject_backward_packed_kernel<BF16, M, ACC>;                                           \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                   \
        kern<<<grid, kBwdWarps * 32, smem, st>>>(q);                                                          \
    }
    if constexpr (ACC) {                 // the register accumulators serve sum / mean only: no max / softmax instantiation
        if (method == MVHMR_SUM) MVHMR_BWD(MVHMR_SUM) else MVHMR_BWD(MVHMR_MEAN)
    } else {
        switch (method) {
        case MVHMR_SUM: MVHMR_BWD(MVHMR_SUM) break;
        case MVHMR_MEAN: MVHMR_BWD(MVHMR_MEAN) break;
        case MVHMR_MAX: MVHMR_BWD(MVHMR_MAX) break;
        default: MVHMR_BWD(MVHMR_SOFTMAX) break;
        }
    }
#undef MVHMR_BWD
}

extern "C" int mvhmr_unproject_aggregate_backward_ws(const float *grad_out, const void *feats, int feat_dtype,
                                                     const float *proj, const float *coord, float *grad_feats,
                                                     int B, int V, int C, int H, int W, long long N, int method,
                                                     void *ws, size_t ws_bytes, void *stream)
{
    if (method < MVHMR_SUM || method > MVHMR_SOFTMAX)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "Unknown aggregation_method: %d", method);
    if (feat_dtype != MVHMR_F32 && feat_dtype != MVHMR_BF16)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: unknown feat_dtype %d", feat_dtype);
    if (B < 0 || V < 1 || C < 1 || H < 1 || W < 1 || N < 0)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: bad shape B=%d V=%d C=%d H=%d W=%d N=%lld", B, V, C, H, W, N);
    if (V > 1024) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: V=%d exceeds 1024", V);
    if (B == 0) return MVHMR_OK;
    if (B > 65535) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: B=%d exceeds 65535", B);
    if (!grad_out || !feats || !proj || !coord || !grad_feats)
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: null pointer");
    const size_t need = mvhmr_unproject_backward_workspace_bytes(feat_dtype, B, V, C, H, W, method);
    if (!ws || ws_bytes < need)
        return fail(MVHMR_ERR_WORKSPACE, "unproject_aggregate_backward: workspace of %zu bytes required, got %zu", need, ws_bytes);
    if (((uintptr_t)ws & 15) || ((uintptr_t)proj & 15))
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: workspace and proj must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const bool bf = feat_dtype == MVHMR_BF16;
    const bool fwd = method == MVHMR_MAX || method == MVHMR_SOFTMAX;
    const int nchunks = nchunks_of(feat_dtype, C);
    const int Hp = H + 2 * kBorder, Wp = W + 2 * kBorder;
    const int CG = nchunks * (bf ? 8 : 4);
    const size_t gbytes = bwd_gpacked_bytes(feat_dtype, B, V, C, H, W);
    if ((long long)V * Hp * Wp * CG * 4 >= (1LL << 40))
        return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: feature maps too large");

    BwdParams q;
    UnprojParams &p = q.u;
    p = UnprojParams();
    q.gpacked = (char *)ws;
    char *fpacked = (char *)ws + gbytes;
    cudaError_t e = cudaMemsetAsync(q.gpacked, 0, gbytes, st);
    if (e != cudaSuccess) return fail(MVHMR_ERR_CUDA, "unproject_aggregate_backward: memset: %s", cudaGetErrorString(e));
    if (fwd) {
        int rc = pack_features_layout(feats, feat_dtype, fpacked, B * V, C, H, W, nchunks, stream);   // dense pixels
        if (rc != MVHMR_OK) return rc;
    }
    p.packed = fwd ? fpacked : nullptr;
    p.proj = proj; p.coord = coord;
    p.V = V; p.C = C; p.W = W; p.H = H; p.Wp = Wp; p.border = kBorder;
    p.nchunks = nchunks; p.lpb = ilog2_exact(nchunks) + 4; p.pstride = 1 << p.lpb;
    p.plane_bytes = ((long long)Hp * Wp) << p.lpb;
    p.Hf = (float)H; p.Wf = (float)W;
    p.sx = (float)(W - 1) / 2.0f; p.sy = (float)(H - 1) / 2.0f;
    p.rH = 1.0f / (float)H; p.rW = 1.0f / (float)W;
    q.gout = grad_out; q.N = N;
    q.glpb = ilog2_exact(CG) + 2;
    q.gplane_bytes = ((long long)Hp * Wp) << q.glpb;

    if (N > 0) {
        const int nch_pass = nchunks < kVecPass ? nchunks : kVecPass;
        const int nvec = bf ? 2 * nch_pass : nch_pass;
        int lz = 256 / nvec;                                           // gradient tile <= 4 KB per warp
        if (lz < 4) lz = 4;
        if (lz > 32) lz = 32;
        if (const char *env = getenv("MVHMR_BWD_LZ")) { const int v = atoi(env); if (v >= 1 && v <= 32) lz = v; }   // tuning knob
        const int VP = (V + 3) & ~3;
        q.rec_bytes = V * 16 + VP * 4;
        while (lz > 1 && (size_t)lz * q.rec_bytes > 12 * 1024) lz >>= 1;   // voxel records <= 12 KB per warp
        q.lz = lz;
        q.off_tile = (lz * q.rec_bytes + (32 / nch_pass) * 16 + 15) & ~15;
        q.warp_smem = q.off_tile + lz * nvec * 16;
        const size_t smem = (size_t)q.warp_smem * kBwdWarps;
        if (smem > 200 * 1024)
            return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: V=%d needs %zu bytes of shared memory", V, smem);
        const long long tasks = (N + lz - 1) / lz;
        const long long nblk = (tasks + kBwdWarps - 1) / kBwdWarps;
        if (nblk > 0x7fffffffLL) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: too many voxels");
        dim3 grid((unsigned)nblk, B);
        if (bf) launch_bwd<true, false>(method, grid, smem, st, q);
        else if (V <= 4 && !fwd) launch_bwd<false, true>(method, grid, smem, st, q);   // max / softmax: the accumulators cost too many registers
        else launch_bwd<false, false>(method, grid, smem, st, q);
        int rc = check_launch("unproject_backward_packed_kernel");
        if (rc != MVHMR_OK) return rc;
    }
    const long long rows = (long long)B * V * H;
    if (rows > 0x7fffffffLL) return fail(MVHMR_ERR_INVALID_ARGUMENT, "unproject_aggregate_backward: too many rows");
    const int tile_rows = CG < 64 ? CG : 64;
    unpack_grad_kernel<<<(unsigned)rows, 256, (size_t)tile_rows * 129 * 4, st>>>((const float *)q.gpacked, grad_feats, C, H, W, CG, Wp, Hp);
    return check_launch("unpack_grad_kernel");
}
