// The L1-gather fused kernel (see unproject.cu for the design notes), templated on the output
// format.  Included by unproject_out0.cu / _out1.cu / _out2.cu, one translation unit per format so
// that they compile in parallel; each defines launch_unproject_gather_out<k>().
#pragma once
#include <cuda_bf16.h>
#include <cstdlib>
#include <cfloat>
#include "mvhmr_common.cuh"
#include "unproject_device.cuh"

namespace mvhmr {

#ifndef MVHMR_WARPS
#define MVHMR_WARPS 16
#endif
#ifndef MVHMR_MINBLOCKS
#define MVHMR_MINBLOCKS 1
#endif
#ifndef MVHMR_CACHE4
#define MVHMR_CACHE4 true
#endif
#ifndef MVHMR_LZCAP
#define MVHMR_LZCAP 32
#endif
// texel gather: read-only path; MVHMR_LDG_EVICT_LAST (tuning build) asks L1 to keep the footprint
__device__ __forceinline__ uint4 ldg_texel(const char *p)
{
#ifdef MVHMR_LDG_EVICT_LAST
    uint4 r;
    asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
#else
    return __ldg(reinterpret_cast<const uint4 *>(p));
#endif
}

constexpr int kWarps = MVHMR_WARPS;       // warps per CTA: consecutive x planes share their texel footprint in L1
constexpr unsigned kNotMine = 0xffffffffu;   // view-0 offset of a voxel outside the shard window (real offsets are multiples of 16)
constexpr int kLzMax = 32;                // voxels of one warp task (z segment): one per lane in phase A

// VMAX : views held in registers at once (V > VMAX walks view blocks, fusion state carried)
// EXACT: V == VMAX — view loops are straight-line code, record layout is a compile-time constant
// CACHE: keep the corner texels of every view across the z walk
// LPB  : log2(pixel bytes) as a compile-time constant (0 = take it from the parameters), so that
//        the second texel of a row is an immediate offset and offsets shift by an immediate
// OUT  : output format — 0 (B,C,N) as the reference, 1 channels-last-3D (B,N,C), 2 fused max_pool3d(2),
//        3 fused 3-D soft-argmax: (B,C,N) as format 0 (or no volume at all, p.out == NULL) plus online-softmax
//        records (max, sum e, sum e*x, sum e*y, sum e*z) of the leading p.sa_J channels per (sample, warp)
//
// Grid: x = blocks of kWarps consecutive x planes, y = voxel row y, z = sample * nseg + z segment.
// One warp = one task = one z segment (<= 32 voxels) of one (sample, x, y) row; the warps of a
// CTA take consecutive x planes, whose projections overlap almost completely in every view, so
// the CTA's texel footprint stays L1-resident.
template <int VMAX, bool EXACT, bool CACHE, bool BF16, int METHOD, int LPB, int OUT>
__global__ void __launch_bounds__(kWarps * 32, MVHMR_MINBLOCKS)
unproject_kernel(const UnprojParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NP = BF16 ? 4 : 2;                 // channel pairs per lane per pass
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *recs = smem_raw + (size_t)warp * p.warp_smem;
    // per-voxel records: V x float4 bilinear weights, then VP x int pixel offsets (-1: voxel not
    // in the shard window).  Records of different lane groups are skewed by 16 bytes so that
    // the groups' broadcast reads fall into different banks.
    float4 *tile = reinterpret_cast<float4 *>(recs + p.off_tile);    // [lz][nvec] output staging

    const int lpb = LPB ? LPB : p.lpb;
    const int nch_pass = LPB ? ((1 << LPB) / 16 < kVecPass ? (1 << LPB) / 16 : kVecPass) : min(p.nchunks, kVecPass);
    const int lpv_log = 31 - __clz(nch_pass);
    const int ngroups = 32 >> lpv_log;               // voxels served per warp step
    const int grp = lane >> lpv_log, chunk = lane & (nch_pass - 1);
    const int nvec = BF16 ? 2 * nch_pass : nch_pass; // float4 vectors per tile row
    const float Vf = (float)p.V;
    // lpb < 0: planes with a padded pixel stride (see make_cell)
    const unsigned px = lpb > 0 ? 1u << lpb : (unsigned)p.pstride, row = (unsigned)p.Wp * px;
    const int wbytes = EXACT ? VMAX * 16 : p.V * 16;
    const int rec_bytes = EXACT ? VMAX * 16 + ((VMAX + 3) & ~3) * 4 : p.rec_bytes;
    // Formats 0 and 3 keep NO separate output tile: row z of the tile is written over the record of voxel z, which
    // is dead by then (its weights and offsets were read earlier in the same step, in program order; lane groups
    // only ever touch the slots of their own run).  The records' pitch is stretched to a tile row where that is
    // longer.  The shared memory this saves (64 KB per CTA at C = 32) is L1 for the texel footprint: with eight
    // views that is worth ~10 %.  Steps beyond the end of a run read a dummy record of zeros
    // instead of record 0, which may already be a tile row.
    // (channel passes — C > 128 fp32 — need the records again: no aliasing there; LPB != 0 implies a single pass)
    // Only the eight-view instantiations do this: with four views the footprint fits the L1 that is left anyway and the
    // stretched records cost 3 % (cfg2 290 -> 299 us).
    const bool ALIAS = (OUT == 0 || OUT == 3) && VMAX > 4 && (LPB != 0 || p.alias);
    // pitch = an ODD number of 16-byte units: the lanes' records in phase A and the lanes' rows in the read-out then
    // spread over all banks without any swizzle
    const int pitch = ALIAS ? (max(rec_bytes, nvec * 16) | 16) : rec_bytes;
    const unsigned char *dummy = ALIAS ? recs + p.off_dummy : recs;
    if (ALIAS) {
        for (int i = lane; i < rec_bytes / 4; i += 32) reinterpret_cast<unsigned *>(recs + p.off_dummy)[i] = 0u;
        __syncwarp();
    }

    // OUT == 3: lane <-> joint (channel) for the whole launch.  The lane keeps ONE online-softmax record
    // over every voxel its warp produces for the current sample; it is flushed when the warp moves on
    // to another sample (tasks are dealt sample-major, so at most once per sample) and at the end.
    float4 *sa_xyz = reinterpret_cast<float4 *>(recs + p.off_xyz);   // [32] coordinates of the task's voxels
    float sa_m = -FLT_MAX;
    u64 sa_SX = pk(0.0f, 0.0f), sa_YZ = pk(0.0f, 0.0f);          // (sum e, sum e*x), (sum e*y, sum e*z)
    int sa_b = -1;
    auto sa_flush = [&]() {
        if (sa_b >= 0 && lane < p.sa_J) {
            float *o = p.sa_rec + (((size_t)sa_b * p.sa_J + lane) * ((size_t)gridDim.x * kWarps) + (size_t)blockIdx.x * kWarps + warp) * 5;
            const f2 sx = upk(sa_SX), yz = upk(sa_YZ);
            o[0] = sa_m; o[1] = sx.x; o[2] = sx.y; o[3] = yz.x; o[4] = yz.y;
        }
        sa_m = -FLT_MAX; sa_SX = sa_YZ = pk(0.0f, 0.0f);
    };

    // Persistent CTAs.  A CTA task = kWarps consecutive x planes of one (sample, z segment, y)
    // row, one plane per warp: their projections overlap almost completely in every view, so the
    // CTA's texel footprint stays L1-resident.  Within a chunk the warps are never synchronised with each
    // other and drift apart in phase (projection / gathers / stores of different warps overlap).
    // Tasks are dealt in chunks of p.ychunk consecutive y rows: all CTAs work on neighbouring chunks (one
    // sample's maps stay in L2) and a CTA's consecutive tasks share texel rows in L1.
    //  * p.deal == NULL (four-view and fused soft-argmax kernels): statically, chunk k to CTA k % gridDim.x; the warps never
    //    meet at all.
    //  * p.deal != NULL (eight-view kernels): the first p.nstatic chunks statically (one round: the pre-assigned
    //    ck = blockIdx.x), every further one DYNAMICALLY from a global counter (zeroed by the launcher).  (Four-view
    //    kernels pay 5-7 % for meeting once per chunk and stay static: p.deal == NULL.)  Thread 0 asks for the CTA's next
    //    chunk while the current one is being worked on and publishes it through shared memory; the CTA meets
    //    once per chunk to read it.  SMs do not all run at the same speed (L2 distance): with a static deal the
    //    slowest of 148 sets the time (+-10 % spread once the larger L1 made these kernels latency-bound).
    const unsigned nchunk = p.nchunk;
    __shared__ unsigned s_deal[2];
    unsigned deal_it = 0;
    for (unsigned ck = blockIdx.x; ck < nchunk; ) {
    // round-robin while the CTA's next chunk is still a static one; from its last static chunk on, the counter
    const bool handover = p.deal && ck + gridDim.x >= p.nstatic;
    if (handover && threadIdx.x == 0) s_deal[deal_it & 1] = p.nstatic + atomicAdd(p.deal, 1u);
    // the last chunks of a dynamic deal are short ones (p.ytail tasks), so that the CTAs finish within a task or two of each other
    unsigned ct = ck < p.nbig ? ck * p.ychunk : p.nbig * p.ychunk + (ck - p.nbig) * p.ytail;
    const unsigned ct_end = min(p.ntasks, ct + (ck < p.nbig ? p.ychunk : p.ytail));
    // task -> (b, z segment, x block, y): divisions once per chunk, then counted up
    unsigned t = ct / (unsigned)p.ty;
    int vy = (int)(ct - t * (unsigned)p.ty);
    unsigned xb = t % p.nxb; t /= p.nxb;
    int seg = (int)(t % (unsigned)p.nseg);
    int b = p.b0 + (int)(t / (unsigned)p.nseg);
    auto next_task = [&]() {
        if (++vy == p.ty) { vy = 0; if (++xb == p.nxb) { xb = 0; if (++seg == p.nseg) { seg = 0; ++b; } } }
    };
    for (; ct < ct_end; ++ct, next_task()) {
    const int xi = (int)xb * kWarps + warp;
    if (xi >= p.tnx) continue;                       // padding of the last x block
    if (OUT == 3 && b != sa_b) { sa_flush(); sa_b = b; }
    const int z0 = seg * p.lz;
    const int zn = min(p.lz, p.gz - z0);             // voxels in this segment (<= 32)
    // pool: every lane group's run starts and ends on an even z, so a 2-voxel pair never straddles groups
    const int steps = OUT == 2 ? ((zn + 2 * ngroups - 1) >> (6 - lpv_log)) << 1 : (zn + ngroups - 1) >> (5 - lpv_log);
    // fused max_pool3d(2): a task is the 2 x 2 rows (x, y) of one pooled row; their maxima meet in the tile
    const int nsub = OUT == 2 ? 4 : 1;
    for (int sub = 0; sub < nsub; ++sub) {
    const int vx = OUT == 2 ? p.x_lo + 2 * xi + (sub >> 1) : p.x_lo + xi;
    const int vyy = OUT == 2 ? 2 * vy + (sub & 1) : vy;
    const long long nrow = ((long long)vx * p.gy + vyy) * p.gz + z0;   // flattened index of the first voxel
    const long long nme = nrow + lane;               // the voxel this lane projects / writes
    const bool mine = (lane < zn) && (nme >= p.n0) && (nme < p.n1);

    // lane / steps by multiplication (exact for lane < 32, steps <= 32): the lane group that serves voxel `lane`
    const unsigned zmagic = zn == p.lz ? p.magic_full : p.magic_last;
    const unsigned g_of_lane = ((unsigned)lane * zmagic) >> 16;
    // row z of the output tile (formats 0 / 3: over the record of voxel z, with its group's skew)
    auto tile_row = [&](int z, unsigned g) -> float4 * {
        return ALIAS ? reinterpret_cast<float4 *>(recs + z * pitch + g * 16) : tile + z * nvec;
    };

    // ---- phase A: one voxel per lane, projected through every view ----
    if (lane < zn) {
        float X = 0.0f, Y = 0.0f, Z = 0.0f;
        if (p.coord) {
            if (mine) {
                const float *xyz = p.coord + ((size_t)b * p.n_extent + (nme - p.n_origin)) * 3;
                X = __ldg(xyz); Y = __ldg(xyz + 1); Z = __ldg(xyz + 2);
            }
        } else {
            // same arithmetic as coord_volume_kernel (bit-identical coordinates, never stored)
            const float *c = p.centers + 3 * b, *R = p.rot + 9 * b;
            const float c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2);
            const float d0 = __fsub_rn(__fadd_rn(p.gpos[0], __fmul_rn(p.gstep[0], (float)vx)), c0);
            const float d1 = __fsub_rn(__fadd_rn(p.gpos[1], __fmul_rn(p.gstep[1], (float)vyy)), c1);
            const float d2 = __fsub_rn(__fadd_rn(p.gpos[2], __fmul_rn(p.gstep[2], (float)(z0 + lane))), c2);
            X = __fadd_rn(rot_row(__ldg(R), __ldg(R + 1), __ldg(R + 2), d0, d1, d2), c0);
            Y = __fadd_rn(rot_row(__ldg(R + 3), __ldg(R + 4), __ldg(R + 5), d0, d1, d2), c1);
            Z = __fadd_rn(rot_row(__ldg(R + 6), __ldg(R + 7), __ldg(R + 8), d0, d1, d2), c2);
        }
        if (OUT == 3) sa_xyz[lane] = make_float4(1.0f, X, Y, Z);   // (1, x, y, z): the sums (S, X) and (Y, Z) advance as two packed FMAs
        unsigned char *rec = recs + lane * pitch + g_of_lane * 16;
        const float4 *Pb = reinterpret_cast<const float4 *>(p.proj + (size_t)b * p.V * 12);
        auto one_view = [&](int v) {
            const float4 P0 = __ldg(Pb + 3 * v), P1 = __ldg(Pb + 3 * v + 1), P2 = __ldg(Pb + 3 * v + 2);
            const ViewCell c = make_cell(P0, P1, P2, X, Y, Z, p, lpb);
            reinterpret_cast<float4 *>(rec)[v] = make_float4(c.w00, c.w01, c.w10, c.w11);
            reinterpret_cast<unsigned *>(rec + wbytes)[v] = (v == 0 && !mine) ? kNotMine : c.off;
        };
        if (EXACT) {                                 // independent chains of all views interleave
#pragma unroll
            for (int v = 0; v < VMAX; ++v) one_view(v);
        } else {
            for (int v = 0; v < p.V; ++v) one_view(v);
        }
    }
    __syncwarp();

    const bool single = EXACT || p.V <= VMAX;        // all views fit one register block
    for (int cb = 0; cb < p.nchunks; cb += kVecPass) {           // channel passes (C > 128 fp32 / 256 bf16)
        const char *lane_base = p.packed + (size_t)b * p.V * p.plane_bytes + ((size_t)(cb + chunk) << 4);
#ifndef MVHMR_TV
#define MVHMR_TV 4
#endif
        constexpr int TV = CACHE ? (VMAX < 4 ? VMAX : 4) : (VMAX < MVHMR_TV ? VMAX : MVHMR_TV);   // views whose texels are in registers at once
        uint4 tex[TV][4];
        unsigned cur[TV];
#pragma unroll
        for (int v = 0; v < TV; ++v) cur[v] = 1u;                  // no texel offset is odd

        // offsets of views [v0, v0+TV) from the voxel record r
        auto load_off = [&](const unsigned char *r, int v0, unsigned *off) {
            if (TV == 4) {
                const uint4 o4 = *reinterpret_cast<const uint4 *>(r + wbytes + v0 * 4);
                off[0] = o4.x; off[1] = o4.y; off[2] = o4.z; off[3] = o4.w;
            } else {
#pragma unroll
                for (int v = 0; v < TV; ++v) off[v] = *reinterpret_cast<const unsigned *>(r + wbytes + (v0 + v) * 4);
            }
        };
        // gathers of up to TV views [v0, v0+nv) at the given offsets; returns false if the voxel is
        // outside the shard window
        auto gather_off = [&](const unsigned *off, int v0, int nv) -> bool {
            const bool inwin = off[0] != kNotMine;                 // only view 0 ever carries the flag
#pragma unroll
            for (int v = 0; v < TV; ++v) {
                if (EXACT || v < nv) {
                    const unsigned o = (v == 0 && !inwin) ? 0u : off[v];
                    if (!CACHE || o != cur[v]) {
                        // 32-bit offset inside the sample's planes: view plane (warp-uniform) + cell
                        const unsigned t = o + (unsigned)(v0 + v) * p.plane32;
                        const char *q0 = lane_base + t;
                        const char *q1 = lane_base + (t + row);
                        tex[v][0] = ldg_texel(q0);
                        tex[v][1] = ldg_texel(q0 + px);
                        tex[v][2] = ldg_texel(q1);
                        tex[v][3] = ldg_texel(q1 + px);
                        cur[v] = o;
                    }
                }
            }
            return inwin;
        };
        auto gather = [&](const unsigned char *r, int v0, int nv) -> bool {
            unsigned off[4];
            load_off(r, v0, off);
            return gather_off(off, v0, nv);
        };
        // row zl of the tile, vector position swizzled by z — or, channels-last-3D output, straight to
        // global memory (a lane group holds all channels of its voxel: one contiguous run), or, pooled
        // output, the running maximum of the 2x2x2 block in row zl / 2
        float *const ovox = p.out + ((size_t)b * p.n_extent + (size_t)(nrow - p.n_origin)) * p.C;   // out_ndhwc: first voxel of the task
        auto emit = [&](int zl, const auto *fz) {
#pragma unroll
            for (int h = 0; h < NP / 2; ++h) {
                const f2 r0 = fz[2 * h].result(Vf), r1 = fz[2 * h + 1].result(Vf);
                const int vec = BF16 ? 2 * chunk + h : chunk;
                float4 val = make_float4(r0.x, r0.y, r1.x, r1.y);
                if (OUT == 1) {
                    const int c = (BF16 ? 8 : 4) * cb + 4 * vec;
                    if (c < p.C) __stcs(reinterpret_cast<float4 *>(ovox + (size_t)zl * p.C + c), val);
                } else if (OUT == 2) {
                    const int row2 = zl >> 1;
                    float4 *slot = tile + row2 * nvec + (vec ^ (row2 & (nvec - 1)));
                    if (sub != 0 || (zl & 1)) {
                        const float4 old = *slot;
                        val = make_float4(max_nan(old.x, val.x), max_nan(old.y, val.y), max_nan(old.z, val.z), max_nan(old.w, val.w));
                    }
                    *slot = val;
                } else {
                    tile_row(zl, grp)[ALIAS ? vec : vec ^ (zl & (nvec - 1))] = val;
                }
            }
        };

        // ---- phase B: each lane group walks its run of consecutive z voxels ----
        const unsigned char *rec = recs + (grp * steps) * pitch + grp * 16;
        int zl = grp * steps;
        const bool piped = single && CACHE && VMAX <= 4;   // the cached path software-pipelines its gathers
        if (piped) {
            // software pipeline: blend this voxel, issue the next voxel's gathers, then do the view
            // fusion (exp, sums) while those loads are in flight.  (Also prefetching the next
            // voxel's weights and the offsets after that costs registers and was slower.)
            bool store_next = gather(zl < zn ? rec : dummy, 0, p.V) && (zl < zn);
            for (int st = 0; st < steps; ++st, ++zl, rec += pitch) {
                const unsigned char *r = zl < zn ? rec : dummy;
                Fuse2<METHOD, VMAX, EXACT> fz[NP];
                const bool store = store_next;
                u64 s[VMAX][NP];
#pragma unroll
                for (int v = 0; v < TV; ++v)
                    if (EXACT || v < p.V)
                        blend_texels<BF16>(s[v], tex[v][0], tex[v][1], tex[v][2], tex[v][3],
                                           reinterpret_cast<const float4 *>(r)[v]);
                if (st + 1 < steps) store_next = gather(zl + 1 < zn ? rec + pitch : dummy, 0, p.V) && (zl + 1 < zn);
#pragma unroll
                for (int i = 0; i < NP; ++i) fz[i].absorb(&s[0][i], NP, p.V, true);
                if (store) emit(zl, fz);
            }
        } else {
            for (int st = 0; st < steps; ++st, ++zl, rec += pitch) {
                const unsigned char *r = zl < zn ? rec : dummy;
                Fuse2<METHOD, VMAX, EXACT> fz[NP];
                bool store = zl < zn;
                for (int vb = 0; vb < p.V; vb += VMAX) {
                    const int nv = EXACT ? VMAX : min(VMAX, p.V - vb);
                    u64 s[VMAX][NP];
#pragma unroll
                    for (int v4 = 0; v4 < VMAX; v4 += TV) {      // TV views at a time
                        if (EXACT || v4 < nv) {
                            const int n4 = EXACT ? TV : min(TV, nv - v4);
                            const bool inwin = gather(r, vb + v4, n4);
                            if (v4 == 0 && vb == 0) store = store && inwin;
#pragma unroll
                            for (int v = 0; v < TV; ++v)
                                if (EXACT || v < n4)
                                    blend_texels<BF16>(s[v4 + v], tex[v][0], tex[v][1], tex[v][2], tex[v][3],
                                                       reinterpret_cast<const float4 *>(r)[vb + v4 + v]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < NP; ++i) fz[i].absorb(&s[0][i], NP, nv, vb == 0);
                }
                if (store) emit(zl, fz);
            }
        }
        __syncwarp();
        // ---- fused 3-D soft-argmax: lane <-> joint, the task's voxels one after the other ----
        // Row z of the tile holds all channels of voxel z: the 32 lanes read 32 different words of one
        // row (conflict-free whatever the swizzle), the voxel's coordinates are one broadcast read.
        // Two straight-line passes: the task's max first (the running sums are rescaled once per task,
        // without a branch), then exp(h - m) and the four sums.  (The fused entry point has no shard
        // window: every voxel of the task counts.)
        if (OUT == 3 && cb == 0) {
            const int c = min(lane, 4 * nvec - 1);
            const int cv = c >> 2;
            // word (z, c) of the tile: row z, vector cv swizzled by z
            auto tile_word = [&](int z, unsigned g) -> float {
                return reinterpret_cast<const float *>(tile_row(z, g))[((ALIAS ? cv : cv ^ (z & (nvec - 1))) << 2) + (c & 3)];
            };
            constexpr int NVC = LPB ? (BF16 ? 2 : 1) * ((1 << LPB) / 16 < kVecPass ? (1 << LPB) / 16 : kVecPass) : 0;   // nvec, when known
            constexpr int SPG = NVC ? (BF16 ? NVC / 2 : NVC) : 1;    // steps per lane group of a full segment = lanes per voxel
            float mt = -FLT_MAX;
            constexpr int PER = NVC < 32 ? NVC : 32;                 // period of the row swizzle of the separate tile
            if (ALIAS && LPB != 0 && zn == kLzMax) {
                // full segment, compile-time pitch: every read carries an immediate offset
                const float *t = reinterpret_cast<const float *>(recs) + c;
#pragma unroll
                for (int z = 0; z < kLzMax; ++z) mt = fmaxf(mt, t[(z * pitch + (z / SPG) * 16) >> 2]);
            } else if (LPB != 0 && zn == kLzMax) {
                // separate tile: rows z and z + PER share their swizzle, so the lane's word offset is computed PER times
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    const float *t = reinterpret_cast<const float *>(tile) + ((cv ^ k) << 2) + (c & 3);
#pragma unroll
                    for (int z = k; z < kLzMax; z += PER) mt = fmaxf(mt, t[z * NVC * 4]);
                }
            } else {
#pragma unroll 4
                for (int z = 0; z < zn; ++z) mt = fmaxf(mt, tile_word(z, ((unsigned)z * zmagic) >> 16));
            }
            const float mn = fmaxf(sa_m, mt);
            const float r = ex2_approx((sa_m - mn) * kLog2e);        // 1 if the max stands, 0 for the first task
            sa_SX = mul2(sa_SX, pk(r, r)); sa_YZ = mul2(sa_YZ, pk(r, r));
            sa_m = mn;
            const float nm = -mn * kLog2e;
            auto absorb_voxel = [&](float h, int z) {
                const float4 c1 = sa_xyz[z];                         // (1, x, y, z)
                const float e = ex2_approx(__fmaf_rn(h, kLog2e, nm));
                const u64 ee = pk(e, e);
                sa_SX = fma2(ee, pk(c1.x, c1.y), sa_SX);
                sa_YZ = fma2(ee, pk(c1.z, c1.w), sa_YZ);
            };
            if (ALIAS && LPB != 0 && zn == kLzMax) {
                const float *t = reinterpret_cast<const float *>(recs) + c;
#pragma unroll
                for (int z = 0; z < kLzMax; ++z) absorb_voxel(t[(z * pitch + (z / SPG) * 16) >> 2], z);
            } else if (LPB != 0 && zn == kLzMax) {
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    const float *t = reinterpret_cast<const float *>(tile) + ((cv ^ k) << 2) + (c & 3);
#pragma unroll
                    for (int z = k; z < kLzMax; z += PER) absorb_voxel(t[z * NVC * 4], z);
                }
            } else {
#pragma unroll 4
                for (int z = 0; z < zn; ++z) absorb_voxel(tile_word(z, ((unsigned)z * zmagic) >> 16), z);
            }
        }
        // ---- read out: lane <-> voxel, one coalesced 128-byte store per channel ----
        if (OUT != 1 && (OUT != 2 || sub == 3) && (OUT != 3 || p.out != nullptr)) {
            const int c_base = (BF16 ? 8 : 4) * cb;
            const bool mine_o = OUT == 2 ? lane < (zn >> 1) : mine;
            const int zr = mine_o ? lane : 0;
            // pool: voxel (x / 2, y / 2, z / 2) of the pooled volume
            const long long vo = OUT == 2 ? ((long long)(p.x_lo / 2 - p.pool_xorg2 + xi) * (p.gy >> 1) + vy) * (p.gz >> 1) + (z0 >> 1) + lane
                                        : nme - p.n_origin;
            const size_t cs = (size_t)p.n_extent_out;
            float *o = p.out + ((size_t)b * p.C + c_base) * cs + (mine_o ? vo : 0);
            const float4 *trow = tile_row(zr, mine_o ? g_of_lane : 0u);
            const int sw = ALIAS ? 0 : zr & (nvec - 1);
            const int kfull = min(nvec, (p.C - c_base) >> 2);        // vectors with all four channels present
            int k = 0;
            for (; k < kfull; ++k, o += 4 * cs) {
                const float4 rr = trow[k ^ sw];
                if (mine_o) { __stcs(o, rr.x); __stcs(o + cs, rr.y); __stcs(o + 2 * cs, rr.z); __stcs(o + 3 * cs, rr.w); }
            }
            if (k < nvec && c_base + 4 * k < p.C) {                  // ragged last vector (C % 4 != 0)
                const float4 rr = trow[k ^ sw];
                const int c = c_base + 4 * k;
                if (mine_o) {
                    __stcs(o, rr.x);
                    if (c + 1 < p.C) __stcs(o + cs, rr.y);
                    if (c + 2 < p.C) __stcs(o + 2 * cs, rr.z);
                }
            }
        }
        __syncwarp();
    }
    }   // sub-rows of a pooled task (the task body)
    }   // tasks of one chunk
    if (handover) {
        __syncthreads();
        ck = s_deal[deal_it & 1];
        ++deal_it;
    } else {
        ck += gridDim.x;                             // static round-robin: the warps do not meet
    }
    }   // persistent chunk loop
    if (OUT == 3) sa_flush();
}

template <int VMAX, bool EXACT, bool CACHE, bool BF16, int LPB, int OUT>
static cudaError_t launch_method(int method, dim3 grid, size_t smem, cudaStream_t st, const UnprojParams &p)
{
#define MVHMR_LAUNCH(M)                                                                                         \
    {                                                                                                           \
        auto kern = unproject_kernel<VMAX, EXACT, CACHE, BF16, M, LPB, OUT>;                                         \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (e != cudaSuccess) return e;                                                                         \
        kern<<<grid, kWarps * 32, smem, st>>>(p);                                                               \
    }
    switch (method) {
    case MVHMR_SUM: MVHMR_LAUNCH(MVHMR_SUM) break;
    case MVHMR_MEAN: MVHMR_LAUNCH(MVHMR_MEAN) break;
    case MVHMR_MAX: MVHMR_LAUNCH(MVHMR_MAX) break;
    default: MVHMR_LAUNCH(MVHMR_SOFTMAX) break;
    }
#undef MVHMR_LAUNCH
    return cudaSuccess;
}

// EXACT view counts get the pixel size as a compile-time constant for the common layouts
// (64 / 128 / 256 bytes per pixel: 32 bf16, 32 fp32 or 64 bf16, 64 fp32 channels).
template <int VMAX, bool EXACT, bool CACHE, bool BF16, int OUT>
static cudaError_t launch_lpb(int method, dim3 grid, size_t smem, cudaStream_t st, const UnprojParams &p)
{
    if (EXACT) {
        if (p.lpb == 6) return launch_method<VMAX, EXACT, CACHE, BF16, EXACT ? 6 : 0, OUT>(method, grid, smem, st, p);
        if (p.lpb == 7) return launch_method<VMAX, EXACT, CACHE, BF16, EXACT ? 7 : 0, OUT>(method, grid, smem, st, p);
        if (p.lpb == 8) return launch_method<VMAX, EXACT, CACHE, BF16, EXACT ? 8 : 0, OUT>(method, grid, smem, st, p);
    }
    return launch_method<VMAX, EXACT, CACHE, BF16, 0, OUT>(method, grid, smem, st, p);
}


// V -> instantiation
template <int OUT>
static int launch_unproject_gather(const UnprojParams &p, bool bf, int method, dim3 grid, size_t smem, cudaStream_t st)
{
    const int V = p.V;
    cudaError_t e;
    // bf16 maps hold 8 channels per lane: the generic view counts (runtime loops) run four views at a
    // time without the texel cache there — the wider instantiations spill (ptxas: 0.4-2.4 KB).
    if (V == 4)
        e = bf ? launch_lpb<4, true, MVHMR_CACHE4, true, OUT>(method, grid, smem, st, p) : launch_lpb<4, true, MVHMR_CACHE4, false, OUT>(method, grid, smem, st, p);
    else if (V == 8)
        e = bf ? launch_lpb<8, true, false, true, OUT>(method, grid, smem, st, p) : launch_lpb<8, true, false, false, OUT>(method, grid, smem, st, p);
    else if (bf)
        e = launch_lpb<4, false, false, true, OUT>(method, grid, smem, st, p);
    else if (V < 4)
        e = launch_lpb<4, false, MVHMR_CACHE4, false, OUT>(method, grid, smem, st, p);
    else
        e = launch_lpb<8, false, false, false, OUT>(method, grid, smem, st, p);
    if (e != cudaSuccess) return fail(MVHMR_ERR_CUDA, "unproject_kernel: %s", cudaGetErrorString(e));
    return check_launch("unproject_kernel");
}

}  // namespace mvhmr
