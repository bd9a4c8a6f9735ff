/*
 * mvhmr_b200.h — C ABI of the B200-native volumetric-aggregation path.
 *
 * This is the drop-in boundary for MultiviewHMR's aggregation hot path
 * (SURVEY.md §8(b)).  The reference has no FFI: its boundary is the Python
 * import of models/aggregation.py, utils/volumetric.py and utils/multiview.py.
 * Every entry point below replaces the torch-op sequence of one reference
 * function (cited per declaration, paths relative to /root/reference) and is
 * what a ctypes/cffi binding on the reference side would bind — see
 * INTEGRATION.md for the stub.
 *
 * Conventions
 *   - Plain pointers and sizes only.  All `const float*` / `void*` data
 *     pointers are DEVICE pointers on the current CUDA device unless a
 *     parameter says "host".  The caller owns every buffer; the library
 *     allocates nothing and keeps no state between calls (two bounded
 *     scratch tables aside: texture-object descriptors of the caller's
 *     workspace, and a 4 KB ring of per-launch work counters in the
 *     library's own device data, zeroed on the caller's stream before use).
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as
 *     void*; NULL = legacy default stream), does no host synchronisation and
 *     no allocation, and is CUDA-graph capturable.
 *   - Return value: MVHMR_OK or a negative MVHMR_ERR_*.  The message of the
 *     last failure on the calling thread is available from
 *     mvhmr_last_error().  The library never aborts and never prints.
 *   - Re-entrant: safe to call concurrently from several host threads.
 *   - fp32 results follow the reference's rounding order op for op
 *     (sum / mean / max fusion and all grids are bit-identical to the
 *     reference's CPU torch path; softmax differs only by the exp
 *     approximation, ~1e-7 relative).
 */
#ifndef MVHMR_B200_H
#define MVHMR_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVHMR_ABI_VERSION 1

/* return codes */
#define MVHMR_OK 0
#define MVHMR_ERR_INVALID_ARGUMENT (-1) /* maps to ValueError on the Python side */
#define MVHMR_ERR_WORKSPACE (-2)        /* workspace missing or too small        */
#define MVHMR_ERR_CUDA (-3)             /* a CUDA runtime call failed            */

/* aggregation_method of models/aggregation.py:71-85 */
#define MVHMR_SUM 0
#define MVHMR_MEAN 1
#define MVHMR_MAX 2
#define MVHMR_SOFTMAX 3

/* storage type of the feature maps */
#define MVHMR_F32 0
#define MVHMR_BF16 1

/* layout of the feature maps handed to mvhmr_unproject_aggregate */
#define MVHMR_LAYOUT_NCHW 0   /* (B,V,C,H,W) as the reference passes them      */
#define MVHMR_LAYOUT_PACKED 1 /* already in the library's gather layout, i.e.  */
                              /* the output of mvhmr_pack_features             */
#define MVHMR_LAYOUT_NHWC 2   /* (B,V,H,W,C) channels-last, gathered in place  */
                              /* (no pack pass, no workspace): what a cuDNN    */
                              /* 1x1 conv emits for a channels_last tensor.    */
                              /* Needs C*elemsize = 16*2^k bytes, H,W >= 2,    */
                              /* 16-byte alignment; texels must be finite (a   */
                              /* corner outside the map is realised as         */
                              /* weight*0 times a real texel).                 */

int mvhmr_abi_version(void);

/* Message of the last error on this thread ("" if none).  Never NULL. */
const char *mvhmr_last_error(void);

/* ---- grids -------------------------------------------------------------- */

/* Per-sample cuboid coordinate volume.
 * Replaces models/aggregation.py:135-161 (grid build), :184-187 (centre,
 * rotate, un-centre) and utils/volumetric.py:102-114 (rotate_coord_volume):
 *   out[b,x,y,z,:] = rot[b] · ((pos + step*idx) - centers[b]) + centers[b]
 * with the reference's fp32 roundings (separate mul and add; K=3 FMA chain).
 * out (B,Gx,Gy,Gz,3); centers (B,3); rot (B,3,3) row-major, identity in eval;
 * pos_host / step_host: 3 floats each in HOST memory (pos = -side/2,
 * step = side/(G-1), already rounded to fp32). */
int mvhmr_build_coord_volumes(float *out, const float *centers, const float *rot,
                              const float *pos_host, const float *step_host,
                              int B, int Gx, int Gy, int Gz, void *stream);

/* utils/volumetric.py:102-114 on an arbitrary (N,3) point array:
 * out[n,:] = rot · pts[n,:].  rot_host: 9 floats, row-major, HOST memory.
 * out may alias pts. */
int mvhmr_rotate_points(float *out, const float *pts, const float *rot_host, size_t N, void *stream);

/* utils/multiview.py:89-110 for torch inputs: out = [pts 1] · P^T, fp32 K=4 FMA
 * chain.  P (3,4) device.  euclid=0: out (N,3) homogeneous (x*w, y*w, w);
 * euclid=1: out (N,2) after utils/multiview.py:72-86 (division by w; w == 0
 * divides by zero exactly as the reference does). */
int mvhmr_project_points(float *out, const float *P, const float *pts, size_t N, int euclid, void *stream);

/* ---- unproject + aggregate ---------------------------------------------- */

/* Bytes of the packed gather layout for BV = B*V feature maps (0 on bad args). */
size_t mvhmr_packed_bytes(int feat_dtype, int BV, int C, int H, int W);

/* NCHW (BV,C,H,W) -> packed gather layout: pixel-major planes (all channels of
 * a pixel contiguous, padded to a power-of-two number of 16-byte vectors),
 * each plane surrounded by a 2-texel zero border so that out-of-map corners
 * read zeros (grid_sample padding_mode='zeros'). */
int mvhmr_pack_features(const void *feats, int feat_dtype, void *packed,
                        int BV, int C, int H, int W, void *stream);

/* Workspace needed by mvhmr_unproject_aggregate for the given layout
 * (= mvhmr_packed_bytes for NCHW input, 0 for PACKED input). */
size_t mvhmr_unproject_workspace_bytes(int feat_dtype, int feat_layout, int B, int V, int C, int H, int W);

/* Fused replacement of models/aggregation.py:20-87 `unprojection`:
 * project every voxel centre with each view's 3x4 matrix, bilinear-sample the
 * view's feature map (align_corners=True, zeros padding, x normalised by H and
 * y by W as the reference does), zero samples with depth <= 0, fuse over views.
 *   feats  (B,V,C,H,W) [NCHW] or packed; feat_dtype fp32 / bf16
 *   proj   (B,V,3,4) fp32
 *   coord  (B,n_extent,3) fp32 and out (B,C,n_extent) fp32 hold voxels
 *          [n_origin, n_origin+n_extent) of the flattened N = gx*gy*gz voxel
 *          volume (pass 0,N for whole-volume buffers; a slab shard may pass
 *          compact buffers covering only its slab)
 *   method MVHMR_SUM / MEAN / MAX / SOFTMAX
 *   [b0,b1) x [n0,n1): shard window — only these samples / voxels are computed
 *   and written (pass 0,B,0,N for everything).  Slab and batch shards of one
 *   problem are bit-identical to the unsharded call.
 *   tile_hint: 0 = automatic; otherwise the number of consecutive z voxels
 *   (<= 32) one warp walks per task (tuning knob; results do not depend on it).
 *   ws / ws_bytes: caller workspace (see mvhmr_unproject_workspace_bytes).
 * gx,gy,gz: volume shape; used only to cut the volume into z runs, any
 * factorisation with gx*gy*gz == N is legal.
 * Limits: the padded maps of one sample stay below 4 GiB; V up to about 650
 * (the per-voxel records of all views must fit the shared memory of a warp
 * task: larger V returns MVHMR_ERR_INVALID_ARGUMENT); any C, H, W >= 1. */
int mvhmr_unproject_aggregate(const void *feats, int feat_dtype, int feat_layout,
                              const float *proj, const float *coord, float *out,
                              int B, int V, int C, int H, int W,
                              int gx, int gy, int gz, int method,
                              int b0, int b1, long long n0, long long n1,
                              long long n_origin, long long n_extent,
                              unsigned tile_hint, void *ws, size_t ws_bytes, void *stream);

/* The cuboid grid of models/aggregation.py:135-187 as a descriptor instead of a
 * materialised (B,Gx,Gy,Gz,3) tensor: voxel (x,y,z) of sample b sits at
 *   rot[b] * ((pos + step*(x,y,z)) - centers[b]) + centers[b]
 * exactly as mvhmr_build_coord_volumes would write it (same fp32 roundings). */
typedef struct mvhmr_grid {
    const float *centers; /* device, (B,3)                                    */
    const float *rot;     /* device, (B,3,3) row-major, identity in eval mode */
    float pos[3];         /* -side/2 per axis, rounded to fp32                */
    float step[3];        /* side/(G-1) per axis, rounded to fp32             */
} mvhmr_grid_t;

/* mvhmr_unproject_aggregate with the coordinates generated in registers from
 * `grid` (host pointer to the descriptor, read during the call): saves the
 * 12 B/voxel write and read of the coord volume.
 * Requires the true volume shape in gx,gy,gz.  Results are bit-identical to
 * building the volume first. */
int mvhmr_unproject_aggregate_grid(const void *feats, int feat_dtype, int feat_layout,
                                   const float *proj, const mvhmr_grid_t *grid, float *out,
                                   int B, int V, int C, int H, int W,
                                   int gx, int gy, int gz, int method,
                                   int b0, int b1, long long n0, long long n1,
                                   long long n_origin, long long n_extent,
                                   unsigned tile_hint, void *ws, size_t ws_bytes, void *stream);

/* Consumer-side output formats (the first thing the reference's
 * consumer does with the aggregate is a 3-D conv block followed by max_pool3d(2),
 * models/regressor.py:70-75):
 *   MVHMR_OUT_NDHWC  out is (B, n_extent, C): a voxel's channels are contiguous — PyTorch's
 *                    channels_last_3d memory format, what a tensor-core 3-D conv wants.  Written
 *                    straight from the fusion registers (no shared-memory transposition);
 *                    needs C % 4 == 0.  Same values as the default layout.
 *   MVHMR_OUT_POOL2  out is (B, C, n_extent / 8): only the maximum of every 2x2x2 voxel block is
 *                    written (max_pool3d(kernel 2, stride 2) of the aggregate, NaN-propagating
 *                    like torch), the full-resolution volume never reaches memory.  Needs an even
 *                    volume shape, windows made of whole x-plane pairs and C <= 128 (fp32).
 * Exactly one of `coord` / `grid` is non-NULL.  All other arguments as mvhmr_unproject_aggregate. */
#define MVHMR_OUT_NDHWC 1u
#define MVHMR_OUT_POOL2 2u
int mvhmr_unproject_aggregate_fmt(const void *feats, int feat_dtype, int feat_layout,
                                  const float *proj, const float *coord, const mvhmr_grid_t *grid,
                                  float *out, unsigned out_flags,
                                  int B, int V, int C, int H, int W,
                                  int gx, int gy, int gz, int method,
                                  int b0, int b1, long long n0, long long n1,
                                  long long n_origin, long long n_extent,
                                  unsigned tile_hint, void *ws, size_t ws_bytes, void *stream);

/* Reduced-precision fast path for callers inside BASELINE.json's bf16 tolerance (1e-2 relative):
 * the bilinear samples are taken by the texture units from fp16 copies of the maps (hardware
 * interpolation weights are 1.8 fixed point).  Measured deviation from the reference on N(0,1)
 * maps: about 4e-3 relative — never use it where the fp32 contract (1e-5) applies.  The error
 * is ABSOLUTE in nature (about |texel| / 256 per corner): a volume that only grazes the edge of the maps
 * (a fraction of a percent of non-zero voxels with values of a few hundredths) can show several
 * percent relative deviation at the same 3e-3 absolute error.
 *   feats  (B,V,C,H,W) NCHW, fp32 or bf16 (|x| > 65504 saturates in fp16)
 *   exactly one of coord / grid non-NULL; out (B,C,n_extent) fp32 as in mvhmr_unproject_aggregate
 *   V <= 8 and V*ceil(C/4)*(H+1) <= 8192, else MVHMR_ERR_INVALID_ARGUMENT (use the exact path)
 *   ws: mvhmr_unproject_tex_workspace_bytes(B,V,C,H,W) bytes (the fp16 planes)
 * The library keeps a small cache of texture-object descriptors (no device memory). */
size_t mvhmr_unproject_tex_workspace_bytes(int B, int V, int C, int H, int W);
int mvhmr_unproject_aggregate_tex(const void *feats, int feat_dtype,
                                  const float *proj, const float *coord, const mvhmr_grid_t *grid, float *out,
                                  int B, int V, int C, int H, int W,
                                  int gx, int gy, int gz, int method,
                                  int b0, int b1, long long n0, long long n1,
                                  long long n_origin, long long n_extent,
                                  void *ws, size_t ws_bytes, void *stream);

/* mvhmr_soft_argmax3d_strided with the voxel coordinates generated from `grid`
 * (same arithmetic as mvhmr_build_coord_volumes, so the result has the same bits
 * as building the coord volume first): with mvhmr_unproject_aggregate_grid no
 * coordinate volume ever exists in device memory. */
int mvhmr_soft_argmax3d_grid(const float *vol, const mvhmr_grid_t *grid, float *out,
                             int B, int J, int gx, int gy, int gz, long long sample_stride,
                             void *ws, size_t ws_bytes, void *stream);

/* Backward of mvhmr_unproject_aggregate w.r.t. the feature maps (the gradient
 * torch autograd produces for models/aggregation.py:20-87; training goes through
 * this op, train.py:110).  grad_out (B,C,N) fp32; feats (B,V,C,H,W) NCHW fp32 or
 * bf16 (re-sampled to rebuild the fusion Jacobian); grad_feats (B,V,C,H,W) fp32,
 * ZEROED BY THE CALLER, accumulated with atomics (summation order, and therefore
 * the last bits, vary from run to run).  V <= 64. */
int mvhmr_unproject_aggregate_backward(const float *grad_out, const void *feats, int feat_dtype,
                                       const float *proj, const float *coord, float *grad_feats,
                                       int B, int V, int C, int H, int W, long long N, int method, void *stream);

/* The same gradient, fast path: the forward's lane-group mapping, scatter with
 * red.global.add.v4.f32 into pixel-major planes inside the workspace, then one
 * un-pack pass.  grad_feats is OVERWRITTEN (no zeroing by the caller).  ws of
 * mvhmr_unproject_backward_workspace_bytes(...) bytes, 16-byte aligned. */
size_t mvhmr_unproject_backward_workspace_bytes(int feat_dtype, int B, int V, int C, int H, int W, int method);
int mvhmr_unproject_aggregate_backward_ws(const float *grad_out, const void *feats, int feat_dtype,
                                          const float *proj, const float *coord, float *grad_feats,
                                          int B, int V, int C, int H, int W, long long N, int method,
                                          void *ws, size_t ws_bytes, void *stream);

/* Self-test: runs the kernel's two exact-division shortcuts (division by a launch
 * constant via reciprocal + FMA correction; shared-reciprocal division) for ALL
 * 2^32 numerators against div.rn.f32 with divisor d and ADDS the number of
 * bitwise mismatches to *mismatches (device pointer, zeroed by the caller). */
int mvhmr_selftest_division(float d, unsigned long long *mismatches, void *stream);

/* ---- 3-D soft-argmax ------------------------------------------------------ */
/* Not in the reference; upstream definition
 * (Learnable-Triangulation integrate_tensor_3d_with_coordinates, cited by URL at
 * models/aggregation.py:13-17):
 *   p = softmax(vol[b,j,:]);  out[b,j,:] = sum_n p[n] * coord[b,n,:]        */

/* Number of (max, sum_e, sum_e*x, sum_e*y, sum_e*z) partial records per (b,j)
 * that mvhmr_soft_argmax3d_partials writes for N voxels. */
int mvhmr_soft_argmax3d_num_slices(long long N);

size_t mvhmr_soft_argmax3d_workspace_bytes(int B, int J, long long N);

/* One-call form: vol (B,J,N), coord (B,N,3) -> out (B,J,3). */
int mvhmr_soft_argmax3d(const float *vol, const float *coord, float *out,
                        int B, int J, long long N, void *ws, size_t ws_bytes, void *stream);

/* The same over the first J channels of a wider volume: sample b starts at
 * vol + b * sample_stride floats (sample_stride >= J*N; e.g. C*N for the leading
 * J joints of a (B,C,G,G,G) aggregate), so no contiguous copy is needed. */
int mvhmr_soft_argmax3d_strided(const float *vol, const float *coord, float *out,
                                int B, int J, long long N, long long sample_stride,
                                void *ws, size_t ws_bytes, void *stream);

/* Shard form.  partials: (B,J,S,5) with S = num_slices(n1-n0), over voxels
 * [n0,n1) of each sample; vol / coord are indexed with the FULL N. */
int mvhmr_soft_argmax3d_partials(const float *vol, const float *coord, float *partials,
                                 int B, int J, long long N, long long n0, long long n1, void *stream);

/* Merge S records per (b,j) (from one or several shards, concatenated along S)
 * into out (B,J,3). */
int mvhmr_soft_argmax3d_finalize(const float *partials, float *out, int B, int J, int S, void *stream);

/* ---- fused unproject + aggregate + 3-D soft-argmax ---------------------------- */
/* BASELINE.json's target path in ONE kernel: mvhmr_unproject_aggregate (models/aggregation.py:20-87)
 * whose warps, before a task's tile leaves shared memory, also fold its voxels into online-softmax
 * records of the leading J channels (lane <-> joint; formula: see "3-D soft-argmax" above), merged by
 * mvhmr_soft_argmax3d_finalize on the same stream.  The aggregate is never re-read from memory.
 *   out     (B,C,N) fp32 as in mvhmr_unproject_aggregate (same bits) — or NULL: heads that only need
 *           the joints skip the B*C*N*4-byte write altogether
 *   joints  (B,J,3) fp32, 1 <= J <= min(32, C)
 *   exactly one of coord / grid non-NULL (the coordinates the softmax expectation is taken over are
 *   the voxel centres the kernel projects)
 *   ws      as for mvhmr_unproject_aggregate;  sa_ws: mvhmr_unproject_softargmax_workspace_bytes(B,J)
 * joints agree with mvhmr_soft_argmax3d over the stored volume to fp32 summation-order noise
 * (<= 1e-5 * max|coord|, the contract of the two-kernel path). */
size_t mvhmr_unproject_softargmax_workspace_bytes(int B, int J);
int mvhmr_unproject_aggregate_softargmax(const void *feats, int feat_dtype, int feat_layout,
                                         const float *proj, const float *coord, const mvhmr_grid_t *grid,
                                         float *out, float *joints, int J,
                                         int B, int V, int C, int H, int W,
                                         int gx, int gy, int gz, int method,
                                         unsigned tile_hint, void *ws, size_t ws_bytes,
                                         void *sa_ws, size_t sa_ws_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MVHMR_B200_H */
