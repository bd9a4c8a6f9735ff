/*
 * mvhmr_oracle.c — CPU restatement of MultiviewHMR's volumetric-aggregation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under multiviewhmr_b200/ may include, link
 * or call this file; it exists so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py can check the CUDA path against an independent
 * plain-C statement of what the reference computes.
 *
 * Parity status: PINNED for rows a2-a12 (coord volume, projection, sampling,
 * fusion) against tests/golden/ *.npz, which were produced by importing the
 * reference itself (tests/golden/make_golden.py).  UNPINNED for the 3-D
 * soft-argmax: the reference has no such function, the
 * formula restated here is upstream Learnable-Triangulation's
 * integrate_tensor_3d_with_coordinates, cited by URL at
 * /root/reference/models/aggregation.py:13-17.
 *
 * The reference's arithmetic lives in PyTorch ATen (torch 2.11, CPU build here):
 *   - `@`/`mm` with K=4 / K=3  == forward FMA chain starting from x*P[r][0]
 *     (0 mismatches in 600k values, verified in this container);
 *   - tensor / python-int on CPU == IEEE division (ATen div_true_kernel);
 *   - F.grid_sample bilinear/zeros/align_corners=True == ATen
 *     GridSamplerKernel.cpp ApplyGridSample<2, Bilinear, Zeros, true>.
 * Build with -ffp-contract=off: every rounding below is intentional.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_SUM 0
#define ORC_MEAN 1
#define ORC_MAX 2
#define ORC_SOFTMAX 3
#define ORC_MAX_VIEWS 1024

int orc_abi_version(void) { return 1; }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------- */
/* a2 + a3: coord volume.                                                    */
/* /root/reference/models/aggregation.py:140-161  grid = pos + (side/(G-1))*idx
 *   (python float scalars are cast to fp32; mul and add round separately)
 * :184  coord - center
 * /root/reference/utils/volumetric.py:102-114  rot.mm(coord.t()).t(), fp32 sgemm
 *   K=3 == forward FMA chain; rot is float64 numpy cast to fp32 by the caller
 * :186  + center                                                             */
void orc_build_coord_volumes(float *out, const float *centers, const float *rot,
                             const float *pos, const float *step, int B, int Gx, int Gy, int Gz)
{
    for (int b = 0; b < B; ++b) {
        const float *c = centers + 3 * b;
        const float *R = rot + 9 * b;
        float *ob = out + (size_t)b * Gx * Gy * Gz * 3;
#pragma omp parallel for schedule(static)
        for (int ix = 0; ix < Gx; ++ix)
            for (int iy = 0; iy < Gy; ++iy)
                for (int iz = 0; iz < Gz; ++iz) {
                    float g[3], d[3];
                    float m0 = step[0] * (float)ix; g[0] = pos[0] + m0;
                    float m1 = step[1] * (float)iy; g[1] = pos[1] + m1;
                    float m2 = step[2] * (float)iz; g[2] = pos[2] + m2;
                    for (int k = 0; k < 3; ++k) d[k] = g[k] - c[k];
                    float *o = ob + (((size_t)ix * Gy + iy) * Gz + iz) * 3;
                    for (int i = 0; i < 3; ++i) {
                        float acc = R[3 * i + 0] * d[0];
                        acc = fmaf(R[3 * i + 1], d[1], acc);
                        acc = fmaf(R[3 * i + 2], d[2], acc);
                        o[i] = acc + c[i];
                    }
                }
    }
}

/* /root/reference/utils/volumetric.py:102-114 rotate_coord_volume on (N,3) points */
void orc_rotate_points(float *out, const float *pts, const float *R, size_t N)
{
#pragma omp parallel for schedule(static)
    for (size_t n = 0; n < N; ++n) {
        const float *p = pts + 3 * n;
        float r[3];
        for (int i = 0; i < 3; ++i) {
            float acc = R[3 * i + 0] * p[0];
            acc = fmaf(R[3 * i + 1], p[1], acc);
            acc = fmaf(R[3 * i + 2], p[2], acc);
            r[i] = acc;
        }
        out[3 * n + 0] = r[0]; out[3 * n + 1] = r[1]; out[3 * n + 2] = r[2];
    }
}

/* a7: /root/reference/utils/multiview.py:55-69,89-110
 *   [X Y Z 1] @ P^T, fp32 sgemm K=4 == forward FMA chain.
 * euclid != 0 additionally applies homogeneous_to_euclidean (:72-86): (N,2). */
static inline void orc_project1(const float *P, float X, float Y, float Z, float *o3)
{
    for (int r = 0; r < 3; ++r) {
        float acc = X * P[4 * r + 0];
        acc = fmaf(Y, P[4 * r + 1], acc);
        acc = fmaf(Z, P[4 * r + 2], acc);
        acc = fmaf(1.0f, P[4 * r + 3], acc);
        o3[r] = acc;
    }
}

void orc_project_points(float *out, const float *P, const float *pts, size_t N, int euclid)
{
#pragma omp parallel for schedule(static)
    for (size_t n = 0; n < N; ++n) {
        float h[3];
        orc_project1(P, pts[3 * n], pts[3 * n + 1], pts[3 * n + 2], h);
        if (euclid) {
            out[2 * n + 0] = h[0] / h[2];
            out[2 * n + 1] = h[1] / h[2];
        } else {
            out[3 * n + 0] = h[0]; out[3 * n + 1] = h[1]; out[3 * n + 2] = h[2];
        }
    }
}

/* a8-a10 position part: pixel-space sampling position of one voxel in one view.
 * /root/reference/models/aggregation.py:42-51 and ATen grid_sampler unnormalize
 * (align_corners=True): ix = (gx + 1) * ((W-1)/2).
 * NOTE the reference normalises x by feature_shape[0] (=H) and y by
 * feature_shape[1] (=W) — kept.                            */
static inline int orc_sample_pos(const float *P, float X, float Y, float Z, int H, int W,
                                 float *ix, float *iy)
{
    float h[3];
    orc_project1(P, X, Y, Z, h);
    int invalid = h[2] <= 0.0f;
    float w = (h[2] == 0.0f) ? 1.0f : h[2];
    float x = h[0] / w;
    float y = h[1] / w;
    float qx = x / (float)H;
    float qy = y / (float)W;
    float gx = 2.0f * (qx - 0.5f);
    float gy = 2.0f * (qy - 0.5f);
    *ix = (gx + 1.0f) * ((float)(W - 1) / 2.0f);
    *iy = (gy + 1.0f) * ((float)(H - 1) / 2.0f);
    return invalid;
}

/* Debug/pinning export: sampling positions + invalid mask for one (b,v). */
void orc_sample_positions(float *ix, float *iy, uint8_t *invalid, const float *P,
                          const float *coord, size_t N, int H, int W)
{
#pragma omp parallel for schedule(static)
    for (size_t n = 0; n < N; ++n)
        invalid[n] = (uint8_t)orc_sample_pos(P, coord[3 * n], coord[3 * n + 1], coord[3 * n + 2],
                                             H, W, ix + n, iy + n);
}

/* Bilinear cell of ATen's CPU kernel: x_w = floor(x); w = x - x_w; e = 1 - w;
 * n = y - y_n; s = 1 - n; nw = s*e, ne = s*w, sw = n*e, se = n*w; a corner
 * outside the map contributes 0 * weight (NaN weights therefore propagate,
 * as they do in the reference for non-finite positions).                    */
typedef struct {
    float wgt[4];      /* nw ne sw se */
    ptrdiff_t off[4];  /* element offsets inside one (H,W) plane, -1 = outside */
} orc_cell_f32;

static inline void orc_cell32(float ix, float iy, int H, int W, orc_cell_f32 *c)
{
    float xw = floorf(ix), yn = floorf(iy);
    float w = ix - xw, e = 1.0f - w, n = iy - yn, s = 1.0f - n;
    c->wgt[0] = s * e; c->wgt[1] = s * w; c->wgt[2] = n * e; c->wgt[3] = n * w;
    /* bounds test in the float domain: no UB on huge / non-finite positions */
    int x0ok = (xw >= 0.0f) && (xw <= (float)(W - 1));
    int x1ok = (xw >= -1.0f) && (xw <= (float)(W - 2));
    int y0ok = (yn >= 0.0f) && (yn <= (float)(H - 1));
    int y1ok = (yn >= -1.0f) && (yn <= (float)(H - 2));
    ptrdiff_t x0 = x0ok || x1ok ? (ptrdiff_t)xw : 0, y0 = y0ok || y1ok ? (ptrdiff_t)yn : 0;
    c->off[0] = (x0ok && y0ok) ? y0 * W + x0 : -1;
    c->off[1] = (x1ok && y0ok) ? y0 * W + x0 + 1 : -1;
    c->off[2] = (x0ok && y1ok) ? (y0 + 1) * W + x0 : -1;
    c->off[3] = (x1ok && y1ok) ? (y0 + 1) * W + x0 + 1 : -1;
}

static inline float orc_fuse32(const float *s, int V, int method)
{
    if (method == ORC_SUM || method == ORC_MEAN) {
        float acc = s[0];
        for (int v = 1; v < V; ++v) acc = acc + s[v];
        return method == ORC_MEAN ? acc / (float)V : acc;
    }
    /* torch.max over the view axis propagates NaN (models/aggregation.py:76): a NaN sample wins */
    float m = s[0];
    for (int v = 1; v < V; ++v) m = (s[v] > m || s[v] != s[v]) ? s[v] : m;
    if (method == ORC_MAX) return m;
    /* softmax over views of the sampled values themselves, then weighted sum:
     * /root/reference/models/aggregation.py:77-83 (ATen softmax = exp(x-max)/sum) */
    float e[ORC_MAX_VIEWS], S = 0.0f;
    for (int v = 0; v < V; ++v) { e[v] = expf(s[v] - m); S = S + e[v]; }
    float acc = 0.0f;
    for (int v = 0; v < V; ++v) acc = acc + s[v] * (e[v] / S);
    return acc;
}

/* a6-a12: fp32, reference op order.
 * feats (B,V,C,H,W) NCHW fp32; proj (B,V,3,4); coord (B,N,3); out (B,C,N).
 * /root/reference/models/aggregation.py:20-87                               */
int orc_unproject_aggregate_f32(const float *feats, const float *proj, const float *coord,
                                float *out, int B, int V, int C, int H, int W, size_t N, int method)
{
    if (V < 1 || V > ORC_MAX_VIEWS || method < 0 || method > 3) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
#pragma omp parallel for schedule(static)
        for (size_t n = 0; n < N; ++n) {
            const float *xyz = coord + ((size_t)b * N + n) * 3;
            orc_cell_f32 cell[ORC_MAX_VIEWS];
            int invalid[ORC_MAX_VIEWS];
            for (int v = 0; v < V; ++v) {
                float ix, iy;
                invalid[v] = orc_sample_pos(proj + ((size_t)b * V + v) * 12, xyz[0], xyz[1], xyz[2],
                                            H, W, &ix, &iy);
                orc_cell32(ix, iy, H, W, &cell[v]);
            }
            for (int c = 0; c < C; ++c) {
                float s[ORC_MAX_VIEWS];
                for (int v = 0; v < V; ++v) {
                    const float *pl = feats + (((size_t)b * V + v) * C + c) * plane;
                    float t[4];
                    for (int k = 0; k < 4; ++k) t[k] = cell[v].off[k] >= 0 ? pl[cell[v].off[k]] : 0.0f;
                    /* ATen's AVX2 kernel contracts the blend: mul, then three FMAs
                     * (bit-equal to F.grid_sample on 800k samples, this container) */
                    float acc = t[0] * cell[v].wgt[0];
                    acc = fmaf(t[1], cell[v].wgt[1], acc);
                    acc = fmaf(t[2], cell[v].wgt[2], acc);
                    acc = fmaf(t[3], cell[v].wgt[3], acc);
                    s[v] = invalid[v] ? 0.0f : acc;   /* :62 zero after sampling */
                }
                out[((size_t)b * C + c) * N + n] = orc_fuse32(s, V, method);
            }
        }
    }
    return 0;
}

/* Same path in float64 ("truth"): inputs are the fp32 tensors promoted, every
 * operation in double.  Measures the fp32 noise floor of reference and kernel. */
int orc_unproject_aggregate_f64(const float *feats, const float *proj, const float *coord,
                                double *out, int B, int V, int C, int H, int W, size_t N, int method)
{
    if (V < 1 || V > ORC_MAX_VIEWS || method < 0 || method > 3) return -1;
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
#pragma omp parallel for schedule(static)
        for (size_t n = 0; n < N; ++n) {
            const float *xyz = coord + ((size_t)b * N + n) * 3;
            double wgt[ORC_MAX_VIEWS][4];
            ptrdiff_t off[ORC_MAX_VIEWS][4];
            int invalid[ORC_MAX_VIEWS];
            for (int v = 0; v < V; ++v) {
                const float *P = proj + ((size_t)b * V + v) * 12;
                double h[3];
                for (int r = 0; r < 3; ++r)
                    h[r] = (double)xyz[0] * P[4 * r] + (double)xyz[1] * P[4 * r + 1] +
                           (double)xyz[2] * P[4 * r + 2] + (double)P[4 * r + 3];
                invalid[v] = h[2] <= 0.0;
                double w = h[2] == 0.0 ? 1.0 : h[2];
                double gx = 2.0 * (h[0] / w / (double)H - 0.5), gy = 2.0 * (h[1] / w / (double)W - 0.5);
                double ix = (gx + 1.0) * ((double)(W - 1) / 2.0), iy = (gy + 1.0) * ((double)(H - 1) / 2.0);
                double xw = floor(ix), yn = floor(iy);
                double fw = ix - xw, fe = 1.0 - fw, fn = iy - yn, fs = 1.0 - fn;
                wgt[v][0] = fs * fe; wgt[v][1] = fs * fw; wgt[v][2] = fn * fe; wgt[v][3] = fn * fw;
                int x0ok = xw >= 0.0 && xw <= (double)(W - 1), x1ok = xw >= -1.0 && xw <= (double)(W - 2);
                int y0ok = yn >= 0.0 && yn <= (double)(H - 1), y1ok = yn >= -1.0 && yn <= (double)(H - 2);
                ptrdiff_t x0 = x0ok || x1ok ? (ptrdiff_t)xw : 0, y0 = y0ok || y1ok ? (ptrdiff_t)yn : 0;
                off[v][0] = (x0ok && y0ok) ? y0 * W + x0 : -1;
                off[v][1] = (x1ok && y0ok) ? y0 * W + x0 + 1 : -1;
                off[v][2] = (x0ok && y1ok) ? (y0 + 1) * W + x0 : -1;
                off[v][3] = (x1ok && y1ok) ? (y0 + 1) * W + x0 + 1 : -1;
            }
            for (int c = 0; c < C; ++c) {
                double s[ORC_MAX_VIEWS];
                for (int v = 0; v < V; ++v) {
                    const float *pl = feats + (((size_t)b * V + v) * C + c) * plane;
                    double acc = 0.0;
                    for (int k = 0; k < 4; ++k)
                        if (off[v][k] >= 0) acc += (double)pl[off[v][k]] * wgt[v][k];
                        else acc += 0.0 * wgt[v][k];
                    s[v] = invalid[v] ? 0.0 : acc;
                }
                double r;
                if (method == ORC_SUM || method == ORC_MEAN) {
                    r = 0.0;
                    for (int v = 0; v < V; ++v) r += s[v];
                    if (method == ORC_MEAN) r /= (double)V;
                } else {
                    double m = s[0];
                    for (int v = 1; v < V; ++v) m = s[v] > m ? s[v] : m;
                    if (method == ORC_MAX) r = m;
                    else {
                        double S = 0.0, A = 0.0;
                        for (int v = 0; v < V; ++v) { double e = exp(s[v] - m); S += e; A += s[v] * e; }
                        r = A / S;
                    }
                }
                out[((size_t)b * C + c) * N + n] = r;
            }
        }
    }
    return 0;
}

/* a13: 3-D soft-argmax (NOT in the reference — parity unpinned, see header).
 * p = softmax(vol[b,j,:]) ; out[b,j,:] = sum_n p[n] * coord[b,n,:]
 * Elementwise math and accumulation in double: this is the "truth" the fp32
 * CUDA kernel and the fp32 torch restatement are both measured against.     */
void orc_soft_argmax3d_f64(const float *vol, const float *coord, double *out, int B, int J, size_t N)
{
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int b = 0; b < B; ++b)
        for (int j = 0; j < J; ++j) {
            const float *x = vol + ((size_t)b * J + j) * N;
            const float *c = coord + (size_t)b * N * 3;
            double m = x[0];
            for (size_t n = 1; n < N; ++n) m = x[n] > m ? x[n] : m;
            double S = 0.0, A[3] = {0.0, 0.0, 0.0};
            for (size_t n = 0; n < N; ++n) {
                double e = exp((double)x[n] - m);
                S += e; A[0] += e * c[3 * n]; A[1] += e * c[3 * n + 1]; A[2] += e * c[3 * n + 2];
            }
            double *o = out + ((size_t)b * J + j) * 3;
            o[0] = A[0] / S; o[1] = A[1] / S; o[2] = A[2] / S;
        }
}
