/*
 * pn2_oracle.c -- CPU restatement of the reference's PointNet++ geometry path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Parity status: the reference ships no tests or golden vectors for this path
 * (SURVEY.md section 4), so this restatement is pinned against outputs of the
 * reference's own CUDA kernels compiled verbatim into oracle/_ref/ref_cuda.so
 * and run on a B200 (live in tests/test_ops_gpu.py; offline through the fixture
 * tests/golden/ref_cuda_r1.npz made by tests/golden/make_golden.py, checked by
 * tests/test_golden.py), and -- for the lifting and the frustum count -- against
 * vectors produced by the reference's own utils/projection.py run on the host
 * (tests/golden/projection_r2.npz, tests/golden/make_golden_projection.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference tree).  All arithmetic is IEEE fp32 with the contraction the
 * reference's nvcc -O2 build performs spelled out with fmaf():
 *     D(p,q) = fma(dz,dz, fma(dx,dx, rn(dy*dy)))           
 * Build with -ffp-contract=off so the compiler adds no contraction of its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(_OPENMP)
#include <omp.h>
#endif

/* squared distance exactly as the three reference kernels compute it after
 * nvcc's fmad contraction: utils/src/sampling_gpu.cu:133,
 * utils/src/ball_query_gpu.cu:33, utils/src/interpolate_gpu.cu:36 */
static inline float dist_ref(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/* utils/src/cuda_utils.h:10-14 -- largest power of two <= min(n, 1024), >= 1.
 * Kept as an integer loop (the reference's log()/log(2) is exact for these). */
int orc_opt_n_threads(int n) {
    int p = 1;
    while (p * 2 <= n && p * 2 <= 1024) p *= 2;
    return p;
}

static inline uint32_t bitrev(uint32_t v, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

/* Furthest point sampling: utils/src/sampling_gpu.cu:93-209 (kernel),
 * :211-253 (block size choice), model/pointnet2_utils.py:25-28 (temp=1e10).
 * The block tree reduction (:86-91,:143-203) keeps the lower slot on ties, and
 * each thread keeps its first strict maximum (:136-137), which together give
 * the total order: larger temp first, then smaller bitrev(k mod bs), then
 * smaller k.  idx[0] = 0 (:113-115).  m <= 0 writes nothing (:101). */
void orc_fps(int B, int N, int M, const float *xyz, int32_t *idx) {
    if (M <= 0 || N <= 0) return;
    int bs = orc_opt_n_threads(N);
    int lg = 0;
    while ((1 << lg) < bs) ++lg;
    uint32_t *rank = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)bs);
    for (int r = 0; r < bs; ++r) rank[r] = bitrev((uint32_t)r, lg);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const float *p = xyz + (size_t)b * N * 3;
        int32_t *o = idx + (size_t)b * M;
        float *temp = (float *)malloc(sizeof(float) * (size_t)N);
        for (int k = 0; k < N; ++k) temp[k] = 1e10f;
        int old = 0;
        o[0] = 0;
        for (int j = 1; j < M; ++j) {
            float x1 = p[old * 3], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            float best = -1.0f;
            uint32_t best_rank = 0;
            int besti = 0;
            for (int k = 0; k < N; ++k) {
                float d = dist_ref(p[k * 3], p[k * 3 + 1], p[k * 3 + 2], x1, y1, z1);
                float d2 = d < temp[k] ? d : temp[k]; /* min(d, temp[k]) :134 */
                temp[k] = d2;
                uint32_t rk = rank[k & (bs - 1)];
                /* ascending k, so on full ties the earlier k is kept */
                if (d2 > best || (d2 == best && rk < best_rank)) {
                    best = d2; best_rank = rk; besti = k;
                }
            }
            old = besti;
            o[j] = old;
        }
        free(temp);
    }
    free(rank);
}

/* gather: utils/src/sampling_gpu.cu:8-24.  points (B,C,N), idx (B,M) -> (B,C,M) */
void orc_gather(int B, int C, int N, int M, const float *points, const int32_t *idx, float *out) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < M; ++j)
                out[((size_t)b * C + c) * M + j] = points[((size_t)b * C + c) * N + idx[(size_t)b * M + j]];
}

/* gather backward: utils/src/sampling_gpu.cu:46-63 (scatter-add; the reference's
 * atomicAdd order is unspecified, here ascending j). grad_points pre-zeroed by caller. */
void orc_gather_grad(int B, int C, int N, int M, const float *grad_out, const int32_t *idx, float *grad_points) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < M; ++j)
                grad_points[((size_t)b * C + c) * N + idx[(size_t)b * M + j]] += grad_out[((size_t)b * C + c) * M + j];
}

/* ball query: utils/src/ball_query_gpu.cu:9-45; idx pre-zeroed by the Python
 * caller (model/pointnet2_utils.py:216) -- done here so an empty ball is zeros.
 * radius2 = rn(radius*radius) in fp32 (:23), strict '<' (:34). */
void orc_ball_query(int B, int N, int M, float radius, int K, const float *new_xyz, const float *xyz, int32_t *idx) {
    float r2 = radius * radius;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int q = 0; q < M; ++q) {
            const float *c = new_xyz + ((size_t)b * M + q) * 3;
            const float *p = xyz + (size_t)b * N * 3;
            int32_t *o = idx + ((size_t)b * M + q) * K;
            for (int l = 0; l < K; ++l) o[l] = 0;
            int cnt = 0;
            for (int k = 0; k < N && cnt < K; ++k) {
                /* the reference computes (new - x); squares make the sign irrelevant */
                float d2 = dist_ref(c[0], c[1], c[2], p[k * 3], p[k * 3 + 1], p[k * 3 + 2]);
                if (d2 < r2) {
                    if (cnt == 0) for (int l = 0; l < K; ++l) o[l] = k;
                    o[cnt++] = k;
                }
            }
        }
    }
}

/* grouping: utils/src/group_points_gpu.cu:47-66. points (B,C,N), idx (B,P,S) -> (B,C,P,S) */
void orc_group(int B, int C, int N, int P, int S, const float *points, const int32_t *idx, float *out) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < P * S; ++j)
                out[((size_t)b * C + c) * P * S + j] = points[((size_t)b * C + c) * N + idx[(size_t)b * P * S + j]];
}

/* grouping backward: utils/src/group_points_gpu.cu:8-25 */
void orc_group_grad(int B, int C, int N, int P, int S, const float *grad_out, const int32_t *idx, float *grad_points) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < P * S; ++j)
                grad_points[((size_t)b * C + c) * N + idx[(size_t)b * P * S + j]] += grad_out[((size_t)b * C + c) * P * S + j];
}

/* three_nn: utils/src/interpolate_gpu.cu:9-52.  best* are doubles initialised
 * to 1e40 holding float values, strict '<' cascade; dist2 written as float
 * (1e40 -> +inf).  Returns SQUARED distances; the sqrt is taken by the Python
 * caller (model/pointnet2_utils.py:97). */
void orc_three_nn(int B, int n, int m, const float *unknown, const float *known, float *dist2, int32_t *idx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < n; ++i) {
            const float *u = unknown + ((size_t)b * n + i) * 3;
            const float *kn = known + (size_t)b * m * 3;
            double b1 = 1e40, b2 = 1e40, b3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                float d = dist_ref(u[0], u[1], u[2], kn[k * 3], kn[k * 3 + 1], kn[k * 3 + 2]);
                if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
                else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
                else if (d < b3) { b3 = d; i3 = k; }
            }
            float *od = dist2 + ((size_t)b * n + i) * 3;
            int32_t *oi = idx + ((size_t)b * n + i) * 3;
            od[0] = (float)b1; od[1] = (float)b2; od[2] = (float)b3;
            oi[0] = i1; oi[1] = i2; oi[2] = i3;
        }
    }
}

/* three_interpolate: utils/src/interpolate_gpu.cu:77-97; nvcc contracts the sum
 * to fma(w2,p2, fma(w0,p0, rn(w1*p1))).
 * points (B,C,m), idx/weight (B,n,3) -> out (B,C,n) */
void orc_three_interpolate(int B, int C, int m, int n, const float *points, const int32_t *idx, const float *weight, float *out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const float *f = points + ((size_t)b * C + c) * m;
            for (int i = 0; i < n; ++i) {
                const int32_t *id = idx + ((size_t)b * n + i) * 3;
                const float *w = weight + ((size_t)b * n + i) * 3;
                float t = w[1] * f[id[1]];
                t = fmaf(w[0], f[id[0]], t);
                out[((size_t)b * C + c) * n + i] = fmaf(w[2], f[id[2]], t);
            }
        }
}

/* three_interpolate backward: utils/src/interpolate_gpu.cu:120-142 */
void orc_three_interpolate_grad(int B, int C, int n, int m, const float *grad_out, const int32_t *idx, const float *weight, float *grad_points) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            float *g = grad_points + ((size_t)b * C + c) * m;
            for (int i = 0; i < n; ++i) {
                const int32_t *id = idx + ((size_t)b * n + i) * 3;
                const float *w = weight + ((size_t)b * n + i) * 3;
                float go = grad_out[((size_t)b * C + c) * n + i];
                g[id[0]] += go * w[0];
                g[id[1]] += go * w[1];
                g[id[2]] += go * w[2];
            }
        }
}

/* ---------------------------------------------------------------------------
 * Multi-view lifting (utils/projection.py:97-130,166-230,237-256 and
 * model/pointnet2multiview.py:30-43,83-102), restated per (cloud, view, point).
 *
 * The reference evaluates these steps with torch.mm / elementwise torch ops
 * whose fp32 summation order is a BLAS implementation detail; this restatement
 * (and the CUDA kernel it checks) fixes the order as the fma chains below.
 * tests/test_lifting.py measures the index agreement of this order against a
 * torch restatement of the reference's op sequence.
 *
 * Per-view inputs prepared by the caller exactly as the reference prepares them
 * (torch.inverse, compute_frustum_corners, compute_frustum_normals):
 *   w2c     (16)  world_to_camera row-major
 *   corner2 (3), corner4 (3)   frustum corners 2 and 4 (world)
 *   normals (6*3)
 * intr = {fx, fy, cx, cy}; image W x H; depth stored [H][W].
 * Returns the pixel index (v*W+u) a point is lifted from, or -1.
 * ------------------------------------------------------------------------- */
static inline float rint_f(float v) { return rintf(v); } /* round-half-even == torch.round */

int32_t orc_project_point(const float *p, const float *w2c, const float *c2, const float *c4,
                          const float *normals, const float *intr, int W, int H,
                          const float *depth, float dmin, float dmax, float acc) {
    /* frustum test: utils/projection.py:108-120; round(100*s)/100 < 0 */
    for (int k = 0; k < 6; ++k) {
        const float *c = k < 3 ? c2 : c4;
        const float *nr = normals + 3 * k;
        float dx = p[0] - c[0], dy = p[1] - c[1], dz = p[2] - c[2];
        float s = dx * nr[0];
        s = fmaf(dy, nr[1], s);
        s = fmaf(dz, nr[2], s);
        float r = rint_f(s * 100.0f) / 100.0f;
        if (!(r < 0.0f)) return -1;
    }
    /* world -> camera: utils/projection.py:199 (rows 0..2 of w2c @ [p;1]) */
    float cam[3];
    for (int r = 0; r < 3; ++r) {
        const float *m = w2c + 4 * r;
        float s = m[0] * p[0];
        s = fmaf(m[1], p[1], s);
        s = fmaf(m[2], p[2], s);
        cam[r] = s + m[3]; /* m[3] * 1 */
    }
    /* pinhole: utils/projection.py:202-204 (multiply, divide, add; round half-even) */
    float u = (cam[0] * intr[0]) / cam[2] + intr[2];
    float v = (cam[1] * intr[1]) / cam[2] + intr[3];
    float ur = rint_f(u), vr = rint_f(v);
    /* bounds: utils/projection.py:207 (NaN fails every comparison) */
    if (!(ur >= 0.0f && vr >= 0.0f && ur < (float)W && vr < (float)H)) return -1;
    int pix = (int)vr * W + (int)ur;
    /* depth test: utils/projection.py:215-216 */
    float z = depth[pix];
    if (!(z >= dmin && z <= dmax && fabsf(z - cam[2]) <= acc)) return -1;
    return pix;
}

/* Batched lifting.  points (B,N,3); feats (B,V,C,H,W) channel-major as ENet
 * produces them; depth (B,V,H,W); w2c (B,V,16); corner2/corner4 (B,V,3);
 * normals (B,V,18).  out (B,C,N).  pix_out (B,V,N) int32 (or NULL).
 * reduce_first = 0: max over views with zeros for invisible views
 *   (model/pointnet2multiview.py:39);
 * reduce_first = 1: first view, later views fill points whose C channels are
 *   all exactly zero (model/pointnet2multiview.py:93-98). */
void orc_lift_views(int B, int N, int V, int C, int H, int W, const float *points, const float *feats,
                    const float *depth, const float *w2c, const float *corner2, const float *corner4,
                    const float *normals, const float *intr, float dmin, float dmax, float acc,
                    int reduce_first, float *out, int32_t *pix_out) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < N; ++i) {
            const float *p = points + ((size_t)b * N + i) * 3;
            int have = 0; /* reduce_first: current value is not all-zero */
            for (int c = 0; c < C; ++c) out[((size_t)b * C + c) * N + i] = 0.0f;
            for (int v = 0; v < V; ++v) {
                size_t bv = (size_t)b * V + v;
                int32_t pix = orc_project_point(p, w2c + bv * 16, corner2 + bv * 3, corner4 + bv * 3,
                                                normals + bv * 18, intr, W, H, depth + bv * H * W, dmin, dmax, acc);
                if (pix_out) pix_out[bv * N + i] = pix;
                if (reduce_first) {
                    if (have) continue;
                    /* candidate column (zeros when not visible) replaces an all-zero column */
                    if (pix < 0) continue;
                    int nz = 0;
                    for (int c = 0; c < C; ++c) {
                        float f = feats[(bv * C + c) * H * W + pix];
                        out[((size_t)b * C + c) * N + i] = f;
                        nz |= (f != 0.0f);
                    }
                    have = nz;
                } else {
                    for (int c = 0; c < C; ++c) {
                        float f = pix >= 0 ? feats[(bv * C + c) * H * W + pix] : 0.0f;
                        float *o = out + ((size_t)b * C + c) * N + i;
                        if (v == 0) *o = f; else *o = f > *o ? f : *o;
                    }
                }
            }
        }
    }
}

/* Frustum membership count of the ScanNet loader's best-view selection:
 * data_utils/ScanNetDataLoader.py:91-97 calls points_in_frustum_cpu
 * (utils/projection.py:132-164) with corners / normals / points cast to fp64.
 * corners (P,8,3) and normals (P,6,3) are the fp32 values of
 * compute_frustum_corners / compute_frustum_normals; planes 0-2 are tested
 * against corner 2, planes 3-5 against corner 4; predicate
 * round(100 * s) / 100 < 0 in fp64.  counts (P); mask (P,N) bytes or NULL. */
void orc_frustum_count(int N, int P, const float *points, const float *corners, const float *normals,
                       int64_t *counts, uint8_t *mask) {
#pragma omp parallel for schedule(static)
    for (int q = 0; q < P; ++q) {
        const float *c2 = corners + ((size_t)q * 8 + 2) * 3, *c4 = corners + ((size_t)q * 8 + 4) * 3;
        int64_t cnt = 0;
        for (int i = 0; i < N; ++i) {
            const float *p = points + (size_t)i * 3;
            int inside = 1;
            for (int k = 0; k < 6; ++k) {
                const float *c = k < 3 ? c2 : c4;
                const float *nr = normals + ((size_t)q * 6 + k) * 3;
                double s = ((double)p[0] - (double)c[0]) * (double)nr[0];
                s += ((double)p[1] - (double)c[1]) * (double)nr[1];
                s += ((double)p[2] - (double)c[2]) * (double)nr[2];
                if (!(rint(s * 100.0) / 100.0 < 0.0)) inside = 0;
            }
            cnt += inside;
            if (mask) mask[(size_t)q * N + i] = (uint8_t)inside;
        }
        counts[q] = cnt;
    }
}

/* ---------------------------------------------------------------------------
 * Shared-MLP helper for the CPU baseline: y[r][co] = act(sum_ci x[r][ci]*w[co][ci] + b[co])
 * (the 1x1 Conv + folded eval BatchNorm + ReLU of model/pointnet_util.py:105-107,
 * 218-220).  Plain fp32, ascending-ci accumulation.  rows are (point) or
 * (centroid, sample) pairs, channel-last.
 * ------------------------------------------------------------------------- */
void orc_mlp_layer(int rows, int cin, int cout, const float *x, const float *w, const float *bias, int relu, float *y) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; ++r) {
        const float *xr = x + (size_t)r * cin;
        float *yr = y + (size_t)r * cout;
        for (int co = 0; co < cout; ++co) {
            const float *wr = w + (size_t)co * cin;
            float s = 0.0f;
            for (int ci = 0; ci < cin; ++ci) s += xr[ci] * wr[ci];
            s += bias[co];
            yr[co] = (relu && s < 0.0f) ? 0.0f : s;
        }
    }
}

int orc_num_threads(void) {
#if defined(_OPENMP)
    return omp_get_max_threads();
#else
    return 1;
#endif
}
