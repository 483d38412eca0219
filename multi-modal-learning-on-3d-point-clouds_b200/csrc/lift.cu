// lift.cu -- multi-view 2D->3D feature lifting in one launch per batch.
//
// Replaces, per (cloud, view): ProjectionHelper.compute_projection (utils/projection.py:166-230,
// ~25 small torch ops and >= 4 host<->device round trips), Projection.forward (:237-256) and the
// per-sample view reduction of PointNet2Multiview* (model/pointnet2multiview.py:30-43, 83-102).
//
// Per point and view: six frustum planes with round(100*s)/100 < 0; camera
// transform; pinhole projection (multiply, divide, add); round-half-even to the NEAREST pixel (no
// interpolation, utils/projection.py:204); bounds; depth-range and |depth - z| <= accuracy tests;
// then the C-channel feature column of that pixel is fetched and reduced over views (max with
// zeros for invisible views, or first view whose column is not all-zero).
//
// fp32 evaluation order (the reference leaves it to BLAS): dot products are
// fma(c, z, fma(b, y, rn(a*x))) (+ rn add of the translation); identical in oracle/pn2_oracle.c.
#include "common.cuh"

namespace pn2 {
namespace {

constexpr int LV_THREADS = 128;
constexpr int LV_MAXV = 8;

struct ViewParams {  // 38 floats per view, staged in shared memory
    float w2c[12];   // rows 0..2 of world_to_camera
    float c2[3], c4[3];
    float nrm[18];
};

__device__ __forceinline__ int project_point(float px, float py, float pz, const ViewParams &vp, float fx, float fy,
                                             float cx, float cy, int W, int H, const float *__restrict__ depth,
                                             float dmin, float dmax, float acc) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float *c = k < 3 ? vp.c2 : vp.c4;
        const float dx = __fsub_rn(px, c[0]), dy = __fsub_rn(py, c[1]), dz = __fsub_rn(pz, c[2]);
        const float s = __fmaf_rn(dz, vp.nrm[3 * k + 2], __fmaf_rn(dy, vp.nrm[3 * k + 1], __fmul_rn(dx, vp.nrm[3 * k])));
        const float r = __fdiv_rn(rintf(__fmul_rn(s, 100.0f)), 100.0f);
        if (!(r < 0.0f)) return -1;
    }
    float cam[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float *m = vp.w2c + 4 * r;
        cam[r] = __fadd_rn(__fmaf_rn(m[2], pz, __fmaf_rn(m[1], py, __fmul_rn(m[0], px))), m[3]);
    }
    const float u = __fadd_rn(__fdiv_rn(__fmul_rn(cam[0], fx), cam[2]), cx);
    const float v = __fadd_rn(__fdiv_rn(__fmul_rn(cam[1], fy), cam[2]), cy);
    const float ur = rintf(u), vr = rintf(v);
    if (!(ur >= 0.0f && vr >= 0.0f && ur < (float)W && vr < (float)H)) return -1;
    const int pix = (int)vr * W + (int)ur;
    const float z = __ldg(depth + pix);
    if (!(z >= dmin && z <= dmax && fabsf(__fsub_rn(z, cam[2])) <= acc)) return -1;
    return pix;
}

__global__ void __launch_bounds__(LV_THREADS)
lift_views_kernel(int n, int nv, int c, int h, int w, const float *__restrict__ points, const float *__restrict__ feats,
                  const float *__restrict__ depth, const float *__restrict__ w2c, const float *__restrict__ corner2,
                  const float *__restrict__ corner4, const float *__restrict__ normals, float fx, float fy, float cx,
                  float cy, float dmin, float dmax, float acc, int reduce, float *__restrict__ out,
                  int32_t *__restrict__ pix_out, int32_t *__restrict__ count) {
    __shared__ ViewParams vps[LV_MAXV];
    __shared__ int vcount[LV_MAXV];
    const int b = blockIdx.y;
    const int hw = h * w;
    if (threadIdx.x < nv) {
        const size_t bv = (size_t)b * nv + threadIdx.x;
        ViewParams &vp = vps[threadIdx.x];
        for (int k = 0; k < 12; ++k) vp.w2c[k] = w2c[bv * 16 + k];
        for (int k = 0; k < 3; ++k) {
            vp.c2[k] = corner2[bv * 3 + k];
            vp.c4[k] = corner4[bv * 3 + k];
        }
        for (int k = 0; k < 18; ++k) vp.nrm[k] = normals[bv * 18 + k];
        vcount[threadIdx.x] = 0;
    }
    __syncthreads();
    const int i = blockIdx.x * LV_THREADS + threadIdx.x;
    const bool live = i < n;
    int pix[LV_MAXV];
    if (live) {
        const float *p = points + ((size_t)b * n + i) * 3;
        const float px = p[0], py = p[1], pz = p[2];
#pragma unroll
        for (int v = 0; v < LV_MAXV; ++v) {
            pix[v] = -1;
            if (v < nv) {
                const size_t bv = (size_t)b * nv + v;
                pix[v] = project_point(px, py, pz, vps[v], fx, fy, cx, cy, w, h, depth + bv * hw, dmin, dmax, acc);
                if (pix_out) pix_out[bv * n + i] = pix[v];
                if (count && pix[v] >= 0) atomicAdd(&vcount[v], 1);
            }
        }
    }
    if (count) {
        __syncthreads();
        if (threadIdx.x < nv && vcount[threadIdx.x]) atomicAdd(count + (size_t)b * nv + threadIdx.x, vcount[threadIdx.x]);
    }
    if (!live) return;
    float *o = out + (size_t)b * c * n + i;
    if (reduce == PN2_REDUCE_MAX) {
        // F.max_pool1d over the stacked per-view maps: invisible views contribute zeros
        for (int ch = 0; ch < c; ++ch) {
            float best = 0.f;
            bool any = false;
#pragma unroll
            for (int v = 0; v < LV_MAXV; ++v) {
                if (v < nv) {
                    const float f = pix[v] >= 0 ? __ldg(feats + (((size_t)b * nv + v) * c + ch) * hw + pix[v]) : 0.f;
                    best = any ? fmaxf(best, f) : f;
                    any = true;
                }
            }
            __stcs(o + (size_t)ch * n, best);
        }
    } else {
        bool have = false;
#pragma unroll
        for (int v = 0; v < LV_MAXV; ++v) {
            if (v < nv && !have && pix[v] >= 0) {
                const float *f = feats + ((size_t)b * nv + v) * c * hw + pix[v];
                bool nz = false;
                for (int ch = 0; ch < c; ++ch) {
                    const float val = __ldg(f + (size_t)ch * hw);
                    nz = nz || (val != 0.f);
                    o[(size_t)ch * n] = val;
                }
                have = nz;
                // a visible but all-zero column stays replaceable by a later view (pointnet2multiview.py:97)
                if (!nz) have = false;
            }
        }
        bool wrote = false;
#pragma unroll
        for (int v = 0; v < LV_MAXV; ++v) wrote = wrote || (v < nv && pix[v] >= 0);
        if (!wrote)
            for (int ch = 0; ch < c; ++ch) o[(size_t)ch * n] = 0.f;
    }
}

// Best-view selection of the ScanNet loader (data_utils/ScanNetDataLoader.py:87-105, 260-278): how many of the N crop
// points fall inside the frustum of EACH of the scene's P camera poses.  The reference calls points_in_frustum_cpu
// (utils/projection.py:132-164) once per pose file on the CPU in fp64; here all P x N tests are one launch, in fp64
// like the reference (same predicate as the lifting: round(100 * s) / 100 < 0 for the six planes).
constexpr int FC_THREADS = 256;
constexpr int FC_POSES = 32;  // poses per CTA (shared-memory tile)

__global__ void __launch_bounds__(FC_THREADS)
frustum_count_kernel(int n, int np, const float *__restrict__ points, const float *__restrict__ corner2,
                     const float *__restrict__ corner4, const float *__restrict__ normals, int32_t *__restrict__ counts) {
    __shared__ double sc2[FC_POSES][3], sc4[FC_POSES][3], sn[FC_POSES][18];
    __shared__ int cnt[FC_POSES];
    const int p0 = blockIdx.y * FC_POSES;
    const int pc = min(FC_POSES, np - p0);
    for (int e = threadIdx.x; e < pc * 18; e += FC_THREADS) sn[e / 18][e % 18] = (double)normals[(size_t)p0 * 18 + e];
    for (int e = threadIdx.x; e < pc * 3; e += FC_THREADS) {
        sc2[e / 3][e % 3] = (double)corner2[(size_t)p0 * 3 + e];
        sc4[e / 3][e % 3] = (double)corner4[(size_t)p0 * 3 + e];
    }
    if (threadIdx.x < FC_POSES) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * FC_THREADS + threadIdx.x;
    const bool live = i < n;
    double px = 0, py = 0, pz = 0;
    if (live) {
        px = (double)points[3 * (size_t)i];
        py = (double)points[3 * (size_t)i + 1];
        pz = (double)points[3 * (size_t)i + 2];
    }
    for (int q = 0; q < pc; ++q) {
        bool inside = live;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double *c = k < 3 ? sc2[q] : sc4[q];
            const double s = (px - c[0]) * sn[q][3 * k] + (py - c[1]) * sn[q][3 * k + 1] + (pz - c[2]) * sn[q][3 * k + 2];
            inside = inside && (rint(s * 100.0) / 100.0 < 0.0);
        }
        const unsigned m = __ballot_sync(0xffffffffu, inside);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&cnt[q], __popc(m));
    }
    __syncthreads();
    if (threadIdx.x < pc && cnt[threadIdx.x]) atomicAdd(counts + p0 + threadIdx.x, cnt[threadIdx.x]);
}

}  // namespace
}  // namespace pn2

extern "C" int pn2_frustum_count(int n, int num_poses, const float *points, const float *corner2, const float *corner4,
                                 const float *normals, int32_t *counts, void *stream) {
    using namespace pn2;
    PN2_REQUIRE(n >= 0 && num_poses >= 0, "frustum_count: bad dims");
    if (n == 0 || num_poses == 0) return PN2_OK;
    PN2_REQUIRE(points && corner2 && corner4 && normals && counts, "frustum_count: null pointer");
    PN2_REQUIRE(ceil_div(num_poses, FC_POSES) <= 65535, "frustum_count: too many poses");
    dim3 grid(ceil_div(n, FC_THREADS), ceil_div(num_poses, FC_POSES));
    frustum_count_kernel<<<grid, FC_THREADS, 0, (cudaStream_t)stream>>>(n, num_poses, points, corner2, corner4, normals, counts);
    PN2_LAUNCH_OK("frustum_count");
    return PN2_OK;
}

extern "C" int pn2_lift_views(int b, int n, int v, int c, int h, int w, const float *points, const float *feats,
                              const float *depth, const float *w2c, const float *corner2, const float *corner4,
                              const float *normals, const float *intr, float depth_min, float depth_max,
                              float accuracy, int reduce, float *out, int32_t *pix, int32_t *count, void *stream) {
    using namespace pn2;
    PN2_REQUIRE(b >= 0 && n >= 0 && v >= 1 && c >= 0 && h >= 1 && w >= 1, "lift_views: bad dims");
    if (v > LV_MAXV) return set_error(PN2_ERR_UNSUPPORTED, "lift_views: at most %d views per cloud (got %d)", LV_MAXV, v);
    PN2_REQUIRE(reduce == PN2_REDUCE_MAX || reduce == PN2_REDUCE_FIRST, "lift_views: unknown reduce %d", reduce);
    if (b == 0 || n == 0) return PN2_OK;
    PN2_REQUIRE(points && depth && w2c && corner2 && corner4 && normals && intr && ((feats && out) || c == 0), "lift_views: null pointer");
    PN2_REQUIRE(b <= 65535, "lift_views: b exceeds the grid limit");
    dim3 grid(ceil_div(n, LV_THREADS), b);
    lift_views_kernel<<<grid, LV_THREADS, 0, (cudaStream_t)stream>>>(n, v, c, h, w, points, feats, depth, w2c, corner2, corner4,
                                                                     normals, intr[0], intr[1], intr[2], intr[3], depth_min,
                                                                     depth_max, accuracy, reduce, out, pix, count);
    PN2_LAUNCH_OK("lift_views");
    return PN2_OK;
}
