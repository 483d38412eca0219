// sphere.cu -- bounding-sphere IoU and greedy non-maximum suppression of the nuScenes proposal layer.
//
// Replaces model/pointmaskrcnn.py:233-288 (iou_spheres: ~20 torch ops with two nonzero() host round trips per call) and
// :290-321 (nms: a Python loop with one host synchronisation per kept sphere).  Semantics kept:
//   iou(a, b):  d = |c_a - c_b|;  d <= |r_a - r_b|  -> (min r / max r)^3          (one sphere inside the other)
//               |r_a - r_b| < d < r_a + r_b         -> I / (4/3 pi (r_a^3 + r_b^3) - I)
//                                                       I = (r_a + r_b - d)^2 (d^2 + 2 d (r_a + r_b) - 3 (r_a - r_b)^2) pi / (12 d)
//               otherwise 0;   every operation rounded to fp32 in the reference's order.
//   nms: visit the spheres by descending score; keep one, drop every later sphere whose IoU with it is > threshold
//        (the reference keeps `iou <= threshold`), until none is left.  Equal scores: lower index first (torch.sort leaves
//        that order unspecified).  Kept indices are returned in selection order.
#include "common.cuh"

namespace pn2 {
namespace {

__device__ __forceinline__ float sphere_iou(float ax, float ay, float az, float ar, float bx, float by, float bz, float br) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const float diff = fabsf(__fsub_rn(ar, br));
    const float sum = __fadd_rn(ar, br);
    if (d <= diff) {
        const float q = __fdiv_rn(fminf(ar, br), fmaxf(ar, br));  // (1.0 * min_r / max_r) ** 3
        return __fmul_rn(__fmul_rn(q, q), q);
    }
    if (d > diff && d < sum) {
        const float t = __fsub_rn(sum, d);
        float inter = __fmul_rn(t, t);
        const float rd = __fsub_rn(ar, br);
        const float poly = __fsub_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(__fmul_rn(2.0f, d), sum)), __fmul_rn(3.0f, __fmul_rn(rd, rd)));
        inter = __fmul_rn(inter, poly);
        inter = __fmul_rn(inter, __fdiv_rn(3.14159265358979323846f, __fmul_rn(12.0f, d)));
        const float cubes = __fadd_rn(__fmul_rn(__fmul_rn(ar, ar), ar), __fmul_rn(__fmul_rn(br, br), br));
        const float uni = __fsub_rn(__fmul_rn(4.18879020478639098f, cubes), inter);  // 4/3. * np.pi as one fp32 constant
        return __fdiv_rn(inter, uni);
    }
    return 0.0f;
}

__global__ void sphere_iou_kernel(int m, int n, const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ iou) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)m * n) return;
    const int i = (int)(e / n), j = (int)(e % n);
    const float4 sa = *reinterpret_cast<const float4 *>(a + 4 * (size_t)i);
    const float4 sb = *reinterpret_cast<const float4 *>(b + 4 * (size_t)j);
    iou[e] = sphere_iou(sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w);
}

constexpr int NMS_THREADS = 256;
constexpr int NMS_MAX_N = 4096;

// one CTA per problem: sort by (score desc, index asc), then the greedy sweep with the survivors flagged in shared memory
__global__ void __launch_bounds__(NMS_THREADS)
sphere_nms_kernel(int n, float threshold, const float *__restrict__ spheres_all, const float *__restrict__ scores_all,
                  const int32_t *__restrict__ count_in, int32_t *__restrict__ keep_all, int32_t *__restrict__ count_out) {
    __shared__ unsigned long long keys[NMS_MAX_N];
    __shared__ unsigned char alive[NMS_MAX_N];
    __shared__ int cursor;
    const int b = blockIdx.x, t = threadIdx.x;
    const int nb = count_in ? min(count_in[b], n) : n;  // problems of a batch may hold fewer than n spheres
    const float *sp = spheres_all + (size_t)b * n * 4;
    const float *sc = scores_all + (size_t)b * n;
    int np2 = 1;
    while (np2 < nb) np2 <<= 1;
    for (int k = t; k < np2; k += NMS_THREADS) {
        unsigned long long key = ~0ull;  // padding sorts last
        if (k < nb) {
            // descending score = ascending key: flip a monotone float -> uint map; index breaks ties (lower first)
            uint32_t u = __float_as_uint(sc[k]);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending order of u == ascending order of the float
            key = ((unsigned long long)(~u) << 32) | (unsigned)k;
        }
        keys[k] = key;
        if (k < NMS_MAX_N) alive[k] = 1;
    }
    __syncthreads();
    for (int size = 2; size <= np2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = t; i < np2 / 2; i += NMS_THREADS) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long x = keys[lo], y = keys[hi];
                if ((x > y) == up) {
                    keys[lo] = y;
                    keys[hi] = x;
                }
            }
            __syncthreads();
        }
    if (t == 0) cursor = 0;
    __syncthreads();
    int32_t *keep = keep_all + (size_t)b * n;
    int kept = 0;
    for (int p = 0; p < nb; ++p) {
        if (!alive[p]) continue;  // uniform: every thread reads the same flag after the barrier below
        const int i = (int)(unsigned)keys[p];
        if (t == 0) keep[kept] = i;
        ++kept;
        const float4 si = *reinterpret_cast<const float4 *>(sp + 4 * (size_t)i);
        for (int q = p + 1 + t; q < nb; q += NMS_THREADS) {
            if (!alive[q]) continue;
            const int j = (int)(unsigned)keys[q];
            const float4 sj = *reinterpret_cast<const float4 *>(sp + 4 * (size_t)j);
            // iou_table[i, j] of the reference: first argument = the kept sphere
            if (!(sphere_iou(si.x, si.y, si.z, si.w, sj.x, sj.y, sj.z, sj.w) <= threshold)) alive[q] = 0;
        }
        __syncthreads();
    }
    if (t == 0) count_out[b] = kept;
    for (int k = kept + t; k < n; k += NMS_THREADS) keep[k] = -1;
}

}  // namespace
}  // namespace pn2

extern "C" int pn2_sphere_iou(int m, int n, const float *spheres_a, const float *spheres_b, float *iou, void *stream) {
    using namespace pn2;
    PN2_REQUIRE(m >= 0 && n >= 0, "sphere_iou: bad dims m=%d n=%d", m, n);
    if (m == 0 || n == 0) return PN2_OK;
    PN2_REQUIRE(spheres_a && spheres_b && iou, "sphere_iou: null pointer");
    PN2_REQUIRE((((uintptr_t)spheres_a | (uintptr_t)spheres_b) & 15) == 0, "sphere_iou: sphere arrays must be 16-byte aligned");
    const long long total = (long long)m * n;
    sphere_iou_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(m, n, spheres_a, spheres_b, iou);
    PN2_LAUNCH_OK("sphere_iou_kernel");
    return PN2_OK;
}

extern "C" int pn2_sphere_nms(int b, int n, const float *spheres, const float *scores, const int32_t *count_in, float threshold,
                              int32_t *keep, int32_t *count_out, void *stream) {
    using namespace pn2;
    PN2_REQUIRE(b >= 0 && n >= 0, "sphere_nms: bad dims b=%d n=%d", b, n);
    if (n > NMS_MAX_N) return set_error(PN2_ERR_UNSUPPORTED, "sphere_nms: at most %d spheres per problem (got %d)", NMS_MAX_N, n);
    if (b == 0) return PN2_OK;
    PN2_REQUIRE(count_out && (n == 0 || (spheres && scores && keep)), "sphere_nms: null pointer");
    PN2_REQUIRE((((uintptr_t)spheres) & 15) == 0, "sphere_nms: spheres must be 16-byte aligned");
    if (n == 0) {
        PN2_CUDA(cudaMemsetAsync(count_out, 0, (size_t)b * sizeof(int32_t), (cudaStream_t)stream));
        return PN2_OK;
    }
    sphere_nms_kernel<<<b, NMS_THREADS, 0, (cudaStream_t)stream>>>(n, threshold, spheres, scores, count_in, keep, count_out);
    PN2_LAUNCH_OK("sphere_nms_kernel");
    return PN2_OK;
}
