// ball_query.cu -- radius neighbour search.  Replaces ball_query_kernel_fast / launcher
// (utils/src/ball_query_gpu.cu:9-62).
//
// Result contract: for query q, scanning k = 0..N-1 ascending, the first `nsample`
// points with D(q,k) < rn(radius*radius) (strict, fp32, reference contraction order); unused slots
// repeat the first hit; a query with no hit gives zeros.
//
// The reference runs ONE THREAD per query over all N points (divergent early exit, M/256 CTAs per
// cloud).  Here a WARP owns a query: the 32 lanes test 32 consecutive points per step from a
// shared-memory tile (SoA, conflict-free), `ballot` + `popc` keep the hits in index order, and the
// warp stops as soon as the ball is full.  A CTA stages each xyz tile once for all of its queries.
#include "common.cuh"

namespace pn2 {
namespace {

constexpr int BQ_WARPS = 8;
constexpr int BQ_THREADS = BQ_WARPS * 32;
constexpr int BQ_QPW = 4;                      // queries per warp
constexpr int BQ_QPB = BQ_WARPS * BQ_QPW;      // queries per block
constexpr int BQ_TILE = 2048;                  // points per shared-memory tile (24 KB)

__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(int n, int m, float radius, int nsample, const float *__restrict__ new_xyz_all,
                  const float *__restrict__ xyz_all, int32_t *__restrict__ idx_all) {
    __shared__ float tx[BQ_TILE], ty[BQ_TILE], tz[BQ_TILE];
    const int b = blockIdx.y;
    const float *xyz = xyz_all + (size_t)b * n * 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = blockIdx.x * BQ_QPB + warp * BQ_QPW;
    const float r2 = __fmul_rn(radius, radius);  // ball_query_gpu.cu:23
    const uint32_t lt_mask = (1u << lane) - 1u;

    float qx[BQ_QPW], qy[BQ_QPW], qz[BQ_QPW];
    int cnt[BQ_QPW], first[BQ_QPW];
#pragma unroll
    for (int i = 0; i < BQ_QPW; ++i) {
        const int q = min(q0 + i, m - 1);
        const float *c = new_xyz_all + ((size_t)b * m + q) * 3;
        qx[i] = c[0];
        qy[i] = c[1];
        qz[i] = c[2];
        cnt[i] = (q0 + i < m) ? 0 : nsample;  // out-of-range queries are "done"
        first[i] = 0;
    }

    for (int base = 0; base < n; base += BQ_TILE) {
        const int tn = min(BQ_TILE, n - base);
        for (int p = threadIdx.x; p < tn; p += BQ_THREADS) {
            const float *s = xyz + (size_t)(base + p) * 3;
            tx[p] = s[0];
            ty[p] = s[1];
            tz[p] = s[2];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < BQ_QPW; ++i) {
            if (cnt[i] >= nsample) continue;  // warp-uniform
            int32_t *out = idx_all + ((size_t)b * m + (q0 + i)) * nsample;
            int c = cnt[i];
            for (int p0 = 0; p0 < tn && c < nsample; p0 += 32) {
                const int p = p0 + lane;
                const bool hit = (p < tn) && (dist_ref(qx[i], qy[i], qz[i], tx[p], ty[p], tz[p]) < r2);
                const uint32_t mask = __ballot_sync(0xffffffffu, hit);
                if (mask) {
                    if (c == 0) first[i] = base + p0 + __ffs(mask) - 1;
                    const int pos = c + __popc(mask & lt_mask);
                    if (hit && pos < nsample) out[pos] = base + p;
                    c += __popc(mask);
                }
            }
            cnt[i] = c;
        }
        bool done = true;
#pragma unroll
        for (int i = 0; i < BQ_QPW; ++i) done = done && (cnt[i] >= nsample);
        if (__syncthreads_and(done)) break;
    }
    // unused slots repeat the first hit (ball_query_gpu.cu:35-39); no hit -> zeros (pointnet2_utils.py:216)
#pragma unroll
    for (int i = 0; i < BQ_QPW; ++i) {
        if (q0 + i >= m) continue;
        int32_t *out = idx_all + ((size_t)b * m + (q0 + i)) * nsample;
        for (int l = cnt[i] + lane; l < nsample; l += 32) out[l] = first[i];
    }
}

}  // namespace
}  // namespace pn2

extern "C" int pn2_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                              int32_t *idx, void *stream) {
    using namespace pn2;
    PN2_REQUIRE(b >= 0 && n >= 1 && m >= 0 && nsample >= 0, "ball_query: bad dims b=%d n=%d m=%d nsample=%d", b, n, m, nsample);
    if (b == 0 || m == 0 || nsample == 0) return PN2_OK;
    PN2_REQUIRE(new_xyz && xyz && idx, "ball_query: null pointer");
    PN2_REQUIRE(b <= 65535, "ball_query: b exceeds the grid limit");
    dim3 grid(ceil_div(m, BQ_QPB), b);
    ball_query_kernel<<<grid, BQ_THREADS, 0, (cudaStream_t)stream>>>(n, m, radius, nsample, new_xyz, xyz, idx);
    PN2_LAUNCH_OK("ball_query");
    return PN2_OK;
}
