// gather_group.cu -- index gathers on the reference's channel-first layout, and their
// scatter-add backwards.  Replaces utils/src/sampling_gpu.cu:8-83 (gather_points[_grad]) and
// utils/src/group_points_gpu.cu:8-83 (group_points[_grad]).
//
// Both forwards are the same operation on a flattened index list:
//     out[b, c, j] = points[b, c, idx[b, j]],   j < J  (J = M, or npoints*nsample)
// The reference launches one thread per (b, c, j) and re-reads idx[b, j] for every channel.  Here a
// thread owns one j for a chunk of CH channels: idx is read once, the CH gathers are independent
// loads in flight together, and the stores of a warp stay coalesced along j.  Offsets are 64-bit
// (the reference's int32 index math overflows at 2^31 elements, group_points_gpu.cu:63).
#include "common.cuh"

namespace pn2 {
namespace {

constexpr int GG_THREADS = 256;
constexpr int GG_CH = 8;  // channels per thread

__global__ void __launch_bounds__(GG_THREADS)
gather_rows_kernel(int c, int n, long long J, const float *__restrict__ points, const int32_t *__restrict__ idx,
                   float *__restrict__ out) {
    const long long j = (long long)blockIdx.x * GG_THREADS + threadIdx.x;
    if (j >= J) return;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * GG_CH;
    const int k = idx[(size_t)b * J + j];
    const float *src = points + ((size_t)b * c + c0) * n + k;
    float *dst = out + ((size_t)b * c + c0) * J + j;
    const int cc = min(GG_CH, c - c0);
    float v[GG_CH];
#pragma unroll
    for (int i = 0; i < GG_CH; ++i)
        if (i < cc) v[i] = __ldg(src + (size_t)i * n);
#pragma unroll
    for (int i = 0; i < GG_CH; ++i)
        if (i < cc) __stcs(dst + (size_t)i * J, v[i]);
}

// grad_points[b, c, idx[b, j]] += grad_out[b, c, j]   (accumulation order unspecified, as in the reference)
__global__ void __launch_bounds__(GG_THREADS)
scatter_add_rows_kernel(int c, int n, long long J, const float *__restrict__ grad_out, const int32_t *__restrict__ idx,
                        float *__restrict__ grad_points) {
    const long long j = (long long)blockIdx.x * GG_THREADS + threadIdx.x;
    if (j >= J) return;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * GG_CH;
    const int k = idx[(size_t)b * J + j];
    const float *src = grad_out + ((size_t)b * c + c0) * J + j;
    float *dst = grad_points + ((size_t)b * c + c0) * n + k;
    const int cc = min(GG_CH, c - c0);
    float v[GG_CH];
#pragma unroll
    for (int i = 0; i < GG_CH; ++i)
        if (i < cc) v[i] = __ldcs(src + (size_t)i * J);
#pragma unroll
    for (int i = 0; i < GG_CH; ++i)
        if (i < cc) atomicAdd(dst + (size_t)i * n, v[i]);
}

// (B, R, Cc) -> (B, Cc, R) through a padded shared-memory tile.
__global__ void __launch_bounds__(256)
transpose_kernel(int r, int c, const float *__restrict__ in, float *__restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    in += (size_t)b * r * c;
    out += (size_t)b * r * c;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rr = r0 + ty + 8 * i, cc = c0 + tx;
        if (rr < r && cc < c) tile[ty + 8 * i][tx] = in[(size_t)rr * c + cc];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int cc = c0 + ty + 8 * i, rr = r0 + tx;
        if (rr < r && cc < c) out[(size_t)cc * r + rr] = tile[tx][ty + 8 * i];
    }
}

int check_common(const char *op, int b, int c, int n, long long J) {
    PN2_REQUIRE(b >= 0 && c >= 0 && n >= 1 && J >= 0, "%s: bad dims b=%d c=%d n=%d J=%lld", op, b, c, n, J);
    PN2_REQUIRE(b <= 65535 && ceil_div(c, GG_CH) <= 65535, "%s: b or c exceeds the grid limits", op);
    PN2_REQUIRE(J <= (long long)GG_THREADS * 2147483647ll, "%s: index list too long", op);
    return PN2_OK;
}

int gather_rows(const char *op, int b, int c, int n, long long J, const float *points, const int32_t *idx, float *out,
                void *stream) {
    if (int st = check_common(op, b, c, n, J)) return st;
    if (b == 0 || c == 0 || J == 0) return PN2_OK;
    PN2_REQUIRE(points && idx && out, "%s: null pointer", op);
    dim3 grid((unsigned)((J + GG_THREADS - 1) / GG_THREADS), ceil_div(c, GG_CH), b);
    gather_rows_kernel<<<grid, GG_THREADS, 0, (cudaStream_t)stream>>>(c, n, J, points, idx, out);
    PN2_LAUNCH_OK(op);
    return PN2_OK;
}

int scatter_rows(const char *op, int b, int c, int n, long long J, const float *grad_out, const int32_t *idx,
                 float *grad_points, void *stream) {
    if (int st = check_common(op, b, c, n, J)) return st;
    if (b == 0 || c == 0 || J == 0) return PN2_OK;
    PN2_REQUIRE(grad_out && idx && grad_points, "%s: null pointer", op);
    dim3 grid((unsigned)((J + GG_THREADS - 1) / GG_THREADS), ceil_div(c, GG_CH), b);
    scatter_add_rows_kernel<<<grid, GG_THREADS, 0, (cudaStream_t)stream>>>(c, n, J, grad_out, idx, grad_points);
    PN2_LAUNCH_OK(op);
    return PN2_OK;
}

}  // namespace
}  // namespace pn2

extern "C" int pn2_gather_points(int b, int c, int n, int npoints, const float *points, const int32_t *idx, float *out,
                                 void *stream) {
    return pn2::gather_rows("gather_points", b, c, n, npoints, points, idx, out, stream);
}

extern "C" int pn2_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int32_t *idx,
                                      float *grad_points, void *stream) {
    return pn2::scatter_rows("gather_points_grad", b, c, n, npoints, grad_out, idx, grad_points, stream);
}

extern "C" int pn2_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int32_t *idx,
                                float *out, void *stream) {
    PN2_REQUIRE(npoints >= 0 && nsample >= 0, "group_points: bad dims");
    return pn2::gather_rows("group_points", b, c, n, (long long)npoints * nsample, points, idx, out, stream);
}

extern "C" int pn2_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                     const int32_t *idx, float *grad_points, void *stream) {
    PN2_REQUIRE(npoints >= 0 && nsample >= 0, "group_points_grad: bad dims");
    return pn2::scatter_rows("group_points_grad", b, c, n, (long long)npoints * nsample, grad_out, idx, grad_points, stream);
}

extern "C" int pn2_transpose(int b, int r, int c, const float *in, float *out, void *stream) {
    using namespace pn2;
    PN2_REQUIRE(b >= 0 && r >= 0 && c >= 0, "transpose: bad dims");
    if (b == 0 || r == 0 || c == 0) return PN2_OK;
    PN2_REQUIRE(in && out, "transpose: null pointer");
    PN2_REQUIRE(b <= 65535 && ceil_div(c, 32) <= 65535, "transpose: dims exceed the grid limits");
    dim3 grid(ceil_div(r, 32), ceil_div(c, 32), b);
    transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(r, c, in, out);
    PN2_LAUNCH_OK("transpose");
    return PN2_OK;
}
