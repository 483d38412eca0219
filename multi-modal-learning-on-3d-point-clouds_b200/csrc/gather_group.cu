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

// Shared-memory variant for long index lists (grouping: J = npoints*nsample >> N).  A CTA stages CC whole source
// rows (CC * N floats, 16-byte asynchronous copies, all in flight) and then streams the index list once: every random
// access is an LDS, HBM sees only the compulsory bytes (source rows once, indices, coalesced output).
// The index loads are the only long-latency step of the streaming loop, and with few staged rows (long rows: N >= 16 k
// leaves room for one to three) there is little work per index to hide them behind: the loop keeps the NEXT iteration's
// U index vectors in flight while the current ones are expanded, and CTAs that stage at most two rows run 1024 threads.
template <int THREADS, int U>
__global__ void __launch_bounds__(THREADS, 1)
gather_rows_smem_kernel(int c, int n, long long J, int cc, const float *__restrict__ points, const int32_t *__restrict__ idx,
                        float *__restrict__ out) {
    extern __shared__ __align__(16) float rows[];  // [cc][n]
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * cc;
    const int nc = min(cc, c - c0);
    const float *src = points + ((size_t)b * c + c0) * n;
    const long long total = (long long)nc * n;
    if (((uintptr_t)src & 15) == 0 && (total & 3) == 0) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(rows);
        for (long long e = threadIdx.x; e < total / 4; e += THREADS)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)e * 16u), "l"(src + 4 * e) : "memory");
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
        for (long long e = threadIdx.x; e < total; e += THREADS) rows[e] = __ldcs(src + e);
    }
    __syncthreads();
    const int32_t *ib = idx + (size_t)b * J;
    float *ob = out + ((size_t)b * c + c0) * J;
    if ((J & 3) == 0 && ((uintptr_t)ib & 15) == 0 && ((uintptr_t)ob & 15) == 0) {
        // four consecutive outputs per thread and vector: one 128-bit index load, nc x (4 LDS + one 128-bit streaming
        // store); U vectors per iteration, the next iteration's U index loads issued before the current ones are used
        const long long J4 = J / 4;
        const long long per = (J4 + gridDim.x - 1) / gridDim.x;
        const long long q0 = (long long)blockIdx.x * per, q1 = min(J4, q0 + per);
        const int4 *ib4 = reinterpret_cast<const int4 *>(ib);
        int4 k[U], kn[U];
        long long q = q0 + threadIdx.x;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            k[u] = make_int4(0, 0, 0, 0);
            if (q + (long long)u * THREADS < q1) k[u] = __ldcs(ib4 + q + (long long)u * THREADS);
        }
        for (; q < q1; q += (long long)U * THREADS) {
            const long long qn = q + (long long)U * THREADS;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                kn[u] = make_int4(0, 0, 0, 0);
                if (qn + (long long)u * THREADS < q1) kn[u] = __ldcs(ib4 + qn + (long long)u * THREADS);
            }
            for (int i = 0; i < nc; ++i) {
                const float *r = rows + (size_t)i * n;
                float4 *dst = reinterpret_cast<float4 *>(ob + (size_t)i * J);
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (q + (long long)u * THREADS < q1)
                        __stcs(dst + q + (long long)u * THREADS, make_float4(r[k[u].x], r[k[u].y], r[k[u].z], r[k[u].w]));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) k[u] = kn[u];
        }
        return;
    }
    const long long per = (J + gridDim.x - 1) / gridDim.x;
    const long long j0 = (long long)blockIdx.x * per, j1 = min(J, j0 + per);
    for (long long j = j0 + threadIdx.x; j < j1; j += THREADS) {
        const int k = __ldcs(ib + j);
        for (int i = 0; i < nc; ++i) __stcs(ob + (size_t)i * J + j, rows[(size_t)i * n + k]);
    }
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

// Narrow transposes (the (B, 3, N) <-> (B, N, 3) conversions at the model boundary): one thread per long-axis element,
// the <= 8 short-axis values in registers; reads and writes are coalesced along the long axis (the 32 x 32 tile kernel
// would run 8192 CTAs with 3 of 32 rows active).
template <bool kRowsNarrow>
__global__ void __launch_bounds__(256)
transpose_narrow_kernel(int r, int c, const float *__restrict__ in, float *__restrict__ out) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * 256 + threadIdx.x;
    in += (size_t)b * r * c;
    out += (size_t)b * r * c;
    float v[8];
    if (kRowsNarrow) {  // (r <= 8, c) -> (c, r)
        if (j >= c) return;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < r) v[i] = in[(size_t)i * c + j];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < r) out[(size_t)j * r + i] = v[i];
    } else {  // (r, c <= 8) -> (c, r)
        if (j >= r) return;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < c) v[i] = in[(size_t)j * c + i];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < c) out[(size_t)i * r + j] = v[i];
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
    // long index lists over rows that fit in shared memory: stage the rows (see gather_rows_smem_kernel)
    const long long smem_floats = 50 * 1024;  // 200 KB
    if (J >= 2ll * n && n <= smem_floats && (long long)b * c >= 32) {
        int cc = (int)(smem_floats / n);
        if (cc > 8) cc = 8;
        if (cc > c) cc = c;
        const int chunks = ceil_div(c, cc);
        // enough CTAs for two waves; each CTA still streams a long slice of the index list
        int jsplit = 1;
        while ((long long)jsplit * chunks * b < 2ll * sm_count() && J / (jsplit * 2) >= 8 * n) jsplit *= 2;
        const size_t smem = (size_t)cc * n * sizeof(float);
        dim3 grid(jsplit, chunks, b);
        if (cc <= 2) {
            auto kern = gather_rows_smem_kernel<1024, 2>;
            PN2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, 1024, smem, (cudaStream_t)stream>>>(c, n, J, cc, points, idx, out);
        } else {
            auto kern = gather_rows_smem_kernel<512, 2>;
            PN2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, 512, smem, (cudaStream_t)stream>>>(c, n, J, cc, points, idx, out);
        }
        PN2_LAUNCH_OK(op);
        return PN2_OK;
    }
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
    if (r <= 8 && c >= 256) {
        transpose_narrow_kernel<true><<<dim3(ceil_div(c, 256), b), 256, 0, (cudaStream_t)stream>>>(r, c, in, out);
        PN2_LAUNCH_OK("transpose_narrow");
        return PN2_OK;
    }
    if (c <= 8 && r >= 256) {
        transpose_narrow_kernel<false><<<dim3(ceil_div(r, 256), b), 256, 0, (cudaStream_t)stream>>>(r, c, in, out);
        PN2_LAUNCH_OK("transpose_narrow");
        return PN2_OK;
    }
    dim3 grid(ceil_div(r, 32), ceil_div(c, 32), b);
    transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(r, c, in, out);
    PN2_LAUNCH_OK("transpose");
    return PN2_OK;
}
