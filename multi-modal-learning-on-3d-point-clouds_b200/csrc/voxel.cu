// voxel.cu -- evaluation voxelisation on the device.
//
// Replaces utils/pc_util.py:39-51 point_cloud_label_to_surface_voxel_label_fast as the evaluation loops call it once per
// scene on the host (train_scannet_semseg.py:226-227, train_scannet_multiview_semseg.py:271-272): over the points of a
// cloud whose sample weight is > 0,
//     nvox = ceil((max - min) / res),  v = ceil((p - min) / res)            (fp32, numpy's float32 arithmetic)
//     vidx = v0 + v1 * nvox0 + v2 * nvox0 * nvox1                           (fp32, left to right)
//     uvidx, first = numpy.unique(vidx, return_index=True)                  (ascending vidx, FIRST point of each voxel)
// The labels of a voxel are those of its first point, so with `first` the per-voxel metrics need no host round trip.
// One CTA per cloud computes the bounds and the keys; a stable device radix sort (cub, plumbing) orders (cloud, vidx)
// keeping ascending point index inside a voxel; one CTA per cloud compacts the segment heads.
#include <cub/cub.cuh>

#include "common.cuh"

namespace pn2 {
namespace {

constexpr int VX_THREADS = 256;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o >= 1; o >>= 1) {
        const float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, u) : fminf(v, u);
    }
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int w = 1; w < VX_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : fminf(r, red[w]);
    return r;
}

// grid = b; keys[b*n + i] = (cloud << 32) | float bits of vidx (non-negative floats order like their bits);
// masked-out points get 0xFFFFFFFF (sorts last inside the cloud, never a head)
__global__ void __launch_bounds__(VX_THREADS)
voxel_keys_kernel(int n, const float *__restrict__ xyz_all, const unsigned char *__restrict__ mask_all, float res,
                  unsigned long long *__restrict__ keys, int32_t *__restrict__ vals, float *__restrict__ nvox_out) {
    __shared__ float red[VX_THREADS / 32];
    const int b = blockIdx.x;
    const float *xyz = xyz_all + (size_t)b * n * 3;
    const unsigned char *mask = mask_all ? mask_all + (size_t)b * n : nullptr;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < n; i += VX_THREADS)
        if (!mask || mask[i]) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v = xyz[3 * i + a];
                lo[a] = fminf(lo[a], v);
                hi[a] = fmaxf(hi[a], v);
            }
        }
    float nv[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = block_reduce(lo[a], false, red);
        hi[a] = block_reduce(hi[a], true, red);
        nv[a] = ceilf(__fdiv_rn(__fsub_rn(hi[a], lo[a]), res));
    }
    if (threadIdx.x < 3 && nvox_out) nvox_out[3 * b + threadIdx.x] = nv[threadIdx.x];
    for (int i = threadIdx.x; i < n; i += VX_THREADS) {
        uint32_t bits = 0xFFFFFFFFu;
        if (!mask || mask[i]) {
            const float v0 = ceilf(__fdiv_rn(__fsub_rn(xyz[3 * i + 0], lo[0]), res));
            const float v1 = ceilf(__fdiv_rn(__fsub_rn(xyz[3 * i + 1], lo[1]), res));
            const float v2 = ceilf(__fdiv_rn(__fsub_rn(xyz[3 * i + 2], lo[2]), res));
            // (v0 + v1*nvox0) + (v2*nvox0)*nvox1, one rounding per operation as numpy evaluates it
            const float vidx = __fadd_rn(__fadd_rn(v0, __fmul_rn(v1, nv[0])), __fmul_rn(__fmul_rn(v2, nv[0]), nv[1]));
            bits = __float_as_uint(vidx);
        }
        keys[(size_t)b * n + i] = ((unsigned long long)b << 32) | bits;
        vals[(size_t)b * n + i] = i;
    }
}

// grid = b; compacts the heads of the sorted segments of cloud b into uvidx / first, pads the rest with -1
__global__ void __launch_bounds__(VX_THREADS)
voxel_heads_kernel(int n, const unsigned long long *__restrict__ keys, const int32_t *__restrict__ vals, float *__restrict__ uvidx,
                   int32_t *__restrict__ first, int32_t *__restrict__ count) {
    __shared__ int warp_sum[VX_THREADS / 32];
    __shared__ int base_s;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long *k = keys + (size_t)b * n;
    const int32_t *v = vals + (size_t)b * n;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += VX_THREADS) {
        const int i = i0 + threadIdx.x;
        uint32_t bits = 0xFFFFFFFFu;
        bool head = false;
        if (i < n) {
            bits = (uint32_t)k[i];
            head = bits != 0xFFFFFFFFu && (i == 0 || (uint32_t)k[i - 1] != bits);
        }
        const unsigned m = __ballot_sync(0xffffffffu, head);
        if (lane == 0) warp_sum[warp] = __popc(m);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < warp; ++w) off += warp_sum[w];
        if (head) {
            const int dst = off + __popc(m & ((1u << lane) - 1u));
            uvidx[(size_t)b * n + dst] = __uint_as_float(bits);
            first[(size_t)b * n + dst] = v[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < VX_THREADS / 32; ++w) t += warp_sum[w];
            base_s += t;
        }
        __syncthreads();
    }
    const int total = base_s;
    if (threadIdx.x == 0) count[b] = total;
    for (int i = total + threadIdx.x; i < n; i += VX_THREADS) {
        uvidx[(size_t)b * n + i] = -1.f;
        first[(size_t)b * n + i] = -1;
    }
}

// Per-class confusion counters of the evaluation loop (train_scannet_semseg.py:218-223 point-wise, :232-239 voxel-wise):
// out[0][l] += #(target == l), out[1][l] += #(target == l and pred == l), out[2][l] += #(target == l or pred == l) over
// the selected points.  Integer atomics: the result does not depend on the order.
__global__ void label_counts_kernel(int n, int num_classes, const int32_t *__restrict__ select, const int32_t *__restrict__ count,
                                    const unsigned char *__restrict__ mask, const long long *__restrict__ target,
                                    const unsigned char *__restrict__ pred, unsigned long long *__restrict__ out) {
    extern __shared__ unsigned int hist[];  // [3][num_classes]
    for (int i = threadIdx.x; i < 3 * num_classes; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const int b = blockIdx.y;
    const int limit = count ? count[b] : n;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < limit; s += gridDim.x * blockDim.x) {
        const int i = select ? select[(size_t)b * n + s] : s;
        if (mask && !mask[(size_t)b * n + i]) continue;
        const long long t = target[(size_t)b * n + i];
        const int p = pred[(size_t)b * n + i];
        if (t >= 0 && t < num_classes) {
            atomicAdd(&hist[t], 1u);
            if (p == t) atomicAdd(&hist[num_classes + t], 1u);
            atomicAdd(&hist[2 * num_classes + t], 1u);
        }
        if (p != t && p < num_classes) atomicAdd(&hist[2 * num_classes + p], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * num_classes; i += blockDim.x)
        if (hist[i]) atomicAdd(&out[i], (unsigned long long)hist[i]);
}

}  // namespace
}  // namespace pn2

extern "C" int pn2_label_counts(int b, int n, int num_classes, const int32_t *select, const int32_t *count, const unsigned char *mask,
                                const long long *target, const unsigned char *pred, long long *out, void *stream) {
    using namespace pn2;
    PN2_REQUIRE(b >= 0 && n >= 0 && num_classes >= 1 && num_classes <= 256, "label_counts: bad arguments b=%d n=%d classes=%d", b, n, num_classes);
    if (b == 0 || n == 0) return PN2_OK;
    PN2_REQUIRE(target && pred && out && (!select || count), "label_counts: null pointer (select needs count)");
    PN2_REQUIRE(b <= 65535, "label_counts: b must be <= 65535");
    dim3 grid((unsigned)((n + 1023) / 1024), (unsigned)b);
    label_counts_kernel<<<grid, 256, 3 * num_classes * sizeof(unsigned int), (cudaStream_t)stream>>>(
        n, num_classes, select, count, mask, target, pred, reinterpret_cast<unsigned long long *>(out));
    PN2_LAUNCH_OK("label_counts_kernel");
    return PN2_OK;
}

extern "C" int pn2_voxel_first_index(int b, int n, const float *xyz, const unsigned char *mask, float res, float *uvidx, int32_t *first,
                                     int32_t *count, float *nvox, void *stream) {
    using namespace pn2;
    PN2_REQUIRE(b >= 0 && n >= 0 && res > 0.f, "voxel_first_index: bad arguments b=%d n=%d res=%g", b, n, (double)res);
    PN2_REQUIRE((long long)b * n < (1ll << 31), "voxel_first_index: more than 2^31 points");
    if (b == 0) return PN2_OK;
    PN2_REQUIRE(count && (n == 0 || (xyz && uvidx && first)), "voxel_first_index: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        PN2_CUDA(cudaMemsetAsync(count, 0, (size_t)b * sizeof(int32_t), s));
        return PN2_OK;
    }
    const long long total = (long long)b * n;
    int end_bit = 33;
    while ((1ll << (end_bit - 32)) < b) ++end_bit;
    size_t tmp_bytes = 0;
    PN2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                             (const int32_t *)nullptr, (int32_t *)nullptr, (int)total, 0, end_bit, s));
    const size_t kb = ((size_t)total * 8 + 255) / 256 * 256, vb = ((size_t)total * 4 + 255) / 256 * 256;
    Scratch scratch_mem(s);  // [keys in | keys out | vals in | vals out | cub temp], released on every return below
    PN2_CUDA(scratch_mem.alloc(2 * kb + 2 * vb + tmp_bytes));
    unsigned char *scratch = (unsigned char *)scratch_mem.ptr;
    unsigned long long *k_in = (unsigned long long *)scratch, *k_out = (unsigned long long *)(scratch + kb);
    int32_t *v_in = (int32_t *)(scratch + 2 * kb), *v_out = (int32_t *)(scratch + 2 * kb + vb);
    voxel_keys_kernel<<<b, VX_THREADS, 0, s>>>(n, xyz, mask, res, k_in, v_in, nvox);
    PN2_LAUNCH_OK("voxel_keys_kernel");
    // stable: inside a voxel the point indices stay ascending, so the head of a segment is numpy.unique's return_index
    PN2_CUDA(cub::DeviceRadixSort::SortPairs(scratch + 2 * kb + 2 * vb, tmp_bytes, k_in, k_out, v_in, v_out, (int)total, 0, end_bit, s));
    voxel_heads_kernel<<<b, VX_THREADS, 0, s>>>(n, k_out, v_out, uvidx, first, count);
    PN2_LAUNCH_OK("voxel_heads_kernel");
    return PN2_OK;
}
