// grid.cu -- exact neighbour search through a per-cloud uniform grid (cell list).
//
// The reference's ball query and 3-NN are brute force: every query scans all N points
// (utils/src/ball_query_gpu.cu:28-44, interpolate_gpu.cu:30-49).  Their RESULTS are what parity is judged on
// (first nsample hits in ascending index order; the three smallest (distance, index) pairs), not the scan.
// Here each cloud is sorted once by grid cell (one CTA per cloud, bitonic sort in shared memory) and a query only
// tests the points of the cells its ball can touch; candidate distances use the same fp32 rounding sequence
// (dist_ref), candidate cells are chosen with a safety margin far above fp32 rounding error, and the ordering
// rules are re-established explicitly (rank by index / lexicographic (d, idx) top-3), so the outputs are
// bit-identical to the brute-force scan.  Queries whose neighbourhood is too dense for the fixed-size buffers
// fall back to the ordered brute-force scan inside the same kernel.
//
// The sorted order doubles as a spatially coherent row order for the feature-propagation gather (row_mlp_tc.cu).
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

namespace pn2 {
namespace {

constexpr int GRID_MAX_N = 32768;       // points per cloud the single-CTA sort handles (128 KB of 32-bit keys)
constexpr int GRID_MAX_CELLS = 1 << 16; // cells per cloud (dense start table)
constexpr int GRID_MAX_DIM = 1024;

struct GridMeta {  // 8 words per cloud
    float ox, oy, oz, inv_h;
    int dx, dy, dz, ncells;
};

__device__ __forceinline__ int cell_coord(float p, float o, float inv_h, int dim) {
    // monotone non-decreasing in p (fp32 sub / mul are monotone), clamped to the grid
    float v = (p - o) * inv_h;
    v = fminf(fmaxf(v, 0.f), (float)(dim - 1));
    return (int)v;
}

// Cell size and grid dimensions from the bounding box of a cloud.
__device__ GridMeta make_meta(const float (&lo)[3], const float (&hi)[3], float h, int n) {
    // h <= 0: pick the cell from the bounding box so that a cell holds about one point
    float hh = h;
    if (!(hh > 0.f)) {
        float e0 = hi[0] - lo[0], e1 = hi[1] - lo[1], e2 = hi[2] - lo[2];
        const float emax = fmaxf(e0, fmaxf(e1, e2)), emin = fminf(e0, fminf(e1, e2));
        const float emid = e0 + e1 + e2 - emax - emin;
        if (emin < 0.05f * emax)
            hh = sqrtf(fmaxf(emax * fmaxf(emid, 1e-6f * emax), 1e-30f) / (float)n);  // (nearly) planar cloud
        else
            hh = cbrtf(e0 * e1 * e2 / (float)n);
        hh = fmaxf(hh, 1e-6f);
    }
    // grow the cell until the dense table fits; a larger cell only adds candidates, never loses one
    int d[3];
    for (int iter = 0; iter < 64; ++iter) {
        const float inv = 1.0f / hh;
        long long cells = 1;
        for (int a = 0; a < 3; ++a) {
            float e = (hi[a] - lo[a]) * inv;
            e = fminf(e, (float)(GRID_MAX_DIM - 1));
            d[a] = (int)e + 1;
            cells *= d[a];
        }
        if (cells <= GRID_MAX_CELLS) break;
        hh *= 1.26f;
    }
    GridMeta m;
    m.ox = lo[0]; m.oy = lo[1]; m.oz = lo[2];
    m.inv_h = 1.0f / hh;
    m.dx = d[0]; m.dy = d[1]; m.dz = d[2];
    m.ncells = d[0] * d[1] * d[2];
    return m;
}

// One CTA per cloud: bounding box, cell of every point, bitonic sort of (cell, index), dense cell-start table.
__global__ void __launch_bounds__(1024, 1)
grid_build_kernel(int n, float h, const float *__restrict__ xyz_all, float4 *__restrict__ sorted_all,
                  int32_t *__restrict__ cell_start_all, int32_t *__restrict__ order_all, GridMeta *__restrict__ meta_all) {
    extern __shared__ uint32_t keys[];  // np2 entries: (cell << 16) | index  (cells and indices both fit 16 bits)
    __shared__ float red[6][32];
    __shared__ GridMeta sm;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *xyz = xyz_all + (size_t)b * n * 3;
    int np2 = 1;
    while (np2 < n) np2 <<= 1;

    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int k = t; k < n; k += 1024) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = xyz[3 * k + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if (lane == 0) {
            red[a][warp] = mn[a];
            red[3 + a][warp] = mx[a];
        }
    }
    __syncthreads();
    if (t == 0) {
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            lo[a] = red[a][0];
            hi[a] = red[3 + a][0];
            for (int w = 1; w < 32; ++w) {
                lo[a] = fminf(lo[a], red[a][w]);
                hi[a] = fmaxf(hi[a], red[3 + a][w]);
            }
        }
        sm = make_meta(lo, hi, h, n);
        meta_all[b] = sm;
    }
    __syncthreads();
    const GridMeta g = sm;
    for (int k = t; k < np2; k += 1024) {
        uint32_t key = 0xffffffffu;
        if (k < n) {
            const int cx = cell_coord(xyz[3 * k], g.ox, g.inv_h, g.dx);
            const int cy = cell_coord(xyz[3 * k + 1], g.oy, g.inv_h, g.dy);
            const int cz = cell_coord(xyz[3 * k + 2], g.oz, g.inv_h, g.dz);
            const unsigned cell = (unsigned)(cx + g.dx * (cy + g.dy * cz));
            key = (cell << 16) | (unsigned)k;
        }
        keys[k] = key;
    }
    __syncthreads();
    for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = t; i < np2 / 2; i += 1024) {
                const int lo_i = 2 * i - (i & (stride - 1));   // insert a zero bit at position log2(stride)
                const int hi_i = lo_i + stride;
                const bool up = (lo_i & size) == 0;
                const uint32_t a = keys[lo_i], c = keys[hi_i];
                if ((a > c) == up) {
                    keys[lo_i] = c;
                    keys[hi_i] = a;
                }
            }
            __syncthreads();
        }
    }
    float4 *sorted = sorted_all + (size_t)b * n;
    int32_t *order = order_all ? order_all + (size_t)b * n : nullptr;
    int32_t *cell_start = cell_start_all + (size_t)b * (GRID_MAX_CELLS + 1);
    for (int i = t; i <= n; i += 1024) {
        const int c_prev = i > 0 ? (int)(keys[i - 1] >> 16) : -1;
        const int c_cur = i < n ? (int)(keys[i] >> 16) : g.ncells;
        for (int c = c_prev + 1; c <= c_cur; ++c) cell_start[c] = i;
        if (i < n) {
            const int k = (int)(keys[i] & 0xffffu);
            sorted[i] = make_float4(xyz[3 * k], xyz[3 * k + 1], xyz[3 * k + 2], __int_as_float(k));
            if (order) order[i] = k;
        }
    }
}

// ---- clouds beyond the single-CTA sort (GRID_MAX_N < n <= GRID_MAX_N_LARGE): counting sort by cell -----------------
// bounding box + meta (one CTA per cloud), per-cell counts (atomics), exclusive scan of the <= 65536 counts (one CTA per
// cloud), scatter through per-cell cursors.  The order of the points INSIDE a cell is whatever the atomics produce; the
// query kernels re-establish the reference's ordering explicitly (rank by index, lexicographic (d, idx)), so their
// results do not depend on it.
constexpr int GRID_MAX_N_LARGE = 1 << 18;

__global__ void __launch_bounds__(1024, 1)
grid_meta_kernel(int n, float h, const float *__restrict__ xyz_all, GridMeta *__restrict__ meta_all) {
    __shared__ float red[6][32];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *xyz = xyz_all + (size_t)b * n * 3;
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int k = t; k < n; k += 1024) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = xyz[3 * k + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if (lane == 0) {
            red[a][warp] = mn[a];
            red[3 + a][warp] = mx[a];
        }
    }
    __syncthreads();
    if (t == 0) {
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            lo[a] = red[a][0];
            hi[a] = red[3 + a][0];
            for (int w = 1; w < 32; ++w) {
                lo[a] = fminf(lo[a], red[a][w]);
                hi[a] = fmaxf(hi[a], red[3 + a][w]);
            }
        }
        meta_all[b] = make_meta(lo, hi, h, n);
    }
}

__device__ __forceinline__ int cell_of(const GridMeta &g, const float *p) {
    const int cx = cell_coord(p[0], g.ox, g.inv_h, g.dx);
    const int cy = cell_coord(p[1], g.oy, g.inv_h, g.dy);
    const int cz = cell_coord(p[2], g.oz, g.inv_h, g.dz);
    return cx + g.dx * (cy + g.dy * cz);
}

__global__ void __launch_bounds__(256)
grid_count_kernel(int n, const float *__restrict__ xyz_all, const GridMeta *__restrict__ meta_all, int32_t *__restrict__ cursor_all) {
    const int b = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    const GridMeta g = meta_all[b];
    atomicAdd(cursor_all + (size_t)b * (GRID_MAX_CELLS + 1) + cell_of(g, xyz_all + ((size_t)b * n + k) * 3), 1);
}

// exclusive scan of the per-cell counts -> cell_start[0..ncells]; the cursors restart at the cell starts
__global__ void __launch_bounds__(1024, 1)
grid_scan_kernel(int n, const GridMeta *__restrict__ meta_all, int32_t *__restrict__ cursor_all, int32_t *__restrict__ cell_start_all) {
    __shared__ int wsum[32];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int ncells = meta_all[b].ncells;
    int32_t *cursor = cursor_all + (size_t)b * (GRID_MAX_CELLS + 1);
    int32_t *cell_start = cell_start_all + (size_t)b * (GRID_MAX_CELLS + 1);
    const int per = (ncells + 1023) / 1024;
    const int c0 = t * per, c1 = min(ncells, c0 + per);
    int local = 0;
    for (int c = c0; c < c1; ++c) local += cursor[c];
    int incl = local;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = wsum[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        wsum[lane] = v;
    }
    __syncthreads();
    int run = incl - local + (warp > 0 ? wsum[warp - 1] : 0);
    for (int c = c0; c < c1; ++c) {
        const int cnt = cursor[c];
        cell_start[c] = run;
        cursor[c] = run;
        run += cnt;
    }
    if (t == 0) cell_start[ncells] = n;
}

__global__ void __launch_bounds__(256)
grid_scatter_kernel(int n, const float *__restrict__ xyz_all, const GridMeta *__restrict__ meta_all, int32_t *__restrict__ cursor_all,
                    float4 *__restrict__ sorted_all, int32_t *__restrict__ order_all) {
    const int b = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
    if (k >= n) return;
    const GridMeta g = meta_all[b];
    const float *p = xyz_all + ((size_t)b * n + k) * 3;
    const int pos = atomicAdd(cursor_all + (size_t)b * (GRID_MAX_CELLS + 1) + cell_of(g, p), 1);
    sorted_all[(size_t)b * n + pos] = make_float4(p[0], p[1], p[2], __int_as_float(k));
    if (order_all) order_all[(size_t)b * n + pos] = k;
}

// ---- ball query ------------------------------------------------------------------------------------------
constexpr int BQG_WARPS = 8;
constexpr int BQG_CAP = 128;   // hits buffered per query before falling back to the ordered scan
constexpr int BQG_MAX_ROWS = 16;

__global__ void __launch_bounds__(BQG_WARPS * 32)
ball_query_grid_kernel(int n, int m, float radius, int nsample, const float *__restrict__ new_xyz_all,
                       const float *__restrict__ xyz_all, const float4 *__restrict__ sorted_all,
                       const int32_t *__restrict__ cell_start_all, const GridMeta *__restrict__ meta_all,
                       int32_t *__restrict__ idx_all) {
    __shared__ int hits[BQG_WARPS][BQG_CAP];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * BQG_WARPS + warp;
    if (q >= m) return;
    const GridMeta g = meta_all[b];
    const float4 *sorted = sorted_all + (size_t)b * n;
    const int32_t *cell_start = cell_start_all + (size_t)b * (GRID_MAX_CELLS + 1);
    const float *c = new_xyz_all + ((size_t)b * m + q) * 3;
    const float qx = c[0], qy = c[1], qz = c[2];
    const float r2 = __fmul_rn(radius, radius);
    const float rm = fabsf(radius) * 1.001f + 1e-30f;  // candidate margin >> fp32 rounding of the distance
    int32_t *out = idx_all + ((size_t)b * m + q) * nsample;
    int *hb = hits[warp];
    const uint32_t lt_mask = (1u << lane) - 1u;

    const int cx0 = cell_coord(qx - rm, g.ox, g.inv_h, g.dx), cx1 = cell_coord(qx + rm, g.ox, g.inv_h, g.dx);
    const int cy0 = cell_coord(qy - rm, g.oy, g.inv_h, g.dy), cy1 = cell_coord(qy + rm, g.oy, g.inv_h, g.dy);
    const int cz0 = cell_coord(qz - rm, g.oz, g.inv_h, g.dz), cz1 = cell_coord(qz + rm, g.oz, g.inv_h, g.dz);
    const int ny = cy1 - cy0 + 1, nrows = ny * (cz1 - cz0 + 1);
    int cnt = 0;
    bool overflow = nrows > BQG_MAX_ROWS || !(radius == radius);
    if (!overflow) {
        // lane l < nrows owns row l: the cells (cx0..cx1, cy, cz) are one contiguous run of the sorted cloud
        int rs = 0, re = 0;
        if (lane < nrows) {
            const int cy = cy0 + lane % ny, cz = cz0 + lane / ny;
            const int base = g.dx * (cy + g.dy * cz);
            rs = __ldg(cell_start + base + cx0);
            re = __ldg(cell_start + base + cx1 + 1);
        }
        for (int row = 0; row < nrows && !overflow; ++row) {
            const int s = __shfl_sync(0xffffffffu, rs, row), e = __shfl_sync(0xffffffffu, re, row);
            for (int p0 = s; p0 < e; p0 += 32) {
                const int p = p0 + lane;
                bool hit = false;
                int k = 0;
                if (p < e) {
                    const float4 v = __ldg(sorted + p);
                    k = __float_as_int(v.w);
                    hit = dist_ref(qx, qy, qz, v.x, v.y, v.z) < r2;
                }
                const uint32_t mask = __ballot_sync(0xffffffffu, hit);
                const int pos = cnt + __popc(mask & lt_mask);
                if (hit && pos < BQG_CAP) hb[pos] = k;
                cnt += __popc(mask);
            }
            if (cnt > BQG_CAP) overflow = true;
        }
    }
    if (overflow) {
        // dense neighbourhood (or oversized ball): ordered scan of the original cloud with early exit, as the reference
        const float *xyz = xyz_all + (size_t)b * n * 3;
        int c2 = 0, first = 0;
        for (int p0 = 0; p0 < n && c2 < nsample; p0 += 32) {
            const int p = p0 + lane;
            const bool hit = (p < n) && (dist_ref(qx, qy, qz, __ldg(xyz + 3 * p), __ldg(xyz + 3 * p + 1), __ldg(xyz + 3 * p + 2)) < r2);
            const uint32_t mask = __ballot_sync(0xffffffffu, hit);
            if (mask) {
                if (c2 == 0) first = p0 + __ffs(mask) - 1;
                const int pos = c2 + __popc(mask & lt_mask);
                if (hit && pos < nsample) out[pos] = p;
                c2 += __popc(mask);
            }
        }
        for (int l = c2 + lane; l < nsample; l += 32) out[l] = first;
        return;
    }
    __syncwarp();
    // the first nsample hits in ascending index order = the nsample smallest indices: rank by counting
    int first = 0x7fffffff;
    for (int i = lane; i < cnt; i += 32) {
        const int mine = hb[i];
        int rank = 0;
        for (int j = 0; j < cnt; ++j) rank += (hb[j] < mine);
        if (rank < nsample) out[rank] = mine;
        first = min(first, mine);
    }
    for (int o = 16; o; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    if (cnt == 0) first = 0;
    for (int l = cnt + lane; l < nsample; l += 32) out[l] = first;
}

// ---- three nearest neighbours (+ the reference's interpolation weights) --------------------------------------
// The running best three are 64-bit keys (float bits of the squared distance << 32 | index): distances are >= 0, so the
// unsigned order of the keys is the lexicographic (d, index) order, and keeping the three smallest keys gives the same
// final state as the reference's ascending strict-'<' scan whatever order the candidates arrive in.  The insertion is a
// branch-free min / max network: with 32 queries per warp nearly every candidate is inserted by SOME lane, so a branchy
// insertion runs all of its nested paths for almost every candidate (ncu: 18 of 32 lanes active on average, 2700 warp
// instructions per 32 queries at the fp1 shape).
struct Key3 {
    unsigned long long k1, k2, k3;
};

constexpr unsigned long long KEY_EMPTY = (0x7f800000ull << 32) | 0x7fffffffull;  // (+inf, no index)

__device__ __forceinline__ void key3_insert(Key3 &t, float d, int k) {
    const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)k;
    const unsigned long long a = min(t.k1, key), r1 = max(t.k1, key);
    const unsigned long long b = min(t.k2, r1), r2 = max(t.k2, r1);
    t.k1 = a;
    t.k2 = b;
    t.k3 = min(t.k3, r2);
}

// One query of the 3-NN search; `sorted` / `cell_start` point to the cloud's cell list in global OR shared memory
// (three_nn_grid_staged_kernel copies small coarse clouds there first).  Lanes past the end repeat the last query and
// store nothing.
// (Measured and not kept: a first pass in which all lanes of a warp scan the ONE block that holds every lane's ring-1
// neighbourhood -- uniform trip counts, broadcast loads.  With the queries in the row-major cell order of the fine grid a
// warp's 32 queries span ~6 coarse cells, the common block holds 2-3 x the candidates and the pass is slower at every
// size: fp1 shape 43 us without it, 45-50 us with block limits of 27-80 cells.)
__device__ __forceinline__ void three_nn_grid_query(int n, int m, int b, int slot, bool valid, const GridMeta &g,
                                                    const float *__restrict__ unknown_all, const float *__restrict__ known_all,
                                                    const float4 *sorted, const int32_t *cell_start,
                                                    const int32_t *__restrict__ query_order, float *__restrict__ dist2_out,
                                                    int32_t *__restrict__ idx_out, float *__restrict__ weight_out) {
    // optional spatially coherent processing order (neighbouring threads then walk the same cells)
    const int i = query_order ? __ldg(query_order + (size_t)b * n + slot) : slot;
    const float *u = unknown_all + ((size_t)b * n + i) * 3;
    const float ux = u[0], uy = u[1], uz = u[2];
    const float h = 1.0f / g.inv_h;
    const float INF = __int_as_float(0x7f800000);
    Key3 t;
    const int cx = cell_coord(ux, g.ox, g.inv_h, g.dx), cy = cell_coord(uy, g.oy, g.inv_h, g.dy),
              cz = cell_coord(uz, g.oz, g.inv_h, g.dz);
    bool done = false;
    // rings 1, 2, then doubling while a block is still much cheaper than the brute-force scan (row look-ups + points
    // grow with the block; in flat or very non-uniform clouds -- lidar sweeps -- the third neighbour is often further
    // than two cells away, and a 9 x 9 x dz block costs a few dozen row look-ups against m distance tests)
    for (int ring = 1; !done; ring = ring < 2 ? 2 : 2 * ring) {
        const long long rows = (long long)min(2 * ring + 1, g.dy) * min(2 * ring + 1, g.dz);
        if (ring > 2 && (rows * 8 > m || ring > 64)) break;
        t.k1 = t.k2 = t.k3 = KEY_EMPTY;
        const int x0 = max(cx - ring, 0), x1 = min(cx + ring, g.dx - 1);
        const int y0 = max(cy - ring, 0), y1 = min(cy + ring, g.dy - 1);
        const int z0 = max(cz - ring, 0), z1 = min(cz + ring, g.dz - 1);
        for (int zc = z0; zc <= z1; ++zc)
            for (int yc = y0; yc <= y1; ++yc) {
                const int base = g.dx * (yc + g.dy * zc);
                const int s = cell_start[base + x0], e = cell_start[base + x1 + 1];
                for (int p = s; p < e; ++p) {
                    const float4 v = sorted[p];
                    key3_insert(t, dist_ref(ux, uy, uz, v.x, v.y, v.z), __float_as_int(v.w));
                }
            }
        // Every known point outside the scanned block is at least `rho` away: the distance from the query to the
        // nearest face of the block that still has grid beyond it (faces on the grid boundary bound nothing:
        // all known points lie inside the grid).
        float rho = INF;
        if (cx - ring > 0) rho = fminf(rho, ux - (g.ox + (float)(cx - ring) * h));
        if (cx + ring < g.dx - 1) rho = fminf(rho, (g.ox + (float)(cx + ring + 1) * h) - ux);
        if (cy - ring > 0) rho = fminf(rho, uy - (g.oy + (float)(cy - ring) * h));
        if (cy + ring < g.dy - 1) rho = fminf(rho, (g.oy + (float)(cy + ring + 1) * h) - uy);
        if (cz - ring > 0) rho = fminf(rho, uz - (g.oz + (float)(cz - ring) * h));
        if (cz + ring < g.dz - 1) rho = fminf(rho, (g.oz + (float)(cz + ring + 1) * h) - uz);
        if (rho == INF) {
            done = true;  // the block covers the whole grid
        } else if (rho > 0.f) {
            const float safe = rho * 0.999f;  // margin >> fp32 rounding of cell edges and distances
            done = __uint_as_float((unsigned)(t.k3 >> 32)) < safe * safe;
        }
    }
    if (!done) {
        // sparse neighbourhood: exact scan of the whole known set
        const float *known = known_all + (size_t)b * m * 3;
        t.k1 = t.k2 = t.k3 = KEY_EMPTY;
        for (int k = 0; k < m; ++k)
            key3_insert(t, dist_ref(ux, uy, uz, __ldg(known + 3 * k), __ldg(known + 3 * k + 1), __ldg(known + 3 * k + 2)), k);
    }
    if (!valid) return;
    const float d1 = __uint_as_float((unsigned)(t.k1 >> 32)), d2 = __uint_as_float((unsigned)(t.k2 >> 32)),
                d3 = __uint_as_float((unsigned)(t.k3 >> 32));
    // fewer than three known points: index 0 / +inf as the reference leaves them
    const int i1 = d1 == INF ? 0 : (int)(unsigned)t.k1, i2 = d2 == INF ? 0 : (int)(unsigned)t.k2, i3 = d3 == INF ? 0 : (int)(unsigned)t.k3;
    const size_t o = ((size_t)b * n + i) * 3;
    idx_out[o] = i1; idx_out[o + 1] = i2; idx_out[o + 2] = i3;
    if (dist2_out) {
        dist2_out[o] = d1; dist2_out[o + 1] = d2; dist2_out[o + 2] = d3;
    }
    if (weight_out) {
        // model/pointnet2_utils.py:97 + model/pointnet_util.py:206-208
        float e1 = __fsqrt_rn(d1), e2 = __fsqrt_rn(d2), e3 = __fsqrt_rn(d3);
        e1 = e1 < 1e-10f ? 1e-10f : e1;
        e2 = e2 < 1e-10f ? 1e-10f : e2;
        e3 = e3 < 1e-10f ? 1e-10f : e3;
        const float w1 = __fdiv_rn(1.0f, e1), w2 = __fdiv_rn(1.0f, e2), w3 = __fdiv_rn(1.0f, e3);
        const float s = __fadd_rn(__fadd_rn(w1, w2), w3);
        weight_out[o] = __fdiv_rn(w1, s); weight_out[o + 1] = __fdiv_rn(w2, s); weight_out[o + 2] = __fdiv_rn(w3, s);
    }
}


__global__ void __launch_bounds__(128)
three_nn_grid_kernel(int n, int m, const float *__restrict__ unknown_all, const float *__restrict__ known_all,
                     const float4 *__restrict__ sorted_all, const int32_t *__restrict__ cell_start_all,
                     const GridMeta *__restrict__ meta_all, const int32_t *__restrict__ query_order,
                     float *__restrict__ dist2_out, int32_t *__restrict__ idx_out, float *__restrict__ weight_out) {
    const int b = blockIdx.y;
    const int slot = blockIdx.x * 128 + threadIdx.x;
    const GridMeta g = meta_all[b];
    three_nn_grid_query(n, m, b, min(slot, n - 1), slot < n, g, unknown_all, known_all, sorted_all + (size_t)b * m,
                        cell_start_all + (size_t)b * (GRID_MAX_CELLS + 1), query_order, dist2_out, idx_out, weight_out);
}

// Small coarse clouds (m <= TNS_MAX_M: the 1024 / 256 / 64-point levels of the semseg stacks): the CTA copies the cloud's
// sorted points and -- when the grid has at most TNS_MAX_CELLS cells -- its cell-start table into shared memory once
// and its 256 queries search there; every cell-table and candidate read is then a shared-memory load instead of a
// dependent L1 / L2 access (a query makes ~18 table reads and ~30 candidate reads in a chain).
constexpr int TNS_THREADS = 256, TNS_MAX_M = 2048, TNS_MAX_CELLS = 8192;

__global__ void __launch_bounds__(TNS_THREADS)
three_nn_grid_staged_kernel(int n, int m, int cell_cap, const float *__restrict__ unknown_all, const float *__restrict__ known_all,
                            const float4 *__restrict__ sorted_all, const int32_t *__restrict__ cell_start_all,
                            const GridMeta *__restrict__ meta_all, const int32_t *__restrict__ query_order,
                            float *__restrict__ dist2_out, int32_t *__restrict__ idx_out, float *__restrict__ weight_out) {
    extern __shared__ __align__(16) unsigned char tns_smem[];
    float4 *s_sorted = reinterpret_cast<float4 *>(tns_smem);
    int32_t *s_cells = reinterpret_cast<int32_t *>(s_sorted + m);
    const int b = blockIdx.y;
    const GridMeta g = meta_all[b];
    const float4 *sorted = sorted_all + (size_t)b * m;
    const int32_t *cell_start = cell_start_all + (size_t)b * (GRID_MAX_CELLS + 1);
    for (int e = threadIdx.x; e < m; e += TNS_THREADS) s_sorted[e] = __ldg(sorted + e);
    const bool cells_fit = g.ncells <= cell_cap;  // uniform over the CTA (cell_cap: what the launch reserved, <= TNS_MAX_CELLS)
    if (cells_fit)
        for (int e = threadIdx.x; e <= g.ncells; e += TNS_THREADS) s_cells[e] = __ldg(cell_start + e);
    __syncthreads();
    const int slot = blockIdx.x * TNS_THREADS + threadIdx.x;
    three_nn_grid_query(n, m, b, min(slot, n - 1), slot < n, g, unknown_all, known_all, s_sorted, cells_fit ? s_cells : cell_start, query_order, dist2_out,
                        idx_out, weight_out);
}

}  // namespace
}  // namespace pn2

extern "C" int pn2_grid_max_points(void) { return pn2::GRID_MAX_N_LARGE; }
extern "C" int pn2_grid_table_stride(void) { return pn2::GRID_MAX_CELLS + 1; }

extern "C" int pn2_grid_build(int b, int n, const float *xyz, float cell, float *sorted, int32_t *cell_start, int32_t *order,
                              float *meta, void *stream) {
    using namespace pn2;
    PN2_REQUIRE(b >= 0 && n >= 1, "grid_build: bad dims b=%d n=%d", b, n);
    if (n > GRID_MAX_N_LARGE)
        return set_error(PN2_ERR_UNSUPPORTED, "grid_build: at most %d points per cloud (got %d)", GRID_MAX_N_LARGE, n);
    PN2_REQUIRE(cell == cell, "grid_build: cell size is NaN (pass <= 0 for automatic)");
    if (b == 0) return PN2_OK;
    PN2_REQUIRE(xyz && sorted && cell_start && meta, "grid_build: null pointer");
    PN2_REQUIRE((((uintptr_t)sorted) & 15) == 0, "grid_build: sorted must be 16-byte aligned");
    if (n > GRID_MAX_N) {
        // counting sort by cell (several CTAs per cloud); per-cell cursors live in stream-ordered scratch
        PN2_REQUIRE(b <= 65535, "grid_build: b exceeds the grid limit");
        cudaStream_t s = (cudaStream_t)stream;
        const size_t bytes = (size_t)b * (GRID_MAX_CELLS + 1) * sizeof(int32_t);
        Scratch cursor_mem(s);  // released on every return below
        PN2_CUDA(cursor_mem.alloc(bytes));
        int32_t *cursor = (int32_t *)cursor_mem.ptr;
        PN2_CUDA(cudaMemsetAsync(cursor, 0, bytes, s));
        GridMeta *gm = reinterpret_cast<GridMeta *>(meta);
        grid_meta_kernel<<<b, 1024, 0, s>>>(n, cell, xyz, gm);
        PN2_LAUNCH_OK("grid_meta_kernel");
        const dim3 pts(ceil_div(n, 256), b);
        grid_count_kernel<<<pts, 256, 0, s>>>(n, xyz, gm, cursor);
        PN2_LAUNCH_OK("grid_count_kernel");
        grid_scan_kernel<<<b, 1024, 0, s>>>(n, gm, cursor, cell_start);
        PN2_LAUNCH_OK("grid_scan_kernel");
        grid_scatter_kernel<<<pts, 256, 0, s>>>(n, xyz, gm, cursor, reinterpret_cast<float4 *>(sorted), order);
        PN2_LAUNCH_OK("grid_scatter_kernel");
        return PN2_OK;
    }
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    const size_t smem = (size_t)np2 * sizeof(uint32_t);
    if (smem > 40 * 1024)
        PN2_CUDA(cudaFuncSetAttribute(grid_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    grid_build_kernel<<<b, 1024, smem, (cudaStream_t)stream>>>(n, cell, xyz, reinterpret_cast<float4 *>(sorted), cell_start, order,
                                                              reinterpret_cast<GridMeta *>(meta));
    PN2_LAUNCH_OK("grid_build");
    return PN2_OK;
}

extern "C" int pn2_ball_query_grid(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                                   const float *sorted, const int32_t *cell_start, const float *meta, int32_t *idx,
                                   void *stream) {
    using namespace pn2;
    PN2_REQUIRE(b >= 0 && n >= 1 && m >= 0 && nsample >= 0, "ball_query_grid: bad dims");
    if (b == 0 || m == 0 || nsample == 0) return PN2_OK;
    PN2_REQUIRE(new_xyz && xyz && sorted && cell_start && meta && idx, "ball_query_grid: null pointer");
    PN2_REQUIRE(b <= 65535, "ball_query_grid: b exceeds the grid limit");
    dim3 grid(ceil_div(m, BQG_WARPS), b);
    ball_query_grid_kernel<<<grid, BQG_WARPS * 32, 0, (cudaStream_t)stream>>>(
        n, m, radius, nsample, new_xyz, xyz, reinterpret_cast<const float4 *>(sorted), cell_start,
        reinterpret_cast<const GridMeta *>(meta), idx);
    PN2_LAUNCH_OK("ball_query_grid");
    return PN2_OK;
}

extern "C" int pn2_three_nn_grid(int b, int n, int m, const float *unknown, const float *known, const float *sorted,
                                 const int32_t *cell_start, const float *meta, const int32_t *query_order, float *dist2,
                                 int32_t *idx, float *weight, void *stream) {
    using namespace pn2;
    PN2_REQUIRE(b >= 0 && n >= 0 && m >= 1, "three_nn_grid: bad dims");
    if (b == 0 || n == 0) return PN2_OK;
    PN2_REQUIRE(unknown && known && sorted && cell_start && meta && idx, "three_nn_grid: null pointer");
    PN2_REQUIRE(b <= 65535, "three_nn_grid: b exceeds the grid limit");
    if (m <= TNS_MAX_M && n >= 2 * TNS_THREADS) {
        // room for a table of up to 4 cells per coarse point (the automatic cell size gives about one); larger tables
        // are read from global memory
        const int cell_cap = std::min(TNS_MAX_CELLS, 4 * m < 512 ? 512 : 4 * m);
        const size_t smem = (size_t)m * sizeof(float4) + (size_t)(cell_cap + 1) * sizeof(int32_t);
        PN2_CUDA(cudaFuncSetAttribute(three_nn_grid_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        three_nn_grid_staged_kernel<<<dim3(ceil_div(n, TNS_THREADS), b), TNS_THREADS, smem, (cudaStream_t)stream>>>(
            n, m, cell_cap, unknown, known, reinterpret_cast<const float4 *>(sorted), cell_start, reinterpret_cast<const GridMeta *>(meta), query_order,
            dist2, idx, weight);
        PN2_LAUNCH_OK("three_nn_grid_staged");
        return PN2_OK;
    }
    dim3 grid(ceil_div(n, 128), b);
    three_nn_grid_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(n, m, unknown, known, reinterpret_cast<const float4 *>(sorted),
                                                                cell_start, reinterpret_cast<const GridMeta *>(meta), query_order,
                                                                dist2, idx, weight);
    PN2_LAUNCH_OK("three_nn_grid");
    return PN2_OK;
}
