// train_mlp.cu -- the shared-MLP layer of the SA / FP blocks in TRAINING mode (batch-statistics BatchNorm), fp32.
//
// Replaces, per layer of model/pointnet_util.py:105-107 / 162-165 / 218-220 run under autograd,
//     Conv2d/Conv1d 1x1 (cuDNN fprop, dgrad, wgrad) + BatchNorm (forward statistics, normalise, backward reductions,
//     backward elementwise) + ReLU (forward, backward mask)
// i.e. nine library / elementwise passes over (rows x channels) activations, by four kernels on channel-last row
// matrices (a row = one (centroid, sample) pair or one point):
//
//   forward      z = act_in(x) W^T + b          and, in the same pass, the per-channel sum / sum of squares of z
//                (act_in = the PREVIOUS layer's BatchNorm + ReLU applied while the operand is loaded: normalised
//                activations are never written; only the pre-BatchNorm z of every layer is kept for the backward)
//   bwd reduce   S1 = sum_r dy, S2 = sum_r dy z  with dy = g [scale z + shift > 0]   (BatchNorm backward needs both)
//   bwd input    g_in = dz W                    with dz = ca dy + cb + cc z built while the operand is loaded
//   bwd weight   dW += dz^T act_in(x)           (row range per CTA, fp32 atomics into dW like the reference's backwards)
//
// BatchNorm algebra (training mode, biased variance, eps as torch): zhat = (z - mu) rstd, y = gamma zhat + beta,
// a = relu(y);  dL/dz = gamma rstd (dy - mean(dy) - zhat mean(dy zhat)) = ca dy + cb + cc z per channel with
//   ca = gamma rstd,  cc = -gamma rstd^3 (S2 - mu S1) / R,  cb = -ca S1 / R - cc mu;  dgamma = rstd (S2 - mu S1), dbeta = S1.
// The host side (pn2_b200/train_mlp.py) turns the sums into these coefficients (tiny per-channel tensors).
//
// All four are the same register-tiled SIMT GEMM (128 x NT output tile, 256 threads, k-chunks of 16 double buffered in
// shared memory, FFMA): the 1e-5 parity path.  Widths are arbitrary (16 .. 1536 in the MSG stack).
#include "common.cuh"

namespace pn2 {
namespace {

constexpr int TG_THREADS = 256;
constexpr int TG_KC = 16;
constexpr int TG_M = 128;

template <int NT, int MT = TG_M>
struct TileShape {
    static constexpr int TX = NT / 4;             // threads along n (4 columns each)
    static constexpr int TY = TG_THREADS / TX;    // threads along m
    static constexpr int TM = MT / TY;            // rows per thread (1, 2 or a multiple of 4)
    static constexpr int AS = MT + 4, BS = NT + 4;
    static constexpr size_t smem_floats = 2 * TG_KC * (AS + BS);
    static_assert(TM >= 1 && (TM < 4 || TM % 4 == 0), "tile shape");
};

// acc[TM][4] += As[k][ty*TM .. ] * Bs[k][tx*4 ..] over one staged k-chunk
template <int NT, int MT = TG_M>
__device__ __forceinline__ void mma_chunk(const float *__restrict__ As, const float *__restrict__ Bs, float (&acc)[TileShape<NT, MT>::TM][4],
                                          int ty, int tx) {
    using S = TileShape<NT, MT>;
#pragma unroll
    for (int kk = 0; kk < TG_KC; ++kk) {
        float a[S::TM], b[4];
        if constexpr (S::TM >= 4) {
#pragma unroll
            for (int i = 0; i < S::TM; i += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(As + kk * S::AS + ty * S::TM + i);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < S::TM; ++i) a[i] = As[kk * S::AS + ty * S::TM + i];
        }
        const float4 w = *reinterpret_cast<const float4 *>(Bs + kk * S::BS + tx * 4);
        b[0] = w.x; b[1] = w.y; b[2] = w.z; b[3] = w.w;
#pragma unroll
        for (int i = 0; i < S::TM; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

// ---- operand loaders: 4 consecutive k of one row (k-contiguous operands) -------------------------------------------------
struct ActIn {  // x with the previous layer's BatchNorm + ReLU folded in (scale == nullptr: identity)
    const float *x, *scale, *shift;
    long long rows;
    int c;
    __device__ __forceinline__ float one(long long r, int k) const {
        if (r >= rows || k >= c) return 0.f;
        float v = __ldg(x + r * c + k);
        if (scale) v = fmaxf(fmaf(v, __ldg(scale + k), __ldg(shift + k)), 0.f);
        return v;
    }
    // The load of an operand chunk is split in two: raw() issues the long-latency loads of the row data and nothing that
    // depends on them, finish() applies the per-channel transform.  The GEMM loops call raw() for chunk c + 1, multiply chunk
    // c, and only then finish() + store: with the transform inside the load the first dependent FFMA stalled the warp on the
    // loads (in-order issue) BEFORE the multiplies of the current chunk, so no load latency was ever hidden (ncu: the
    // transform's FFMA / FSEL carry the long-scoreboard samples, FMA pipe 10-13 % busy).
    struct Raw {
        float4 v;
    };
    __device__ __forceinline__ bool fast(long long r, int k) const { return r < rows && k + 3 < c && (c & 3) == 0; }
    __device__ __forceinline__ Raw raw(long long r, int k) const {  // k % 4 == 0
        Raw q;
        if (fast(r, k)) q.v = __ldg(reinterpret_cast<const float4 *>(x + r * c + k));
        else q.v = make_float4(one(r, k), one(r, k + 1), one(r, k + 2), one(r, k + 3));  // ragged edge: final values at once
        return q;
    }
    __device__ __forceinline__ float4 finish(const Raw &q, long long r, int k) const {
        float4 v = q.v;
        if (scale && fast(r, k)) {
            const float4 s = __ldg(reinterpret_cast<const float4 *>(scale + k)), t = __ldg(reinterpret_cast<const float4 *>(shift + k));
            v.x = fmaxf(fmaf(v.x, s.x, t.x), 0.f); v.y = fmaxf(fmaf(v.y, s.y, t.y), 0.f);
            v.z = fmaxf(fmaf(v.z, s.z, t.z), 0.f); v.w = fmaxf(fmaf(v.w, s.w, t.w), 0.f);
        }
        return v;
    }
};

struct DzIn {  // dz = ca * g * [scale z + shift > 0] + cb + cc * z, built from g and z
    const float *g, *z, *scale, *shift, *ca, *cb, *cc;
    long long rows;
    int c;
    __device__ __forceinline__ float one(long long r, int k) const {
        if (r >= rows || k >= c) return 0.f;
        const float zz = __ldg(z + r * c + k), gg = __ldg(g + r * c + k);
        const float dy = fmaf(zz, __ldg(scale + k), __ldg(shift + k)) > 0.f ? gg : 0.f;
        return fmaf(__ldg(cc + k), zz, fmaf(__ldg(ca + k), dy, __ldg(cb + k)));
    }
    struct Raw {
        float4 z, g;  // ragged edge: z holds the final values
    };
    __device__ __forceinline__ bool fast(long long r, int k) const { return r < rows && k + 3 < c && (c & 3) == 0; }
    __device__ __forceinline__ Raw raw(long long r, int k) const {
        Raw q;
        if (fast(r, k)) {
            q.z = __ldg(reinterpret_cast<const float4 *>(z + r * c + k));
            q.g = __ldg(reinterpret_cast<const float4 *>(g + r * c + k));
        } else {
            q.z = make_float4(one(r, k), one(r, k + 1), one(r, k + 2), one(r, k + 3));
            q.g = q.z;
        }
        return q;
    }
    __device__ __forceinline__ float4 finish(const Raw &q, long long r, int k) const {
        if (!fast(r, k)) return q.z;
        const float4 zz = q.z, gg = q.g;
        const float4 s = __ldg(reinterpret_cast<const float4 *>(scale + k)), t = __ldg(reinterpret_cast<const float4 *>(shift + k));
        const float4 a = __ldg(reinterpret_cast<const float4 *>(ca + k)), b = __ldg(reinterpret_cast<const float4 *>(cb + k));
        const float4 cq = __ldg(reinterpret_cast<const float4 *>(cc + k));
        float4 o;
        o.x = fmaf(cq.x, zz.x, fmaf(a.x, fmaf(zz.x, s.x, t.x) > 0.f ? gg.x : 0.f, b.x));
        o.y = fmaf(cq.y, zz.y, fmaf(a.y, fmaf(zz.y, s.y, t.y) > 0.f ? gg.y : 0.f, b.y));
        o.z = fmaf(cq.z, zz.z, fmaf(a.z, fmaf(zz.z, s.z, t.z) > 0.f ? gg.z : 0.f, b.z));
        o.w = fmaf(cq.w, zz.w, fmaf(a.w, fmaf(zz.w, s.w, t.w) > 0.f ? gg.w : 0.f, b.w));
        return o;
    }
};

struct WeightIn {  // w (n_total, k_total) row-major: 4 consecutive k of output column n
    const float *w;
    int n_total, k_total;
    __device__ __forceinline__ float4 four(int n, int k) const {
        if (n < n_total && k + 3 < k_total && (k_total & 3) == 0) return __ldg(reinterpret_cast<const float4 *>(w + (size_t)n * k_total + k));
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (n < n_total && k + e < k_total) ? __ldg(w + (size_t)n * k_total + k + e) : 0.f;
        return make_float4(v[0], v[1], v[2], v[3]);
    }
};

// C[128 x NT] = A[rows m0.., K] * B[n0.., K]^T with k-contiguous operands (forward and input-gradient products).
template <int NT, class LA>
__device__ __forceinline__ void gemm_rows(const LA &la, const WeightIn &lb, long long m0, int n0, int K, float *smem,
                                          float (&acc)[TileShape<NT>::TM][4]) {
    using S = TileShape<NT>;
    float *As = smem, *Bs = smem + 2 * TG_KC * S::AS;
    const int tid = threadIdx.x, tx = tid % S::TX, ty = tid / S::TX;
    // A: 128 rows x 16 k = 512 float4; thread -> (row = q / 4, kq = q % 4) for q = tid, tid + 256
    // B: NT columns x 16 k = NT * 4 float4
    constexpr int BQ = (NT * 4 + TG_THREADS - 1) / TG_THREADS;
    typename LA::Raw ra[2];
    float4 rb[BQ];
    auto load = [&](int k0) {  // long-latency loads only (see ActIn::raw)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int q = tid + u * TG_THREADS;
            ra[u] = la.raw(m0 + (q >> 2), k0 + 4 * (q & 3));
        }
#pragma unroll
        for (int u = 0; u < BQ; ++u) {
            const int q = tid + u * TG_THREADS;
            rb[u] = q < NT * 4 ? lb.four(n0 + (q >> 2), k0 + 4 * (q & 3)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store = [&](int buf, int k0) {
        float *a = As + buf * TG_KC * S::AS, *b = Bs + buf * TG_KC * S::BS;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int q = tid + u * TG_THREADS, r = q >> 2, kq = 4 * (q & 3);
            const float4 v = la.finish(ra[u], m0 + r, k0 + kq);
            a[(kq + 0) * S::AS + r] = v.x; a[(kq + 1) * S::AS + r] = v.y;
            a[(kq + 2) * S::AS + r] = v.z; a[(kq + 3) * S::AS + r] = v.w;
        }
#pragma unroll
        for (int u = 0; u < BQ; ++u) {
            const int q = tid + u * TG_THREADS;
            if (q < NT * 4) {
                const int n = q >> 2, kq = 4 * (q & 3);
                b[(kq + 0) * S::BS + n] = rb[u].x; b[(kq + 1) * S::BS + n] = rb[u].y;
                b[(kq + 2) * S::BS + n] = rb[u].z; b[(kq + 3) * S::BS + n] = rb[u].w;
            }
        }
    };
#pragma unroll
    for (int i = 0; i < S::TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int nchunks = (K + TG_KC - 1) / TG_KC;
    load(0);
    store(0, 0);
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) load((c + 1) * TG_KC);
        mma_chunk<NT>(As + (c & 1) * TG_KC * S::AS, Bs + (c & 1) * TG_KC * S::BS, acc, ty, tx);
        if (c + 1 < nchunks) store((c + 1) & 1, (c + 1) * TG_KC);
        __syncthreads();
    }
}

// ---- forward: z = act_in(x) W^T + b, per-channel sum / sum of squares of z ---------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(TG_THREADS, 2)
train_linear_fwd_kernel(ActIn xin, WeightIn w, const float *__restrict__ bias, float *__restrict__ z, double *__restrict__ stats,
                        long long tiles) {
    using S = TileShape<NT>;
    extern __shared__ __align__(16) float smem[];
    __shared__ float cpart[2][TG_THREADS / 32][NT];  // per-warp partial sums of the tile's columns (sum, sum of squares)
    const int cout = w.n_total, cin = w.k_total;
    const int tid = threadIdx.x, tx = tid % S::TX, ty = tid / S::TX;
    const int ntn = (cout + NT - 1) / NT;
    for (long long t = blockIdx.x; t < tiles * ntn; t += gridDim.x) {
        const long long mt = t / ntn;
        const int n0 = (int)(t - mt * ntn) * NT;
        const long long m0 = mt * TG_M;
        float acc[S::TM][4];
        gemm_rows<NT>(xin, w, m0, n0, cin, smem, acc);
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        const int col = n0 + tx * 4;
        float bv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = col + j < cout ? __ldg(bias + col + j) : 0.f;
#pragma unroll
        for (int i = 0; i < S::TM; ++i) {
            const long long r = m0 + ty * S::TM + i;
            if (r >= xin.rows) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = acc[i][j] + bv[j];
                if (col + j < cout) {
                    s1[j] += v[j];
                    s2[j] = fmaf(v[j], v[j], s2[j]);
                }
            }
            if (col + 3 < cout && (cout & 3) == 0) {
                *reinterpret_cast<float4 *>(z + r * cout + col) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (col + j < cout) z[r * cout + col + j] = v[j];
            }
        }
        // column sums: lanes of a warp that share tx (32 / TX rows of threads) by shuffles, the 8 warps through shared
        // memory (shared-memory float atomics compile to compare-and-swap loops: 8-way contended, they were 7 % of the
        // kernel's instructions and a quarter of its stall samples)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int o = S::TX; o < 32; o <<= 1) {
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
            }
        }
        if ((tid & 31) < S::TX) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                cpart[0][tid >> 5][tx * 4 + j] = s1[j];
                cpart[1][tid >> 5][tx * 4 + j] = s2[j];
            }
        }
        __syncthreads();
        if (tid < 2 * NT && n0 + (tid % NT) < cout) {
            float t = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < TG_THREADS / 32; ++w8) t += cpart[tid / NT][w8][tid % NT];
            atomicAdd(stats + (size_t)(tid / NT) * cout + n0 + (tid % NT), (double)t);
        }
        __syncthreads();
    }
}

// ---- backward reductions of BatchNorm: S1 = sum dy, S2 = sum dy z -------------------------------------------------------------
__global__ void __launch_bounds__(256)
train_bn_bwd_reduce_kernel(long long rows, int c, const float *__restrict__ g, const float *__restrict__ z, const float *__restrict__ scale,
                           const float *__restrict__ shift, double *__restrict__ sums) {
    // lanes along channels (coalesced rows), the 8 warps of a CTA and the CTAs of the grid along rows
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ float red[2][8][32];
    for (int c0 = 0; c0 < c; c0 += 32) {
        const int ch = c0 + lane;
        float s1 = 0.f, s2 = 0.f;
        if (ch < c) {
            const float sc = __ldg(scale + ch), sh = __ldg(shift + ch);
            for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += (long long)gridDim.x * 8) {
                const float zz = __ldg(z + r * c + ch), gg = __ldg(g + r * c + ch);
                const float dy = fmaf(zz, sc, sh) > 0.f ? gg : 0.f;
                s1 += dy;
                s2 = fmaf(dy, zz, s2);
            }
        }
        red[0][warp][lane] = s1;
        red[1][warp][lane] = s2;
        __syncthreads();
        if (warp < 2 && ch < c) {
            float t = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) t += red[warp][w8][lane];
            atomicAdd(sums + (size_t)warp * c + ch, (double)t);
        }
        __syncthreads();
    }
}

// ---- backward, input gradient: g_in = dz W  (W^T given as wt (cin, cout) row-major) -------------------------------------------
template <int NT>
__global__ void __launch_bounds__(TG_THREADS, 2)
train_linear_bwd_input_kernel(DzIn dz, WeightIn wt, float *__restrict__ g_in, long long tiles) {
    using S = TileShape<NT>;
    extern __shared__ __align__(16) float smem[];
    const int cin = wt.n_total, cout = wt.k_total;
    const int tid = threadIdx.x, tx = tid % S::TX, ty = tid / S::TX;
    const int ntn = (cin + NT - 1) / NT;
    for (long long t = blockIdx.x; t < tiles * ntn; t += gridDim.x) {
        const long long mt = t / ntn;
        const int n0 = (int)(t - mt * ntn) * NT;
        const long long m0 = mt * TG_M;
        float acc[S::TM][4];
        gemm_rows<NT>(dz, wt, m0, n0, cout, smem, acc);
        const int col = n0 + tx * 4;
#pragma unroll
        for (int i = 0; i < S::TM; ++i) {
            const long long r = m0 + ty * S::TM + i;
            if (r >= dz.rows) continue;
            if (col + 3 < cin && (cin & 3) == 0) {
                *reinterpret_cast<float4 *>(g_in + r * cin + col) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (col + j < cin) g_in[r * cin + col + j] = acc[i][j];
            }
        }
    }
}

// ---- backward, weight gradient: dW[co][ci] += sum_r dz[r][co] act_in(x)[r][ci] over a row range per CTA -----------------------
// Both operands are contiguous along their channel (row-major activations), so the k-chunk (16 rows) is staged as it lies:
// As[k][co], Bs[k][ci].  Output tile 128 (co) x NT (ci); the CTA walks its rows once per output tile.
template <int NT, int MT>
__global__ void __launch_bounds__(TG_THREADS, 2)
train_linear_bwd_weight_kernel(DzIn dz, ActIn xin, float *__restrict__ dw, long long rows_per_cta) {
    using S = TileShape<NT, MT>;
    extern __shared__ __align__(16) float smem[];
    float *As = smem, *Bs = smem + 2 * TG_KC * S::AS;
    const int cout = dz.c, cin = xin.c;
    const int tid = threadIdx.x, tx = tid % S::TX, ty = tid / S::TX;
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    const long long r_end = min(dz.rows, r_begin + rows_per_cta);
    if (r_begin >= r_end) return;
    const int nchunks = (int)((r_end - r_begin + TG_KC - 1) / TG_KC);
    constexpr int BQ = (NT * 4 + TG_THREADS - 1) / TG_THREADS;
    // blockIdx.y = output tile (m0, n0): deep layers have few rows but many tiles, shallow ones the opposite
    const int ntn = (cin + NT - 1) / NT;
        {
            const int m0 = (int)(blockIdx.y / ntn) * MT, n0 = (int)(blockIdx.y % ntn) * NT;
            float acc[S::TM][4];
#pragma unroll
            for (int i = 0; i < S::TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            constexpr int AQ = (MT * 4 + TG_THREADS - 1) / TG_THREADS;
            DzIn::Raw ra[AQ];
            ActIn::Raw rb[BQ];
            // A chunk: 16 rows x MT co = MT * 4 float4: q -> (k = q / (MT / 4), co4 = q % (MT / 4));  B chunk: 16 rows x NT ci
            // rows past the CTA's range are read as rows past the end of the matrix (zeros)
            auto load = [&](long long r0) {  // long-latency loads only (see ActIn::raw)
#pragma unroll
                for (int u = 0; u < AQ; ++u) {
                    const int q = tid + u * TG_THREADS;
                    const long long r = r0 + q / (MT / 4);
                    if (q < MT * 4) ra[u] = dz.raw(r < r_end ? r : dz.rows, m0 + 4 * (q % (MT / 4)));
                }
#pragma unroll
                for (int u = 0; u < BQ; ++u) {
                    const int q = tid + u * TG_THREADS;
                    const long long r = r0 + q / (NT / 4);
                    if (q < NT * 4) rb[u] = xin.raw(r < r_end ? r : xin.rows, n0 + 4 * (q % (NT / 4)));
                }
            };
            auto store = [&](int buf, long long r0) {
                float *a = As + buf * TG_KC * S::AS, *b = Bs + buf * TG_KC * S::BS;
#pragma unroll
                for (int u = 0; u < AQ; ++u) {
                    const int q = tid + u * TG_THREADS;
                    const long long r = r0 + q / (MT / 4);
                    if (q < MT * 4)
                        *reinterpret_cast<float4 *>(a + (q / (MT / 4)) * S::AS + 4 * (q % (MT / 4))) =
                            dz.finish(ra[u], r < r_end ? r : dz.rows, m0 + 4 * (q % (MT / 4)));
                }
#pragma unroll
                for (int u = 0; u < BQ; ++u) {
                    const int q = tid + u * TG_THREADS;
                    const long long r = r0 + q / (NT / 4);
                    if (q < NT * 4)
                        *reinterpret_cast<float4 *>(b + (q / (NT / 4)) * S::BS + 4 * (q % (NT / 4))) =
                            xin.finish(rb[u], r < r_end ? r : xin.rows, n0 + 4 * (q % (NT / 4)));
                }
            };
            __syncthreads();  // the previous output tile's last chunk has been consumed
            load(r_begin);
            store(0, r_begin);
            __syncthreads();
            for (int c = 0; c < nchunks; ++c) {
                if (c + 1 < nchunks) load(r_begin + (long long)(c + 1) * TG_KC);
                mma_chunk<NT, MT>(As + (c & 1) * TG_KC * S::AS, Bs + (c & 1) * TG_KC * S::BS, acc, ty, tx);
                if (c + 1 < nchunks) store((c + 1) & 1, r_begin + (long long)(c + 1) * TG_KC);
                __syncthreads();
            }
#pragma unroll
            for (int i = 0; i < S::TM; ++i) {
                const int co = m0 + ty * S::TM + i;
                if (co >= cout) continue;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ci = n0 + tx * 4 + j;
                    if (ci < cin) atomicAdd(dw + (size_t)co * cin + ci, acc[i][j]);
                }
            }
        }
}

// ---- the BatchNorm + ReLU of the LAST layer of a stack, materialised (what the max-pool / the next block consumes) ------------
__global__ void __launch_bounds__(256)
train_bn_relu_kernel(long long total4, int c, const float *__restrict__ z, const float *__restrict__ scale, const float *__restrict__ shift,
                     float *__restrict__ a) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total4; e += (long long)gridDim.x * 256) {
        const int k = (int)((e * 4) % c);
        const float4 v = __ldg(reinterpret_cast<const float4 *>(z) + e);
        const float4 s = __ldg(reinterpret_cast<const float4 *>(scale + k)), t = __ldg(reinterpret_cast<const float4 *>(shift + k));
        reinterpret_cast<float4 *>(a)[e] = make_float4(fmaxf(fmaf(v.x, s.x, t.x), 0.f), fmaxf(fmaf(v.y, s.y, t.y), 0.f),
                                                       fmaxf(fmaf(v.z, s.z, t.z), 0.f), fmaxf(fmaf(v.w, s.w, t.w), 0.f));
    }
}
__global__ void __launch_bounds__(256)
train_bn_relu_scalar_kernel(long long total, int c, const float *__restrict__ z, const float *__restrict__ scale,
                            const float *__restrict__ shift, float *__restrict__ a) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int k = (int)(e % c);
        a[e] = fmaxf(fmaf(__ldg(z + e), __ldg(scale + k), __ldg(shift + k)), 0.f);
    }
}

// ---- per-channel coefficient kernels (one tiny launch instead of ~15 elementwise torch launches per layer and direction) ----
// forward: (sum z, sum z^2) -> mean, biased var, rstd (fp64 -> fp32 where stored), scale = gamma rstd, shift = beta - mean scale;
// running statistics updated like torch.nn.BatchNorm in training mode (momentum, UNBIASED variance).
__global__ void train_bn_finalize_kernel(int c, double rows, const double *__restrict__ stats, const float *__restrict__ gamma,
                                         const float *__restrict__ beta, double eps, double momentum, float *__restrict__ scale,
                                         float *__restrict__ shift, double *__restrict__ mean_rstd, float *__restrict__ running_mean,
                                         float *__restrict__ running_var) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= c) return;
    const double mean = stats[k] / rows;
    double var = stats[c + k] / rows - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const double rstd = rsqrt(var + eps);
    const double g = (double)gamma[k];
    scale[k] = (float)(g * rstd);
    shift[k] = (float)((double)beta[k] - mean * g * rstd);
    mean_rstd[k] = mean;
    mean_rstd[c + k] = rstd;
    if (running_mean) {
        running_mean[k] = (float)((1.0 - momentum) * (double)running_mean[k] + momentum * (double)(float)mean);
        const double unbiased = (double)(float)var * (rows / (rows > 1.0 ? rows - 1.0 : 1.0));
        running_var[k] = (float)((1.0 - momentum) * (double)running_var[k] + momentum * unbiased);
    }
}

// backward: (S1 = sum dy, S2 = sum dy z) -> ca, cb, cc of dz = ca dy + cb + cc z; dgamma = rstd (S2 - mean S1), dbeta = S1
__global__ void train_bn_bwd_coeffs_kernel(int c, double rows, const double *__restrict__ sums, const double *__restrict__ mean_rstd,
                                           const float *__restrict__ gamma, float *__restrict__ coef, float *__restrict__ dgamma,
                                           float *__restrict__ dbeta) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= c) return;
    const double S1 = sums[k], S2 = sums[c + k], mean = mean_rstd[k], rstd = mean_rstd[c + k], g = (double)gamma[k];
    const double t = S2 - mean * S1;
    const double ca = g * rstd;
    const double cc = -g * rstd * rstd * rstd * t / rows;
    const double cb = -ca * S1 / rows - cc * mean;
    coef[k] = (float)ca;
    coef[c + k] = (float)cb;
    coef[2 * c + k] = (float)cc;
    dgamma[k] = (float)(rstd * t);
    dbeta[k] = (float)S1;
}

int pick_nt(int n) { return n <= 32 ? 32 : (n <= 64 ? 64 : 128); }

template <int NT>
size_t tile_smem() { return TileShape<NT>::smem_floats * sizeof(float); }

}  // namespace
}  // namespace pn2

using namespace pn2;

extern "C" int pn2_train_linear_fwd(long long rows, int cin, int cout, const float *x, const float *in_scale, const float *in_shift,
                                    const float *w, const float *bias, float *z, double *stats, void *stream) {
    PN2_REQUIRE(rows >= 0 && cin >= 1 && cout >= 1, "train_linear_fwd: bad dims rows=%lld cin=%d cout=%d", rows, cin, cout);
    if (rows == 0) return PN2_OK;
    PN2_REQUIRE(x && w && bias && z && stats && ((in_scale == nullptr) == (in_shift == nullptr)), "train_linear_fwd: null pointer");
    PN2_REQUIRE((((uintptr_t)x | (uintptr_t)w | (uintptr_t)z | (uintptr_t)in_scale | (uintptr_t)in_shift) & 15) == 0, "train_linear_fwd: 16-byte alignment");
    cudaStream_t s = (cudaStream_t)stream;
    ActIn a = {x, in_scale, in_shift, rows, cin};
    WeightIn b = {w, cout, cin};
    const long long tiles = (rows + TG_M - 1) / TG_M;
    const int nt = pick_nt(cout);
    const long long work = tiles * ((cout + nt - 1) / nt);
    const unsigned grid = (unsigned)(work < 2ll * sm_count() ? work : 2ll * sm_count());
    if (nt == 32) {
        train_linear_fwd_kernel<32><<<grid, TG_THREADS, tile_smem<32>(), s>>>(a, b, bias, z, stats, tiles);
    } else if (nt == 64) {
        train_linear_fwd_kernel<64><<<grid, TG_THREADS, tile_smem<64>(), s>>>(a, b, bias, z, stats, tiles);
    } else {
        train_linear_fwd_kernel<128><<<grid, TG_THREADS, tile_smem<128>(), s>>>(a, b, bias, z, stats, tiles);
    }
    PN2_LAUNCH_OK("train_linear_fwd_kernel");
    return PN2_OK;
}

extern "C" int pn2_train_bn_bwd_reduce(long long rows, int c, const float *g, const float *z, const float *scale, const float *shift,
                                       double *sums, void *stream) {
    PN2_REQUIRE(rows >= 0 && c >= 1, "train_bn_bwd_reduce: bad dims");
    if (rows == 0) return PN2_OK;
    PN2_REQUIRE(g && z && scale && shift && sums, "train_bn_bwd_reduce: null pointer");
    long long want = (rows + 63) / 64;
    const unsigned grid = (unsigned)(want < 4ll * sm_count() ? (want < 1 ? 1 : want) : 4ll * sm_count());
    train_bn_bwd_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rows, c, g, z, scale, shift, sums);
    PN2_LAUNCH_OK("train_bn_bwd_reduce_kernel");
    return PN2_OK;
}

extern "C" int pn2_train_linear_bwd(long long rows, int cin, int cout, const float *x, const float *in_scale, const float *in_shift,
                                    const float *wt, const float *g, const float *z, const float *scale, const float *shift,
                                    const float *ca, const float *cb, const float *cc, float *g_in, float *dw, void *stream) {
    PN2_REQUIRE(rows >= 0 && cin >= 1 && cout >= 1, "train_linear_bwd: bad dims rows=%lld cin=%d cout=%d", rows, cin, cout);
    if (rows == 0) return PN2_OK;
    PN2_REQUIRE(x && wt && g && z && scale && shift && ca && cb && cc && dw && ((in_scale == nullptr) == (in_shift == nullptr)),
                "train_linear_bwd: null pointer");
    // 128-bit loads are only used for widths that are multiples of 4: those tensors must then be 16-byte aligned
    PN2_REQUIRE((((uintptr_t)x | (uintptr_t)wt | (uintptr_t)g | (uintptr_t)z | (uintptr_t)g_in | (uintptr_t)in_scale | (uintptr_t)in_shift) & 15) == 0,
                "train_linear_bwd: 16-byte alignment");
    PN2_REQUIRE((cout & 3) != 0 || (((uintptr_t)scale | (uintptr_t)shift | (uintptr_t)ca | (uintptr_t)cb | (uintptr_t)cc) & 15) == 0,
                "train_linear_bwd: 16-byte alignment of the per-channel vectors");
    cudaStream_t s = (cudaStream_t)stream;
    DzIn dz = {g, z, scale, shift, ca, cb, cc, rows, cout};
    ActIn a = {x, in_scale, in_shift, rows, cin};
    const long long tiles = (rows + TG_M - 1) / TG_M;
    if (g_in) {
        WeightIn b = {wt, cin, cout};
        const int nt = pick_nt(cin);
        const long long work = tiles * ((cin + nt - 1) / nt);
        const unsigned grid = (unsigned)(work < 2ll * sm_count() ? work : 2ll * sm_count());
        if (nt == 32) train_linear_bwd_input_kernel<32><<<grid, TG_THREADS, tile_smem<32>(), s>>>(dz, b, g_in, tiles);
        else if (nt == 64) train_linear_bwd_input_kernel<64><<<grid, TG_THREADS, tile_smem<64>(), s>>>(dz, b, g_in, tiles);
        else train_linear_bwd_input_kernel<128><<<grid, TG_THREADS, tile_smem<128>(), s>>>(dz, b, g_in, tiles);
        PN2_LAUNCH_OK("train_linear_bwd_input_kernel");
    }
    {
        // grid = (row ranges, output tiles): about four CTAs per SM in total, at least 128 rows per range (the partial sums
        // go to dW with atomics)
        const int nt = pick_nt(cin), mt = pick_nt(cout);  // output tile (cout x cin) = mt x nt: no padded FMAs for narrow layers
        const long long out_tiles = (long long)((cout + mt - 1) / mt) * ((cin + nt - 1) / nt);
        PN2_REQUIRE(out_tiles <= 65535, "train_linear_bwd: too many output tiles");
        long long ranges = (4ll * sm_count() + out_tiles - 1) / out_tiles;
        long long per = (rows + ranges - 1) / ranges;
        if (per < 128) per = 128;
        per = (per + TG_KC - 1) / TG_KC * TG_KC;
        ranges = (rows + per - 1) / per;
        const dim3 grid((unsigned)ranges, (unsigned)out_tiles);
#define PN2_WG(NTV, MTV) train_linear_bwd_weight_kernel<NTV, MTV><<<grid, TG_THREADS, TileShape<NTV, MTV>::smem_floats * sizeof(float), s>>>(dz, a, dw, per)
        if (mt == 32) {
            if (nt == 32) PN2_WG(32, 32); else if (nt == 64) PN2_WG(64, 32); else PN2_WG(128, 32);
        } else if (mt == 64) {
            if (nt == 32) PN2_WG(32, 64); else if (nt == 64) PN2_WG(64, 64); else PN2_WG(128, 64);
        } else {
            if (nt == 32) PN2_WG(32, 128); else if (nt == 64) PN2_WG(64, 128); else PN2_WG(128, 128);
        }
#undef PN2_WG
        PN2_LAUNCH_OK("train_linear_bwd_weight_kernel");
    }
    return PN2_OK;
}

extern "C" int pn2_train_bn_relu(long long rows, int c, const float *z, const float *scale, const float *shift, float *a, void *stream) {
    PN2_REQUIRE(rows >= 0 && c >= 1, "train_bn_relu: bad dims");
    if (rows == 0) return PN2_OK;
    PN2_REQUIRE(z && scale && shift && a, "train_bn_relu: null pointer");
    const long long total = rows * c;
    cudaStream_t s = (cudaStream_t)stream;
    if ((c & 3) == 0 && (((uintptr_t)z | (uintptr_t)a | (uintptr_t)scale | (uintptr_t)shift) & 15) == 0) {
        const long long t4 = total / 4;
        const unsigned grid = (unsigned)((t4 + 255) / 256 < 8ll * sm_count() ? (t4 + 255) / 256 : 8ll * sm_count());
        train_bn_relu_kernel<<<grid, 256, 0, s>>>(t4, c, z, scale, shift, a);
    } else {
        const unsigned grid = (unsigned)((total + 255) / 256 < 8ll * sm_count() ? (total + 255) / 256 : 8ll * sm_count());
        train_bn_relu_scalar_kernel<<<grid, 256, 0, s>>>(total, c, z, scale, shift, a);
    }
    PN2_LAUNCH_OK("train_bn_relu_kernel");
    return PN2_OK;
}

extern "C" int pn2_train_bn_finalize(long long rows, int c, const double *stats, const float *gamma, const float *beta, double eps,
                                     double momentum, float *scale, float *shift, double *mean_rstd, float *running_mean,
                                     float *running_var, void *stream) {
    PN2_REQUIRE(rows >= 1 && c >= 1, "train_bn_finalize: bad dims");
    PN2_REQUIRE(stats && gamma && beta && scale && shift && mean_rstd && ((running_mean == nullptr) == (running_var == nullptr)),
                "train_bn_finalize: null pointer");
    train_bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(c, (double)rows, stats, gamma, beta, eps, momentum, scale,
                                                                                shift, mean_rstd, running_mean, running_var);
    PN2_LAUNCH_OK("train_bn_finalize_kernel");
    return PN2_OK;
}

extern "C" int pn2_train_bn_bwd_coeffs(long long rows, int c, const double *sums, const double *mean_rstd, const float *gamma, float *coef,
                                       float *dgamma, float *dbeta, void *stream) {
    PN2_REQUIRE(rows >= 1 && c >= 1, "train_bn_bwd_coeffs: bad dims");
    PN2_REQUIRE(sums && mean_rstd && gamma && coef && dgamma && dbeta, "train_bn_bwd_coeffs: null pointer");
    train_bn_bwd_coeffs_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(c, (double)rows, sums, mean_rstd, gamma, coef, dgamma, dbeta);
    PN2_LAUNCH_OK("train_bn_bwd_coeffs_kernel");
    return PN2_OK;
}
