// row_mlp_tile.cuh -- the fp32 register-tiled "rows x weights^T" tile shared by the fused inference blocks (row_mlp.cu)
// and the training kernels (train_mlp.cu).  Activations live in shared memory k-major ([channel][row], row stride TR + 4)
// so that a thread's A fragment is one or two LDS.128; weight k-chunks are staged through a double-buffered tile with their
// columns permuted to match.
#pragma once
#include "common.cuh"

namespace pn2 {
namespace {

constexpr int RM_THREADS = 256;
constexpr int KC = 16;          // k-chunk of the staged weight tile
constexpr int WSP = 128 + 4;    // row stride of the weight tile (floats)

// n-tile width of a layer; the 16-row tile has 64 thread columns, so its tiles are at least 64 wide
__host__ __device__ inline int pick_nt(int cout, int tr) { return cout <= 32 ? (tr == 16 ? 64 : 32) : (cout <= 64 ? 64 : 128); }
__host__ __device__ inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

// One n-tile of one layer: acc = X_in[.,rows] * W[n0 + cols, .]^T, then bias/ReLU and a k-major store.
template <int TR, int NT>
__device__ __forceinline__ void layer_tile(const float *xin, float *xout, int out_ch0,
                                           const float *__restrict__ W, const float *__restrict__ bias, int cin,
                                           int cout, int n0, int relu, float *ws) {
    constexpr int TRP = TR + 4;
    constexpr int TM = (TR == 128 && NT >= 64) ? 8 : 4;
    constexpr int TY = TR / TM;
    constexpr int TXN = RM_THREADS / TY;
    constexpr int TN = NT / TXN;
    constexpr int EPT = KC * NT / RM_THREADS;  // weight elements staged per thread per chunk
    static_assert(TN >= 1 && EPT >= 1, "bad tile");
    const int tid = threadIdx.x;
    // thread -> (tx, ty).  A 128-bit shared load costs one wavefront per distinct 128-byte line and HALF warp (ncu: 2 for an
    // A fragment shared by 16 lanes, 4 for 16 lanes x 16 B of B read by both halves), so a warp is laid out as 8 tx x 4 ty:
    // each half warp reads one 128-byte line of B and two 16-byte pieces of one line of A.  Narrow thread tiles (TN < 4:
    // 64- / 32-bit B loads) keep consecutive lanes along tx.
    constexpr int LTX = (TN >= 4 && TXN >= 8) ? 8 : (TXN < 32 ? TXN : 32);
    constexpr int WX = TXN / LTX;
    const int lane_ = tid & 31, warp_ = tid >> 5;
    const int tx = (warp_ % WX) * LTX + lane_ % LTX, ty = (warp_ / WX) * (32 / LTX) + lane_ / LTX;

    // staging coordinates: smem position p <-> logical column.  A thread's TN columns are tx + TXN * j; they sit in
    // groups of four so that the lanes of a quarter warp read CONSECUTIVE 16-byte slots (p = (j / 4) * 4 TXN + 4 tx + j % 4):
    // with the eight columns of a thread contiguous, lanes 32 bytes apart met two to a bank group on every 128-bit load.
    const int sp = tid % NT, sg = tid / NT;
    constexpr int GW = TN >= 4 ? 4 : TN;  // columns of one thread that are contiguous in the tile
    const int scol = n0 + ((sp % (TXN * GW)) / GW) + TXN * (GW * (sp / (TXN * GW)) + sp % GW);
    const float *wrow = W + (size_t)scol * cin;
    const bool col_ok = scol < cout;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float wreg[EPT];
    const int nchunks = (cin + KC - 1) / KC;
    // prologue: stage chunk 0
    // a thread's EPT weights are consecutive k of one column: 128-bit loads when the rows allow it (a scalar load per
    // element costs a 32-byte sector request per lane -- the columns of a warp are cin floats apart)
    const bool wvec = (EPT % 4 == 0) && (cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    auto load_chunk = [&](int k0) {
        if (wvec) {
#pragma unroll
            for (int e = 0; e < EPT; e += 4) {
                const int k = k0 + sg * EPT + e;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (col_ok && k < cin) v = __ldg(reinterpret_cast<const float4 *>(wrow + k));
                wreg[e] = v.x; wreg[e + 1] = v.y; wreg[e + 2] = v.z; wreg[e + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int k = k0 + sg * EPT + e;
                wreg[e] = (col_ok && k < cin) ? __ldg(wrow + k) : 0.f;
            }
        }
    };
    load_chunk(0);
#pragma unroll
    for (int e = 0; e < EPT; ++e) ws[(sg * EPT + e) * WSP + sp] = wreg[e];
    __syncthreads();

    for (int c = 0; c < nchunks; ++c) {
        float *wcur = ws + (c & 1) * (KC * WSP);
        float *wnxt = ws + ((c + 1) & 1) * (KC * WSP);
        const bool more = (c + 1) < nchunks;
        if (more) load_chunk((c + 1) * KC);
        const float *xa = xin + (size_t)(c * KC) * TRP + ty * TM;
        const int kk_end = min(KC, cin - c * KC);
        auto step = [&](int kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(xa + kk * TRP + i);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
            const float *wb = wcur + kk * WSP + tx * GW;
            if constexpr (TN >= 4) {
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    const float4 v = *reinterpret_cast<const float4 *>(wb + j * TXN);
                    b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
                }
            } else if constexpr (TN == 2) {
                const float2 v = *reinterpret_cast<const float2 *>(wb);
                b[0] = v.x; b[1] = v.y;
            } else {
                b[0] = wb[0];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        };
        if (kk_end == KC) {  // whole chunk: no per-step bound checks, loads scheduled across the 16 steps
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) step(kk);
        } else {
            for (int kk = 0; kk < kk_end; ++kk) step(kk);
        }
        if (more) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) wnxt[(sg * EPT + e) * WSP + sp] = wreg[e];
        }
        __syncthreads();
    }

    // epilogue: bias + activation, k-major store (channel = out_ch0 + logical column offset)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int col = n0 + tx + TXN * j;
        const float bv = col < cout ? __ldg(bias + col) : 0.f;
        float *dst = xout + (size_t)(out_ch0 + tx + TXN * j) * TRP + ty * TM;
#pragma unroll
        for (int i = 0; i < TM; i += 4) {
            float4 v;
            v.x = acc[i][j] + bv; v.y = acc[i + 1][j] + bv; v.z = acc[i + 2][j] + bv; v.w = acc[i + 3][j] + bv;
            if (relu) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            }
            *reinterpret_cast<float4 *>(dst + i) = v;
        }
    }
}

}  // namespace
}  // namespace pn2
