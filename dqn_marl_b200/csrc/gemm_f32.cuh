// fp32 CUDA-core GEMM with implicit-convolution operand loaders — the parity path of the Q-network.
//
// The reference's DQNNetwork (Louvre_Evacuation/agents/dqn_agent.py:15-61) is evaluated by PyTorch in fp32
// and north_star asks for Q-values / losses within 1e-5 relative of it, which rules out TF32/bf16 tensor
// cores for THIS path: every contraction here is plain FFMA with fp32 accumulation.  One templated kernel
// covers all twelve contractions of a learn step (3 convs + 2 fcs forward, their dgrad and wgrad) through
// operand-loader modes, so the 3x3/pad-1 convolutions never materialise an im2col buffer in HBM.
//
// Tiling: CTA tile 128 x BN x 16 (BN = 128/64/32), 256 threads, 8 x (BN/16) register micro-tile, operands
// staged k-major in shared memory and double-buffered through registers.  Split-K over gridDim.z for the
// skinny shapes (fc1 at small batch, every wgrad), partials reduced by `splitk_epilogue_kernel`.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mq {

enum AMode { A_ROW = 0, A_COL = 1, A_IM2COL = 2, A_IM2COL_T = 3, A_IM2COL_FLIP = 4 };
enum BMode { B_ROW = 0, B_COL = 1, B_CONVW_T = 2 };

struct GemmParams {
    int M, N, K;
    const float* A; const float* B; float* C;
    __nv_bfloat16* Cb;           // optional bf16 copy of C (same ld): operand of the tensor-core path
    int lda, ldb, ldc;
    int batch;                   // conv: number of images (M or K = batch * 121)
    // epilogue: C = f(acc + bias[n]) with optional relu, * (mask_act[m][n] > 0), * drop[m][n] * drop_scale
    const float* bias;
    int relu;
    const float* mask_act;       // same shape / ld as C
    const uint8_t* drop;         // [M][N] u8 keep-mask or null
    float drop_scale;
    // split-K
    float* partial;              // [splits][M][N] when splits > 1
    int splits, k_chunk;
    // Small-batch form (launch_gemm_swapped): the kernel computes C^T — its rows are the ORIGINAL columns — so that a batch of
    // <= 32 rows becomes the narrow BN = 32 side of the tile instead of filling 1..32 of its 128 rows.  bias / mask_act / drop
    // and the store are indexed in ORIGINAL coordinates (row = kernel column); drop_ld = row length of `drop` (original N).
    int trans_out, drop_ld;
};

constexpr int GEMM_BM = 128, GEMM_BK = 16, GEMM_THREADS = 256;
constexpr int OBS_HW = 11, OBS_PIX = 121;

// ---- operand loaders (element-wise; divisors are compile-time constants) -------------------------------
// A_IM2COL:      X NHWC [batch][11][11][C]; m = (b, i, j); k = (tap, c) -> X[b][i+di-1][j+dj-1][c]
// A_IM2COL_T:    transposed view of the same matrix: m = (tap, c); k = (b, i, j)           (conv wgrad)
// A_IM2COL_FLIP: dY NHWC [batch][11][11][C]; m = (b, i, j); k = (tap, n) -> dY[b][i-(di-1)][j-(dj-1)][n]  (conv dgrad)
template <int AMODE, int CDIM>
__device__ __forceinline__ float load_a(const GemmParams& p, int m, int k) {
    if (m >= p.M || k >= p.K) return 0.f;
    if (AMODE == A_ROW) return __ldg(p.A + (size_t)m * p.lda + k);
    if (AMODE == A_COL) return __ldg(p.A + (size_t)k * p.lda + m);
    int pix, tc;
    if (AMODE == A_IM2COL_T) { pix = k; tc = m; } else { pix = m; tc = k; }
    const int b = pix / OBS_PIX, q = pix - b * OBS_PIX;
    const int i = q / OBS_HW, j = q - i * OBS_HW;
    const int tap = tc / CDIM, c = tc - tap * CDIM;
    const int di = tap / 3 - 1, dj = tap - (tap / 3) * 3 - 1;
    const int ii = (AMODE == A_IM2COL_FLIP) ? i - di : i + di;
    const int jj = (AMODE == A_IM2COL_FLIP) ? j - dj : j + dj;
    if ((unsigned)ii >= (unsigned)OBS_HW || (unsigned)jj >= (unsigned)OBS_HW) return 0.f;
    return __ldg(p.A + ((size_t)(b * OBS_PIX + ii * OBS_HW + jj)) * CDIM + c);
}
// B_ROW:     B[k][n] (row-major, n contiguous)
// B_COL:     B[n][k] (PyTorch Linear weight)
// B_CONVW_T: conv weight Wc[(tap*Cin + c)][Cout] read as (k = (tap, n), col = c); CDIM = Cout, N = Cin  (conv dgrad)
template <int BMODE, int CDIM>
__device__ __forceinline__ float load_b(const GemmParams& p, int k, int n) {
    if (k >= p.K || n >= p.N) return 0.f;
    if (BMODE == B_ROW) return __ldg(p.B + (size_t)k * p.ldb + n);
    if (BMODE == B_COL) return __ldg(p.B + (size_t)n * p.ldb + k);
    const int tap = k / CDIM, co = k - tap * CDIM;
    return __ldg(p.B + ((size_t)(tap * p.N + n)) * CDIM + co);
}

// ---- 16-byte operand loads along the contiguous dimension of each loader mode ----------------------------------------
// K-vector: elements (m, k .. k+3) of A / (k .. k+3, n) of B; MN-vector: (m .. m+3, k) / (k, n .. n+3).  k and m / n are
// multiples of 4 by construction.  One address computation and bounds test per four elements instead of per element, and a
// warp's request covers whole sectors: the scalar loaders kept the load / store unit, not the FMA pipe, busy (ncu: MIO
// throttle and short-scoreboard stalls first, L1 at 81 %).  Anything irregular (partial vector at an edge, unaligned base)
// falls back to the element-wise loaders, so every shape stays legal.
template <int AMODE, int CDIM> struct AVec {
    static constexpr bool K = AMODE == A_ROW || ((AMODE == A_IM2COL || AMODE == A_IM2COL_FLIP) && CDIM % 4 == 0);
    static constexpr bool M = AMODE == A_COL || (AMODE == A_IM2COL_T && CDIM % 4 == 0);
};
template <int BMODE, int CDIM> struct BVec {
    static constexpr bool N = BMODE == B_ROW;
    static constexpr bool K = BMODE == B_COL || (BMODE == B_CONVW_T && CDIM % 4 == 0);
};
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int AMODE, int CDIM>
__device__ __forceinline__ float4 load_a_kvec(const GemmParams& p, int m, int k, int k_end) {
    if (m < p.M && k + 3 < k_end) {
        const float* ptr = nullptr;
        bool zero = false;
        if (AMODE == A_ROW) ptr = p.A + (size_t)m * p.lda + k;
        else {
            const int b = m / OBS_PIX, q = m - b * OBS_PIX, i = q / OBS_HW, j = q - i * OBS_HW;
            const int tap = k / CDIM, c = k - tap * CDIM;                 // c .. c+3 lie in the same tap (CDIM % 4 == 0)
            const int di = tap / 3 - 1, dj = tap - (tap / 3) * 3 - 1;
            const int ii = (AMODE == A_IM2COL_FLIP) ? i - di : i + di, jj = (AMODE == A_IM2COL_FLIP) ? j - dj : j + dj;
            zero = (unsigned)ii >= (unsigned)OBS_HW || (unsigned)jj >= (unsigned)OBS_HW;
            ptr = p.A + ((size_t)(b * OBS_PIX + ii * OBS_HW + jj)) * CDIM + c;
        }
        if (zero) return make_float4(0.f, 0.f, 0.f, 0.f);
        if (aligned16(ptr)) return __ldg(reinterpret_cast<const float4*>(ptr));
    }
    float4 v;
    v.x = k < k_end ? load_a<AMODE, CDIM>(p, m, k) : 0.f;
    v.y = k + 1 < k_end ? load_a<AMODE, CDIM>(p, m, k + 1) : 0.f;
    v.z = k + 2 < k_end ? load_a<AMODE, CDIM>(p, m, k + 2) : 0.f;
    v.w = k + 3 < k_end ? load_a<AMODE, CDIM>(p, m, k + 3) : 0.f;
    return v;
}
template <int AMODE, int CDIM>
__device__ __forceinline__ float4 load_a_mvec(const GemmParams& p, int m, int k, int k_end) {
    if (k >= k_end) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (m + 3 < p.M) {
        const float* ptr = nullptr;
        bool zero = false;
        if (AMODE == A_COL) ptr = p.A + (size_t)k * p.lda + m;
        else {                                                           // A_IM2COL_T: m = (tap, c .. c+3), k = pixel
            const int b = k / OBS_PIX, q = k - b * OBS_PIX, i = q / OBS_HW, j = q - i * OBS_HW;
            const int tap = m / CDIM, c = m - tap * CDIM;
            const int ii = i + tap / 3 - 1, jj = j + (tap - (tap / 3) * 3) - 1;
            zero = (unsigned)ii >= (unsigned)OBS_HW || (unsigned)jj >= (unsigned)OBS_HW;
            ptr = p.A + ((size_t)(b * OBS_PIX + ii * OBS_HW + jj)) * CDIM + c;
        }
        if (zero) return make_float4(0.f, 0.f, 0.f, 0.f);
        if (aligned16(ptr)) return __ldg(reinterpret_cast<const float4*>(ptr));
    }
    return make_float4(load_a<AMODE, CDIM>(p, m, k), load_a<AMODE, CDIM>(p, m + 1, k), load_a<AMODE, CDIM>(p, m + 2, k),
                       load_a<AMODE, CDIM>(p, m + 3, k));
}
template <int BMODE, int CDIM>
__device__ __forceinline__ float4 load_b_nvec(const GemmParams& p, int k, int n, int k_end) {
    if (k >= k_end) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (n + 3 < p.N) {
        const float* ptr = p.B + (size_t)k * p.ldb + n;                  // B_ROW
        if (aligned16(ptr)) return __ldg(reinterpret_cast<const float4*>(ptr));
    }
    return make_float4(load_b<BMODE, CDIM>(p, k, n), load_b<BMODE, CDIM>(p, k, n + 1), load_b<BMODE, CDIM>(p, k, n + 2),
                       load_b<BMODE, CDIM>(p, k, n + 3));
}
template <int BMODE, int CDIM>
__device__ __forceinline__ float4 load_b_kvec(const GemmParams& p, int k, int n, int k_end) {
    if (n < p.N && k + 3 < k_end) {
        const float* ptr;
        if (BMODE == B_COL) ptr = p.B + (size_t)n * p.ldb + k;
        else { const int tap = k / CDIM, co = k - tap * CDIM; ptr = p.B + ((size_t)(tap * p.N + n)) * CDIM + co; }
        if (aligned16(ptr)) return __ldg(reinterpret_cast<const float4*>(ptr));
    }
    float4 v;
    v.x = k < k_end ? load_b<BMODE, CDIM>(p, k, n) : 0.f;
    v.y = k + 1 < k_end ? load_b<BMODE, CDIM>(p, k + 1, n) : 0.f;
    v.z = k + 2 < k_end ? load_b<BMODE, CDIM>(p, k + 2, n) : 0.f;
    v.w = k + 3 < k_end ? load_b<BMODE, CDIM>(p, k + 3, n) : 0.f;
    return v;
}

// (m, n) = kernel coordinates; with trans_out the output element is C[n][m]
__device__ __forceinline__ size_t out_index(const GemmParams& p, int m, int n) {
    return p.trans_out ? (size_t)n * p.ldc + m : (size_t)m * p.ldc + n;
}
__device__ __forceinline__ float epilogue_value(const GemmParams& p, float v, int m, int n) {
    if (p.trans_out) { const int t = m; m = n; n = t; }
    if (p.bias) v += __ldg(p.bias + n);
    if (p.relu) v = fmaxf(v, 0.f);
    if (p.mask_act) v = (__ldg(p.mask_act + (size_t)m * p.ldc + n) > 0.f) ? v : 0.f;
    if (p.drop) v = __ldg(p.drop + (size_t)m * (p.drop_ld ? p.drop_ld : p.N) + n) ? v * p.drop_scale : 0.f;      // drop_ld = 0: a caller that fills GemmParams itself
    return v;
}

template <int AMODE, int BMODE, int BN, int CDIM>
// resident CTAs per SM the register allocation must allow: 2 at BN = 128 (<= 128 registers; two loader modes needed 130 / 139
// and ran ONE CTA per SM), 3 at BN = 64, 4 at BN = 32
__global__ void __launch_bounds__(GEMM_THREADS, BN >= 128 ? 2 : (BN >= 64 ? 3 : 4))
gemm_f32_kernel(GemmParams p) {
    constexpr int TN = BN / 16;                       // micro-tile columns per thread
    constexpr int B_PER_THREAD = GEMM_BK * BN / GEMM_THREADS;   // 8 / 4 / 2
    __shared__ __align__(16) float As[2][GEMM_BK][GEMM_BM + 4];
    __shared__ __align__(16) float Bs[2][GEMM_BK][BN + 4];

    const int t = threadIdx.x;
    const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * p.k_chunk;
    const int k_end = min(p.K, k_begin + p.k_chunk);
    const int tx = t & 15, ty = t >> 4;

    // global -> register staging assignment
    const int a_m = t & 127, a_k0 = (t >> 7) * 8;
    const int b_n = t % BN, b_k0 = (t / BN) * B_PER_THREAD;
    // micro-tile column j of thread tx: two interleaved groups of 4 for BN = 128 (conflict-free float4 reads)
    auto col_of = [&](int j) { return TN == 8 ? (j >> 2) * 64 + tx * 4 + (j & 3) : tx * TN + j; };
    using AV = AVec<AMODE, CDIM>;
    using BV = BVec<BMODE, CDIM>;
    constexpr bool A_VEC = AV::K || AV::M, B_VEC = BV::K || BV::N;
    constexpr int NB4 = BN >= 64 ? BN / 64 : 1;       // float4 of B per thread (BN = 32: the first 128 threads hold one)
    float ra[A_VEC ? 1 : 8], rb[B_VEC ? 1 : B_PER_THREAD];
    float4 ra4[A_VEC ? 2 : 1], rb4[B_VEC ? NB4 : 1];
    // accumulators as pairs of adjacent columns: the inner product issues packed FMAs (fma.rn.f32x2, two IEEE fp32 FMAs per
    // instruction: same results, half the issue slots of the FMA stream)
    float2 acc2[8][TN / 2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) acc2[i][j] = make_float2(0.f, 0.f);

    // vector staging assignments (f = t + 256 i numbers the float4 of a tile):
    //   A K-vectors: row = t >> 1, k-quad = (t & 1) + 2 i      (two lanes share a row: 32 contiguous bytes; the transposing
    //                scalar stores of a warp then hit 16 rows x 2 quads = 32 different banks)
    //   A M-vectors: k = (t >> 5) + 8 i, m-quad = t & 31       (one 16-byte shared store)
    //   B N-vectors: k = f / (BN/4), n-quad = f % (BN/4);  B K-vectors: n = (f % 2BN) >> 1, k-quad = (f & 1) + 2 (f / 2BN)
    auto fetch = [&](int k0) {
        if constexpr (AV::K) {
#pragma unroll
            for (int i = 0; i < 2; ++i) ra4[i] = load_a_kvec<AMODE, CDIM>(p, m0 + (t >> 1), k0 + ((t & 1) + 2 * i) * 4, k_end);
        } else if constexpr (AV::M) {
#pragma unroll
            for (int i = 0; i < 2; ++i) ra4[i] = load_a_mvec<AMODE, CDIM>(p, m0 + (t & 31) * 4, k0 + (t >> 5) + 8 * i, k_end);
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) ra[q] = (k0 + a_k0 + q < k_end) ? load_a<AMODE, CDIM>(p, m0 + a_m, k0 + a_k0 + q) : 0.f;
        }
        if constexpr (BV::N) {
#pragma unroll
            for (int i = 0; i < NB4; ++i) {
                const int f = t + GEMM_THREADS * i;
                if (f < 4 * BN) rb4[i] = load_b_nvec<BMODE, CDIM>(p, k0 + f / (BN / 4), n0 + (f % (BN / 4)) * 4, k_end);
            }
        } else if constexpr (BV::K) {
#pragma unroll
            for (int i = 0; i < NB4; ++i) {
                const int f = t + GEMM_THREADS * i;
                if (f < 4 * BN) rb4[i] = load_b_kvec<BMODE, CDIM>(p, k0 + ((f & 1) + 2 * (f / (2 * BN))) * 4, n0 + ((f % (2 * BN)) >> 1), k_end);
            }
        } else {
#pragma unroll
            for (int q = 0; q < B_PER_THREAD; ++q)
                rb[q] = (k0 + b_k0 + q < k_end) ? load_b<BMODE, CDIM>(p, k0 + b_k0 + q, n0 + b_n) : 0.f;
        }
    };
    auto stash = [&](int buf) {
        if constexpr (AV::K) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int kq = ((t & 1) + 2 * i) * 4, row = t >> 1;
                As[buf][kq][row] = ra4[i].x; As[buf][kq + 1][row] = ra4[i].y; As[buf][kq + 2][row] = ra4[i].z; As[buf][kq + 3][row] = ra4[i].w;
            }
        } else if constexpr (AV::M) {
#pragma unroll
            for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(&As[buf][(t >> 5) + 8 * i][(t & 31) * 4]) = ra4[i];
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) As[buf][a_k0 + q][a_m] = ra[q];
        }
        if constexpr (BV::N) {
#pragma unroll
            for (int i = 0; i < NB4; ++i) {
                const int f = t + GEMM_THREADS * i;
                if (f < 4 * BN) *reinterpret_cast<float4*>(&Bs[buf][f / (BN / 4)][(f % (BN / 4)) * 4]) = rb4[i];
            }
        } else if constexpr (BV::K) {
#pragma unroll
            for (int i = 0; i < NB4; ++i) {
                const int f = t + GEMM_THREADS * i;
                if (f < 4 * BN) {
                    const int kq = ((f & 1) + 2 * (f / (2 * BN))) * 4, n = (f % (2 * BN)) >> 1;
                    Bs[buf][kq][n] = rb4[i].x; Bs[buf][kq + 1][n] = rb4[i].y; Bs[buf][kq + 2][n] = rb4[i].z; Bs[buf][kq + 3][n] = rb4[i].w;
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < B_PER_THREAD; ++q) Bs[buf][b_k0 + q][b_n] = rb[q];
        }
    };

    int buf = 0;
    if (k_begin < k_end) { fetch(k_begin); stash(0); }
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += GEMM_BK) {
        const bool more = k0 + GEMM_BK < k_end;
        if (more) fetch(k0 + GEMM_BK);
#pragma unroll
        for (int kk = 0; kk < GEMM_BK; ++kk) {
            float a[8], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            if (TN >= 4) {
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][col_of(j)]);
                    b[j] = bv.x; b[j + 1] = bv.y; b[j + 2] = bv.z; b[j + 3] = bv.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][tx * TN + j];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 a2 = make_float2(a[i], a[i]);
#pragma unroll
                for (int j = 0; j < TN; j += 2) acc2[i][j / 2] = __ffma2_rn(a2, make_float2(b[j], b[j + 1]), acc2[i][j / 2]);
            }
        }
        if (more) { stash(buf ^ 1); }
        __syncthreads();
        buf ^= 1;
    }

    // Stores: four adjacent columns of the micro-tile per 16-byte store where the row length and the base allow it (the scalar
    // form issued 64 four-byte stores per thread, each touching a quarter of a sector).
    if (p.splits > 1) {
        float* dst = p.partial + (size_t)blockIdx.z * p.M * p.N;
        const bool vec = TN >= 4 && p.N % 4 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + ty * 8 + i;
            if (m >= p.M) continue;
            if (TN >= 4 && vec) {
#pragma unroll
                for (int j = 0; j + 3 < TN; j += 4) {
                    const int n = n0 + col_of(j);
                    if (n + 3 < p.N) *reinterpret_cast<float4*>(dst + (size_t)m * p.N + n) = make_float4(acc2[i][j / 2].x, acc2[i][j / 2].y, acc2[i][j / 2 + 1].x, acc2[i][j / 2 + 1].y);
                    else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) if (n + q < p.N) dst[(size_t)m * p.N + n + q] = (q & 1) ? acc2[i][(j + q) / 2].y : acc2[i][(j + q) / 2].x;
                    }
                }
                continue;
            }
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = n0 + col_of(j);
                if (n < p.N) dst[(size_t)m * p.N + n] = (j & 1) ? acc2[i][j / 2].y : acc2[i][j / 2].x;
            }
        }
    } else {
        const bool vec = TN >= 4 && !p.trans_out && p.ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.Cb) & 7) == 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + ty * 8 + i;
            if (m >= p.M) continue;
            if (TN >= 4 && vec) {
#pragma unroll
                for (int j = 0; j + 3 < TN; j += 4) {
                    const int n = n0 + col_of(j);
                    float v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = (n + q < p.N) ? epilogue_value(p, (q & 1) ? acc2[i][(j + q) / 2].y : acc2[i][(j + q) / 2].x, m, n + q) : 0.f;
                    const size_t o = (size_t)m * p.ldc + n;
                    if (n + 3 < p.N) {
                        *reinterpret_cast<float4*>(p.C + o) = make_float4(v[0], v[1], v[2], v[3]);
                        if (p.Cb) {
                            const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
                            *reinterpret_cast<uint2*>(p.Cb + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) if (n + q < p.N) { p.C[o + q] = v[q]; if (p.Cb) p.Cb[o + q] = __float2bfloat16(v[q]); }
                    }
                }
                continue;
            }
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = n0 + col_of(j);
                if (n < p.N) {
                    const float v = epilogue_value(p, (j & 1) ? acc2[i][j / 2].y : acc2[i][j / 2].x, m, n);
                    const size_t o = out_index(p, m, n);
                    p.C[o] = v;
                    if (p.Cb) p.Cb[o] = __float2bfloat16(v);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) splitk_epilogue_kernel(GemmParams p) {
    const size_t total = (size_t)p.M * p.N;
    // The partials of one output element are summed in split order (deterministic); the loads of eight splits are issued
    // together (16 / 32 at a time) — a loop of dependent load -> add pairs made the reduce of a small-batch layer (148 splits of a 32 x 512 output)
    // a 10 us latency chain.
    if (!p.trans_out && p.N % 4 == 0 && p.ldc % 4 == 0 && (((uintptr_t)p.C | (uintptr_t)p.partial) & 15) == 0 && ((uintptr_t)p.Cb & 7) == 0) {
        // four columns per thread: 16-byte loads of the partials, one pass
        const size_t total4 = total / 4;
        for (size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < total4; i4 += (size_t)gridDim.x * blockDim.x) {
            const size_t idx = i4 * 4;
            const int m = (int)(idx / p.N), n = (int)(idx - (size_t)m * p.N);
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            int s = 0;
            for (; s + 16 <= p.splits; s += 16) {                                         // fixed order: deterministic
                float4 w[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) w[u] = *reinterpret_cast<const float4*>(p.partial + (size_t)(s + u) * total + idx);
#pragma unroll
                for (int u = 0; u < 16; ++u) { a.x += w[u].x; a.y += w[u].y; a.z += w[u].z; a.w += w[u].w; }
            }
            for (; s < p.splits; ++s) {
                const float4 w = *reinterpret_cast<const float4*>(p.partial + (size_t)s * total + idx);
                a.x += w.x; a.y += w.y; a.z += w.z; a.w += w.w;
            }
            float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = epilogue_value(p, v[j], m, n + j);
            const size_t o = (size_t)m * p.ldc + n;
            if (p.C) *reinterpret_cast<float4*>(p.C + o) = make_float4(v[0], v[1], v[2], v[3]);
            if (p.Cb) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
                *reinterpret_cast<uint2*>(p.Cb + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
            }
        }
        return;
    }
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(idx / p.N), n = (int)(idx - (size_t)m * p.N);
        float v = 0.f;
        int s = 0;
        for (; s + 32 <= p.splits; s += 32) {                                             // fixed order: deterministic
            float w[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) w[u] = p.partial[(size_t)(s + u) * total + idx];
#pragma unroll
            for (int u = 0; u < 32; ++u) v += w[u];
        }
        for (; s + 8 <= p.splits; s += 8) {
            float w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) w[u] = p.partial[(size_t)(s + u) * total + idx];
#pragma unroll
            for (int u = 0; u < 8; ++u) v += w[u];
        }
        for (; s < p.splits; ++s) v += p.partial[(size_t)s * total + idx];
        v = epilogue_value(p, v, m, n);
        const size_t o = out_index(p, m, n);
        p.C[o] = v;
        if (p.Cb) p.Cb[o] = __float2bfloat16(v);
    }
}

// Launch helper.  `partial_cap` = floats available in p.partial.  Returns the number of kernels launched.
template <int AMODE, int BMODE, int BN, int CDIM>
inline int launch_gemm(GemmParams p, size_t partial_cap, int n_sms, cudaStream_t stream) {
    const int tiles_m = (p.M + GEMM_BM - 1) / GEMM_BM, tiles_n = (p.N + BN - 1) / BN;
    const int tiles = tiles_m * tiles_n;
    const int k_tiles = (p.K + GEMM_BK - 1) / GEMM_BK;
    // split-K so that the grid fills whole waves: `slots` CTAs are resident at once (launch bounds above), a grid of
    // tiles x splits CTAs runs in ceil(tiles x splits / slots) waves of which the last may be nearly empty.  Small grids take
    // the largest split that still fits ONE wave; grids of a few waves take the split (<= 4) with the fullest last wave.
    int splits = 1;
    const int slots = n_sms * (BN >= 128 ? 2 : (BN >= 64 ? 3 : 4));
    if (k_tiles >= 8) {
        if (tiles <= slots / 2) {
            splits = slots / tiles;
        } else if (tiles < 6 * slots) {
            double best = 0.0;
            for (int sp = 1; sp <= 4; ++sp) {
                const double waves = (double)tiles * sp / slots;
                const double eff = waves / (double)(long long)(waves + 0.999999) - 0.01 * (sp - 1);      // a split costs a partial round trip
                if (eff > best) { best = eff; splits = sp; }
            }
        }
        if (splits > k_tiles / 4) splits = k_tiles / 4;
        while (splits > 1 && (size_t)splits * p.M * p.N > partial_cap) --splits;
        if (splits < 1) splits = 1;
    }
    int chunk_tiles = (k_tiles + splits - 1) / splits;
    splits = (k_tiles + chunk_tiles - 1) / chunk_tiles;
    if (p.drop_ld == 0) p.drop_ld = p.trans_out ? p.M : p.N;
    p.splits = splits;
    p.k_chunk = chunk_tiles * GEMM_BK;
    dim3 grid(tiles_n, tiles_m, splits);
    gemm_f32_kernel<AMODE, BMODE, BN, CDIM><<<grid, GEMM_THREADS, 0, stream>>>(p);
    if (splits > 1) {
        size_t total = (size_t)p.M * p.N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 4 * n_sms) blocks = 4 * n_sms;
        splitk_epilogue_kernel<<<blocks, 256, 0, stream>>>(p);
        return 2;
    }
    return 1;
}

// Small-batch fully connected layers.  C[m][n] (m < M <= 32 batch rows) is computed as C^T = B^T A^T: the kernel's 128-row side
// runs over the layer's output features and the batch is the narrow BN = 32 side, so no FFMA is spent on padding rows (the
// 128 x BN tile on a 1- or 32-row batch spent 75-99 % of its FMAs there: fc1 forward 47 us at B = 1 and at B = 32).
//   wt_mode = A_ROW: the layer's weight is W[n][k] (forward, PyTorch Linear layout), x[m][k] row-major
//   wt_mode = A_COL: the weight is W[k][n] (data gradient: dX = dY W), x[m][k] row-major
// `p` is filled in ORIGINAL coordinates (M batch rows, N features, K, A = x / lda, B = W / ldb, C / ldc, bias, mask_act, drop).
template <int WT_MODE>
inline int launch_gemm_swapped(GemmParams p, size_t partial_cap, int n_sms, cudaStream_t stream) {
    GemmParams q = p;
    q.M = p.N; q.N = p.M;
    q.A = p.B; q.lda = p.ldb;
    q.B = p.A; q.ldb = p.lda;
    q.trans_out = 1; q.drop_ld = p.N;
    return launch_gemm<WT_MODE, B_COL, 32, 1>(q, partial_cap, n_sms, stream);
}
constexpr int GEMM_SWAP_MAX_M = 32;

}  // namespace mq
