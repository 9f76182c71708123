// Helper kernels of the Q-network's bf16 tensor-core path: operand preparation for gemm_tc.cuh, which wants both
// operands K-contiguous.  All of these are pure HBM streams (16-byte accesses), the flops live in the tcgen05 GEMMs.
//
//   im2col_bf16      X NHWC [B][11][11][C] (f32 or bf16) -> rows [B*121][9*C] bf16, taps in (kh, kw, c) order; `flip`
//                    mirrors the taps (conv dgrad reads dY[i-(kh-1)][j-(kw-1)])
//   transpose_bf16   [R][C] -> [C][R]
//   cast / weight re-layouts  fp32 master weights (kernel layouts of qnet.cu) -> bf16 GEMM operands
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mq {
namespace bf {

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }

// one thread = 8 consecutive channels of one (pixel, tap): a 16-byte store
template <typename Tin, int C>
__global__ void __launch_bounds__(256)
im2col_bf16_kernel(const Tin* __restrict__ src, bf16* __restrict__ dst, long long M, int flip) {
    constexpr int CH8 = C / 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = M * 9 * CH8;
    if (idx >= total) return;
    const int c8 = (int)(idx % CH8);
    const long long r = idx / CH8;
    const int tap = (int)(r % 9);
    const long long m = r / 9;
    const long long b = m / 121;
    const int q = (int)(m - b * 121), i = q / 11, j = q - i * 11;
    const int di = tap / 3 - 1, dj = tap % 3 - 1;
    const int ii = flip ? i - di : i + di, jj = flip ? j - dj : j + dj;
    __align__(16) bf16 v[8];
    if ((unsigned)ii < 11u && (unsigned)jj < 11u) {
        const Tin* s = src + ((b * 121 + ii * 11 + jj) * C + c8 * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __float2bfloat16(to_f<Tin>(s[k]));
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __float2bfloat16(0.f);
    }
    *reinterpret_cast<uint4*>(dst + (m * 9 + tap) * C + c8 * 8) = *reinterpret_cast<const uint4*>(v);
}

// [R][C] -> [C][R], 64 x 64 tiles through shared memory, 16-byte global accesses on both sides
// (C % 8 == 0 and R % 8 == 0 on the fast path; callers guarantee it: B is a multiple of 8)
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, long long R, int C) {
    __shared__ __align__(16) bf16 tile[64][72];
    const long long r0 = (long long)blockIdx.x * 64;
    const int c0 = blockIdx.y * 64;
    const bool vec = (C % 8 == 0) && (R % 8 == 0);
    for (int k = threadIdx.x; k < 64 * 8; k += 256) {       // 64 rows x 8 chunks of 8 elements
        const int r = k >> 3, c = (k & 7) * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r0 + r < R && c0 + c < C) {
            if (vec) v = *reinterpret_cast<const uint4*>(src + (r0 + r) * C + c0 + c);
            else {
                __align__(16) bf16 t[8];
                for (int q = 0; q < 8; ++q) t[q] = (c0 + c + q < C) ? src[(r0 + r) * C + c0 + c + q] : __float2bfloat16(0.f);
                v = *reinterpret_cast<const uint4*>(t);
            }
        }
        *reinterpret_cast<uint4*>(&tile[r][c]) = v;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 64 * 8; k += 256) {       // 64 output rows (columns of src) x 8 chunks of 8 source rows
        const int c = k >> 3, r = (k & 7) * 8;
        if (c0 + c >= C || r0 + r >= R) continue;
        __align__(16) bf16 t[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) t[q] = tile[r + q][c];
        bf16* d = dst + (long long)(c0 + c) * R + r0 + r;
        if (vec) *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(t);
        else for (int q = 0; q < 8; ++q) if (r0 + r + q < R) d[q] = t[q];
    }
}

// f32 [R][C] -> bf16 [R][C] and (optionally) bf16 [C][R]
__global__ void __launch_bounds__(256)
cast_transpose_kernel(const float* __restrict__ src, bf16* __restrict__ dst, bf16* __restrict__ dst_t, long long R, int C) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * C) return;
    const long long r = idx / C;
    const int c = (int)(idx - r * C);
    const bf16 v = __float2bfloat16(src[idx]);
    if (dst) dst[idx] = v;
    if (dst_t) dst_t[(long long)c * R + r] = v;
}

// conv weight re-layouts from the fp32 kernel layout Wc[(tap*Cin + c)][Cout]:
//   fwd  operand  Wf[n][(tap*Cin + c)]      (B operand of the forward GEMM, K = 9*Cin)
//   dgrad operand Wd[c][(tap*Cout + n)]     (B operand of the dgrad GEMM,   K = 9*Cout)
__global__ void __launch_bounds__(256)
conv_weight_bf16_kernel(const float* __restrict__ wc, bf16* __restrict__ wf, bf16* __restrict__ wd, int Cin, int Cout) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 9 * Cin * Cout) return;
    const int n = idx % Cout, tc = idx / Cout;
    const int tap = tc / Cin, c = tc - tap * Cin;
    const bf16 v = __float2bfloat16(wc[idx]);
    wf[(size_t)n * 9 * Cin + tc] = v;
    if (wd) wd[(size_t)c * 9 * Cout + tap * Cout + n] = v;
}

// column sums of a bf16 matrix (bias gradients), two deterministic stages like colsum_*_kernel in qnet.cu
__global__ void __launch_bounds__(256)
colsum_partial_bf16_kernel(const bf16* __restrict__ X, long long M, int N, int rows_per_block, float* __restrict__ partial) {
    const long long m0 = (long long)blockIdx.x * rows_per_block;
    const long long m1 = m0 + rows_per_block < M ? m0 + rows_per_block : M;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float acc = 0.f;
        for (long long m = m0; m < m1; ++m) acc += __bfloat162float(X[m * N + n]);
        partial[(size_t)blockIdx.x * N + n] = acc;
    }
}

}  // namespace bf
}  // namespace mq
