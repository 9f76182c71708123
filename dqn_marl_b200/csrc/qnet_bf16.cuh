// Helper kernels of the Q-network's bf16 tensor-core path.  Activations need no preparation any more (the convolutions
// are implicit GEMMs over shifted TMA boxes and the weight gradients use MN-major operands, gemm_tc.cuh); what is left:
//
//   cast / weight re-layouts  fp32 master weights (kernel layouts of qnet.cu) -> bf16 GEMM operands, once per optimizer step
//   colsum_partial_bf16       bias gradients
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

// conv1 on the tensor cores: the 6-channel observation is too narrow for a TMA box (12-byte pixels), so its im2col rows
// A1 bf16 [B*121][64] — column k = tap*6 + c for k < 54, 1.0 in column 54 so that the weight-gradient GEMM yields the bias
// gradient as row 54, zero above — are built in shared memory by tc::conv1_obs_resident_kernel (gemm_tc.cuh).
// conv1 weights Wc[(tap*6 + c)][32] f32 -> GEMM operand W1f[32][64] bf16 (K padded 54 -> 64 with zeros)
__global__ void __launch_bounds__(256)
conv1_weight_bf16_kernel(const float* __restrict__ wc, bf16* __restrict__ w1f) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 32 * 64) return;
    const int n = idx >> 6, k = idx & 63;
    w1f[idx] = __float2bfloat16(k < 54 ? wc[k * 32 + n] : 0.f);
}
// conv1 weight gradient: partial[split][n (32)][k (64)] (the TN GEMM computes dY^T A1) -> dWc[(tap*6 + c)][32], fixed order
__global__ void __launch_bounds__(256)
conv1_wgrad_reduce_kernel(const float* __restrict__ partial, int splits, float* __restrict__ dwc, float* __restrict__ dbias) {
    // one CTA per k (54 weight rows + the ones column k = 54 -> bias gradient): thread = (n, split group of 8); the groups are
    // combined through shared memory in a fixed order
    __shared__ float red[8][32];
    const int k = blockIdx.x, n = threadIdx.x & 31, sg = threadIdx.x >> 5;
    float acc = 0.f;
    for (int s = sg; s < splits; s += 8) acc += partial[((size_t)s * 32 + n) * 64 + k];
    red[sg][n] = acc;
    __syncthreads();
    if (sg == 0) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += red[g][n];
        if (k < 54) dwc[k * 32 + n] = t;
        else dbias[n] = t;
    }
}

// f32 [R][C] -> bf16 [R][C] and (optionally) bf16 [C][R]
__global__ void __launch_bounds__(256)
cast_transpose_kernel(const float* __restrict__ src, bf16* __restrict__ dst, bf16* __restrict__ dst_t, long long R, int C) {
    // 32 x 32 tiles through shared memory: both the row-major and the transposed copy are written coalesced
    __shared__ float tile[32][33];
    const int tiles_c = (C + 31) / 32;
    const long long r0 = (long long)(blockIdx.x / tiles_c) * 32;
    const int c0 = (blockIdx.x % tiles_c) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long r = r0 + ty + 8 * k;
        const int c = c0 + tx;
        float v = 0.f;
        if (r < R && c < C) {
            v = src[r * C + c];
            if (dst) dst[r * C + c] = __float2bfloat16(v);
        }
        tile[ty + 8 * k][tx] = v;
    }
    if (!dst_t) return;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k;
        const long long r = r0 + tx;
        if (r < R && c < C) dst_t[(long long)c * R + r] = __float2bfloat16(tile[tx][ty + 8 * k]);
    }
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
