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
