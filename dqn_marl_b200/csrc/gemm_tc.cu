// C-ABI entry of the stand-alone tensor-core GEMM (used by tests and by bench.py's tensor-pipe roofline leg);
// the Q-network's bf16 path calls mq::tc::launch directly.
#include <new>
#include "common.h"
#include "gemm_tc.cuh"

namespace mq {
__global__ void __launch_bounds__(256) tc_splitk_reduce_kernel(const float* __restrict__ partial, int splits, size_t total, float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        float v = 0.f;
        for (int s = 0; s < splits; ++s) v += partial[(size_t)s * total + i];
        out[i] = v;
    }
}
}  // namespace mq

// C[M][N] (fp32, row-major) = A[M][K] (bf16, K contiguous) * B[N][K]^T (bf16, K contiguous); fp32 accumulation in TMEM.
// bn = 128 / 64 / 32 selects the tile width; splits > 1 needs workspace >= splits*M*N floats.
extern "C" int mq_gemm_bf16(const void* A, const void* B, float* C, int32_t M, int32_t N, int32_t K, int32_t bn, int32_t splits,
                            float* workspace, void* stream) {
    MQ_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "mq_gemm_bf16: bad argument");
    MQ_REQUIRE(K % 8 == 0, "mq_gemm_bf16: K must be a multiple of 8 (TMA row pitch of 16 bytes)");
    MQ_REQUIRE(splits <= 1 || workspace, "mq_gemm_bf16: split-K needs a workspace");
    cudaStream_t s = (cudaStream_t)stream;
    mq::tc::Epilogue ep{};
    ep.out_f32 = C; ep.ldc = N; ep.partial = splits > 1 ? workspace : nullptr;
    int sp = splits < 1 ? 1 : splits;
    cudaError_t e;
    const __nv_bfloat16* a = (const __nv_bfloat16*)A;
    const __nv_bfloat16* b = (const __nv_bfloat16*)B;
    if (bn == 128) e = mq::tc::launch<128, 3>(a, K, b, K, M, N, K, ep, &sp, s);
    else if (bn == 64) e = mq::tc::launch<64, 4>(a, K, b, K, M, N, K, ep, &sp, s);
    else if (bn == 32) e = mq::tc::launch<32, 4>(a, K, b, K, M, N, K, ep, &sp, s);
    else return mq::fail(MQ_ERR_ARG, "mq_gemm_bf16: bn must be 128, 64 or 32");
    if (e != cudaSuccess) return mq::fail(MQ_ERR_CUDA, "mq_gemm_bf16: launch failed: %s", cudaGetErrorString(e));
    if (sp > 1) {
        size_t total = (size_t)M * N;
        int blocks = (int)((total + 255) / 256); if (blocks > 1184) blocks = 1184;
        mq::tc_splitk_reduce_kernel<<<blocks, 256, 0, s>>>(workspace, sp, total, C);
        MQ_CUDA(cudaGetLastError());
    }
    return MQ_OK;
}
