// C-ABI entry of the stand-alone tensor-core GEMM (used by tests and by bench.py's tensor-pipe roofline leg);
// the Q-network's bf16 path calls mq::tc::launch directly.
#include <new>
#include <cstdlib>
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

// C[M][N] (fp32 in C and / or bf16 in C_bf16, row-major) = A[M][K] (bf16, K contiguous) * B[N][K]^T (bf16, K contiguous); fp32
// accumulation in TMEM.  A bf16-only output with N a multiple of the tile width goes through the TMA-store epilogue.
// bn = 128 / 64 / 32 selects the tile width; splits > 1 needs workspace >= splits*M*N floats.
extern "C" int mq_gemm_bf16(const void* A, const void* B, float* C, void* C_bf16, int32_t M, int32_t N, int32_t K, int32_t bn, int32_t splits,
                            float* workspace, void* stream) {
    MQ_REQUIRE(A && B && (C || C_bf16) && M > 0 && N > 0 && K > 0, "mq_gemm_bf16: bad argument");
    MQ_REQUIRE(C || splits <= 1, "mq_gemm_bf16: split-K needs the fp32 output");
    MQ_REQUIRE(K % 8 == 0, "mq_gemm_bf16: K must be a multiple of 8 (TMA row pitch of 16 bytes)");
    MQ_REQUIRE(splits <= 1 || workspace, "mq_gemm_bf16: split-K needs a workspace");
    MQ_ON_DEVICE_OF(A);
    cudaStream_t s = (cudaStream_t)stream;
    mq::tc::Epilogue ep{};
    ep.out_f32 = C; ep.out_bf16 = (__nv_bfloat16*)C_bf16; ep.ldc = N; ep.partial = splits > 1 ? workspace : nullptr;
    int sp = splits < 1 ? 1 : splits;
    cudaError_t e;
    const __nv_bfloat16* a = (const __nv_bfloat16*)A;
    const __nv_bfloat16* b = (const __nv_bfloat16*)B;
    if (bn == 2032) {                  // 2000 + bn: persistent kernel with the B operand resident (N == bn, K a few 64-blocks)
        MQ_REQUIRE(N == 32 && K % 64 == 0 && sp == 1, "mq_gemm_bf16: the resident-B form needs N == 32, K %% 64 == 0, no split");
        int dev = 0, n_sms = 0;
        MQ_CUDA(cudaGetDevice(&dev));
        MQ_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
        e = mq::tc::launch_resident<32, 8>(a, K, b, K, M, N, K, ep, n_sms, s);
    } else if (bn == 1256) e = mq::tc::launch_pair<256, 6>(a, K, b, K, M, N, K, ep, &sp, s);       // 1000 + bn: CTA pairs (cta_group::2), 256 x bn per pair
    else if (bn == 1128) e = mq::tc::launch_pair<128, 8>(a, K, b, K, M, N, K, ep, &sp, s);
    else if (bn == 512) e = mq::tc::launch<256, 3, 2>(a, K, b, K, M, N, K, ep, &sp, s);         // 256 x 256 outputs per CTA
    else if (bn == 384) e = mq::tc::launch<128, 4, 2>(a, K, b, K, M, N, K, ep, &sp, s);    // 256 x 128
    else if (bn == 256) e = mq::tc::launch<256, 4>(a, K, b, K, M, N, K, ep, &sp, s);
    else if (bn == 128) e = mq::tc::launch<128, 3>(a, K, b, K, M, N, K, ep, &sp, s);
    else if (bn == 64) e = mq::tc::launch<64, 4>(a, K, b, K, M, N, K, ep, &sp, s);
    else if (bn == 32) e = mq::tc::launch<32, 4>(a, K, b, K, M, N, K, ep, &sp, s);
    else return mq::fail(MQ_ERR_ARG, "mq_gemm_bf16: bn must be 32, 64, 128, 256 (128-row tiles), 384, 512 (256 x 128 / 256 x 256 tiles) 1128, 1256 (CTA pairs) or 2032 (resident B)");
    if (e != cudaSuccess) return mq::fail(MQ_ERR_CUDA, "mq_gemm_bf16: launch failed: %s", cudaGetErrorString(e));
    if (sp > 1) {
        size_t total = (size_t)M * N;
        int blocks = (int)((total + 255) / 256); if (blocks > 1184) blocks = 1184;
        mq::tc_splitk_reduce_kernel<<<blocks, 256, 0, s>>>(workspace, sp, total, C);
        MQ_CUDA(cudaGetLastError());
    }
    return MQ_OK;
}

static int finish_split(float* workspace, int sp, size_t total, float* C, cudaStream_t s) {
    if (sp > 1) {
        int blocks = (int)((total + 255) / 256); if (blocks > 1184) blocks = 1184;
        mq::tc_splitk_reduce_kernel<<<blocks, 256, 0, s>>>(workspace, sp, total, C);
        MQ_CUDA(cudaGetLastError());
    }
    return MQ_OK;
}

// C[M][N] (fp32) = At[K][M]^T * Bt[K][N]: both operands bf16 with K as the ROW index (MN-major UMMA descriptors) —
// the weight-gradient shapes (K = batch rows) without a transpose.  M, N multiples of 8.
extern "C" int mq_gemm_bf16_tn(const void* At, const void* Bt, float* C, int32_t M, int32_t N, int32_t K, int32_t splits, float* workspace,
                               void* stream) {
    MQ_REQUIRE(At && Bt && C && M > 0 && N > 0 && K > 0, "mq_gemm_bf16_tn: bad argument");
    MQ_REQUIRE(M % 8 == 0 && N % 8 == 0, "mq_gemm_bf16_tn: M and N must be multiples of 8 (TMA row pitch of 16 bytes)");
    MQ_REQUIRE(splits <= 1 || workspace, "mq_gemm_bf16_tn: split-K needs a workspace");
    MQ_ON_DEVICE_OF(At);
    cudaStream_t s = (cudaStream_t)stream;
    mq::tc::Epilogue ep{};
    ep.out_f32 = C; ep.ldc = N; ep.partial = splits > 1 ? workspace : nullptr;
    int sp = splits < 1 ? 1 : splits;
    cudaError_t e = mq::tc::launch_tn<128, 3>((const __nv_bfloat16*)At, M, (const __nv_bfloat16*)Bt, N, M, N, K, ep, &sp, s);
    if (e != cudaSuccess) return mq::fail(MQ_ERR_CUDA, "mq_gemm_bf16_tn: launch failed: %s", cudaGetErrorString(e));
    return finish_split(workspace, sp, (size_t)M * N, C, s);
}

// Y[B*121][Cout] (fp32 in Y and / or bf16 in Y_bf16) = conv3x3/pad1 of X [B][11][11][Cin] (bf16, NHWC) with Wk[Cout][9*Cin] (bf16, taps (kh,kw,c)); implicit
// GEMM, the im2col matrix is never written: each tap is a shifted, zero-filled 4-D TMA box.  flip = 1 mirrors the taps
// (data gradient).  Cin must be a multiple of 32 (32 -> 64B swizzle, else 128B swizzle); bn = 128 / 64 / 32 = tile width of the
// one-sample-per-CTA kernel, bn = 0 = persistent kernel with the weights resident in shared memory.
extern "C" int mq_conv3x3_bf16(const void* X, const void* Wk, float* Y, void* Y_bf16, int64_t batch, int32_t Cin, int32_t Cout, int32_t flip,
                               int32_t bn, void* stream) {
    MQ_REQUIRE(X && Wk && (Y || Y_bf16) && batch > 0 && Cin > 0 && Cout > 0, "mq_conv3x3_bf16: bad argument");
    MQ_REQUIRE(Cin % 32 == 0, "mq_conv3x3_bf16: Cin must be a multiple of 32");
    MQ_ON_DEVICE_OF(X);
    cudaStream_t s = (cudaStream_t)stream;
    mq::tc::Epilogue ep{};
    ep.out_f32 = Y; ep.out_bf16 = (__nv_bfloat16*)Y_bf16; ep.ldc = Cout;
    const __nv_bfloat16* x = (const __nv_bfloat16*)X;
    const __nv_bfloat16* w = (const __nv_bfloat16*)Wk;
    cudaError_t e;
    if (bn == 0) {          // persistent kernel: weights resident in shared memory, Cout = tile width
        int dev = 0, n_sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
        if (Cin % 64 == 0 && Cout == 128) e = mq::tc::launch_conv_persistent<128, 64, 3>(x, w, batch, Cin, Cout, flip, ep, n_sms, s);
        else if (Cin % 64 == 0 && Cout == 64) e = mq::tc::launch_conv_persistent<64, 64, 3>(x, w, batch, Cin, Cout, flip, ep, n_sms, s);
        else if (Cin % 64 == 0 && Cout == 32) e = mq::tc::launch_conv_persistent<32, 64, 4>(x, w, batch, Cin, Cout, flip, ep, n_sms, s);
        else if (Cout == 64) e = mq::tc::launch_conv_persistent<64, 32, 6>(x, w, batch, Cin, Cout, flip, ep, n_sms, s);
        else return mq::fail(MQ_ERR_ARG, "mq_conv3x3_bf16: the persistent kernel (bn = 0) supports Cout = 128/64/32 (Cin %% 64 == 0) or Cout = 64");
    } else if (Cin % 64 == 0) {
        if (bn == 128) e = mq::tc::launch_conv<128, 4, 64>(x, w, batch, Cin, Cout, flip, ep, s);
        else if (bn == 64) e = mq::tc::launch_conv<64, 4, 64>(x, w, batch, Cin, Cout, flip, ep, s);
        else if (bn == 32) e = mq::tc::launch_conv<32, 4, 64>(x, w, batch, Cin, Cout, flip, ep, s);
        else return mq::fail(MQ_ERR_ARG, "mq_conv3x3_bf16: bn must be 128, 64 or 32");
    } else {
        if (bn == 64) e = mq::tc::launch_conv<64, 6, 32>(x, w, batch, Cin, Cout, flip, ep, s);
        else return mq::fail(MQ_ERR_ARG, "mq_conv3x3_bf16: Cin = 32 (mod 64) supports bn = 64 only");
    }
    if (e != cudaSuccess) return mq::fail(MQ_ERR_CUDA, "mq_conv3x3_bf16: launch failed: %s", cudaGetErrorString(e));
    return MQ_OK;
}

// dW[9*Cin][Cout] (fp32) = im2col(X)^T * dY: X [B][11][11][Cin] bf16, dY [B*121][Cout] bf16; Cin % 32 == 0, Cout % 8 == 0.
// splits > 1 partitions the samples; workspace >= splits * 9*Cin * Cout floats.
// dBias (optional): [Cout] column sums of dY (the bias gradient), produced by the same kernel through a spare A row of ones;
// needs 9*Cin not to be a multiple of 128 (MQ_ERR_UNSUPPORTED otherwise) and splits * Cout more workspace floats.
extern "C" int mq_conv3x3_wgrad_bf16(const void* X, const void* dY, float* dW, float* dBias, int64_t batch, int32_t Cin, int32_t Cout,
                                     int32_t splits, float* workspace, void* stream) {
    MQ_REQUIRE(X && dY && dW && batch > 0, "mq_conv3x3_wgrad_bf16: bad argument");
    MQ_REQUIRE(Cin % 32 == 0 && Cout % 8 == 0, "mq_conv3x3_wgrad_bf16: Cin must be a multiple of 32 and Cout of 8");
    MQ_REQUIRE(splits <= 1 || workspace, "mq_conv3x3_wgrad_bf16: split needs a workspace");
    MQ_REQUIRE(!dBias || workspace, "mq_conv3x3_wgrad_bf16: the bias gradient needs a workspace");
    MQ_ON_DEVICE_OF(X);
    cudaStream_t s = (cudaStream_t)stream;
    mq::tc::Epilogue ep{};
    int sp = splits < 1 ? 1 : splits;
    if (sp > batch) sp = (int)batch;
    const size_t wtotal = (size_t)9 * Cin * Cout;
    ep.out_f32 = dW; ep.ldc = Cout; ep.partial = sp > 1 ? workspace : nullptr;
    ep.colsum_partial = dBias ? workspace + (sp > 1 ? (size_t)sp * wtotal : 0) : nullptr;       // behind the (at most sp) weight partials
    bool fused = false;
    cudaError_t e;
    const __nv_bfloat16* x = (const __nv_bfloat16*)X;
    const __nv_bfloat16* dy = (const __nv_bfloat16*)dY;
    if (Cin % 64 != 0) {
        MQ_REQUIRE(Cout <= 64, "mq_conv3x3_wgrad_bf16: Cin = 32 (mod 64) supports Cout <= 64 only");
        e = mq::tc::launch_conv_wgrad<64, 4, 32>(x, dy, batch, Cin, Cout, ep, &sp, s, &fused);
    } else if (Cout > 64) e = mq::tc::launch_conv_wgrad<128, 3, 64>(x, dy, batch, Cin, Cout, ep, &sp, s, &fused);
    else e = mq::tc::launch_conv_wgrad<64, 4, 64>(x, dy, batch, Cin, Cout, ep, &sp, s, &fused);
    if (e != cudaSuccess) return mq::fail(MQ_ERR_CUDA, "mq_conv3x3_wgrad_bf16: launch failed: %s", cudaGetErrorString(e));
    if (dBias) {
        if (!fused) return mq::fail(MQ_ERR_UNSUPPORTED, "mq_conv3x3_wgrad_bf16: no spare operand row for the bias gradient (9*Cin %% 128 == 0)");
        mq::tc_splitk_reduce_kernel<<<1, 256, 0, s>>>(ep.colsum_partial, sp, (size_t)Cout, dBias);      // sp >= 1 rows of [Cout]
        MQ_CUDA(cudaGetLastError());
    }
    return finish_split(workspace, sp, wtotal, dW, s);
}

#ifdef MQ_CONV_TRACE
extern "C" int mq_debug_conv_trace(long long* host_out, int reset) {
    if (reset) { static long long z[256 * 8]; return (int)cudaMemcpyToSymbol(mq::tc::g_conv_trace, z, sizeof(z)); }
    return (int)cudaMemcpyFromSymbol(host_out, mq::tc::g_conv_trace, sizeof(long long) * 256 * 8);
}
#endif
