// bf16 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ->
// tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16) with the fp32 accumulator in TMEM ->
// tcgen05.ld -> fused epilogue -> global.   Hand-written PTX, no CUTLASS.
//
//   C[M][N] = A[M][K] * B[N][K]^T          A, B bf16, both K-major (row-major with K contiguous)
//
// This is the throughput path of the Q-network (DQNNetwork, Louvre_Evacuation/agents/dqn_agent.py:15-61): the
// contractions with >= 16k rows (conv2/conv3 as GEMMs over im2col rows, fc1) run here; the fp32 FFMA kernels of
// gemm_f32.cuh remain the 1e-5 parity path.  Warp roles per CTA (192 threads): warp 0 = TMA producer (one lane),
// warp 1 = TMEM allocator + MMA issuer (one lane), warps 2..5 = epilogue (each owns 32 TMEM lanes).
// One 128 x BN output tile per CTA, optional split-K over gridDim.z (fp32 partials).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mq {
namespace tc {

constexpr int BM = 128, BK = 64, UMMA_K = 16, THREADS = 192;

struct Epilogue {
    float* out_f32;              // [M][ldc] or null
    __nv_bfloat16* out_bf16;     // [M][ldc] or null
    int ldc;
    const float* bias;           // [N] or null
    int relu;
    const __nv_bfloat16* mask_bf16;   // multiply by (mask[m][n] > 0), same ld as out, or null
    const float* mask_f32;
    const uint8_t* drop;         // [M][N] keep mask or null
    float drop_scale;
    float* partial;              // split-K: [splits][M][N] fp32 (then no other epilogue op is applied)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer) : "memory");
}
// shared-memory matrix descriptor: K-major tile of [rows][64 bf16] written by TMA with SWIZZLE_128B
// (8-row x 128-byte swizzle atoms, 1024 B apart)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address
    d |= (uint64_t)0 << 16;                                // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;                     // stride byte offset: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN, int STAGES>
struct SmemLayout {
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(THREADS)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M, int N, int K,
                    int k_chunk, Epilogue ep) {
    extern __shared__ unsigned char smem_raw[];
    using L = SmemLayout<BN, STAGES>;
    unsigned char* tiles = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);     // SWIZZLE_128B: 1024 B alignment
    uint64_t* full_bar = (uint64_t*)(tiles + STAGES * L::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_ptr = (uint32_t*)(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * k_chunk;
    const int k_end = min(K, k_begin + k_chunk);
    const int num_kb = (k_end - k_begin + BK - 1) / BK;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    }
    if (warp == 1) {      // TMEM allocation by one full warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t phase = 1;          // fresh barriers: the first pass over the ring does not wait
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&empty_bar[s], phase);
                unsigned char* a_dst = tiles + s * L::STAGE_BYTES;
                unsigned char* b_dst = a_dst + L::A_BYTES;
                mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
                const int k = k_begin + kb * BK;
                tma_load_2d(a_dst, &tmap_a, &full_bar[s], k, m0);
                tma_load_2d(b_dst, &tmap_b, &full_bar[s], k, n0);
                if (s == STAGES - 1) phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&full_bar[s], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(tiles + s * L::STAGE_BYTES);
                const uint32_t b_addr = a_addr + L::A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t adesc = make_smem_desc(a_addr + k * UMMA_K * 2);
                    const uint64_t bdesc = make_smem_desc(b_addr + k * UMMA_K * 2);
                    umma_bf16(tmem_base, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);          // frees the stage once these MMAs have read it
                if (s == STAGES - 1) phase ^= 1;
            }
            umma_commit(tmem_full_bar);              // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
        const int q = warp & 3;
        mbar_wait(tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int m = m0 + q * 32 + lane;
        // fast path: no per-element masks, full 32-column chunks, 16-byte aligned rows -> vector stores
        const bool plain = !ep.mask_bf16 && !ep.mask_f32 && !ep.drop && (ep.ldc % 8 == 0);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            if (m >= M || num_kb <= 0) continue;
            const int nb = n0 + c0;
            if (nb >= N) continue;
            if (ep.partial) {
                float* dst = ep.partial + ((size_t)blockIdx.z * M + m) * N + nb;
                if (nb + 32 <= N && (N % 4 == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(dst + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (nb + j < N) dst[j] = __uint_as_float(r[j]);
                }
                continue;
            }
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (ep.bias) {
#pragma unroll
                for (int j = 0; j < 32; ++j) if (nb + j < N) v[j] += __ldg(ep.bias + nb + j);
            }
            if (ep.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            const size_t o = (size_t)m * ep.ldc + nb;
            if (plain && nb + 32 <= N) {
                if (ep.out_bf16) {
                    uint4* dst = reinterpret_cast<uint4*>(ep.out_bf16 + o);
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]), p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                        dst[j / 8] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                                                *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
                    }
                }
                if (ep.out_f32) {
                    float4* dst = reinterpret_cast<float4*>(ep.out_f32 + o);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) dst[j / 4] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            } else {
                if (ep.mask_bf16 && nb + 32 <= N && (ep.ldc % 8 == 0)) {
                    const uint4* mk = reinterpret_cast<const uint4*>(ep.mask_bf16 + o);
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const uint4 w = mk[j / 8];
                        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&ws[k]);
                            if (!(__low2float(b2) > 0.f)) v[j + 2 * k] = 0.f;
                            if (!(__high2float(b2) > 0.f)) v[j + 2 * k + 1] = 0.f;
                        }
                    }
                } else if (ep.mask_bf16) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (nb + j < N && !(__bfloat162float(ep.mask_bf16[o + j]) > 0.f)) v[j] = 0.f;
                }
                if (ep.mask_f32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (nb + j < N && !(ep.mask_f32[o + j] > 0.f)) v[j] = 0.f;
                }
                if (ep.drop) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (nb + j < N) v[j] = ep.drop[(size_t)m * N + nb + j] ? v[j] * ep.drop_scale : 0.f;
                }
                if (nb + 32 <= N && (ep.ldc % 8 == 0)) {
                    if (ep.out_bf16) {
                        uint4* dst = reinterpret_cast<uint4*>(ep.out_bf16 + o);
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]), p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                            __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                            dst[j / 8] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                                                    *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
                        }
                    }
                    if (ep.out_f32) {
                        float4* dst = reinterpret_cast<float4*>(ep.out_f32 + o);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) dst[j / 4] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (nb + j >= N) continue;
                        if (ep.out_f32) ep.out_f32[o + j] = v[j];
                        if (ep.out_bf16) ep.out_bf16[o + j] = __float2bfloat16(v[j]);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2D bf16 row-major matrix [rows][cols] (cols contiguous, leading dimension ld elements), box = [box_rows][64]
inline bool make_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int STAGES>
inline cudaError_t launch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K, Epilogue ep,
                          int* splits_inout, cudaStream_t stream) {
    int splits = splits_inout ? *splits_inout : 1;
    CUtensorMap ta, tb;
    if (!make_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM) || !make_tmap(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN))
        return cudaErrorInvalidValue;
    using L = SmemLayout<BN, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int k_tiles = (K + BK - 1) / BK;
    if (splits < 1) splits = 1;
    if (splits > k_tiles) splits = k_tiles;
    const int chunk_tiles = (k_tiles + splits - 1) / splits;
    splits = (k_tiles + chunk_tiles - 1) / chunk_tiles;
    if (splits_inout) *splits_inout = splits;
    if (splits > 1 && !ep.partial) return cudaErrorInvalidValue;
    if (splits == 1) ep.partial = nullptr;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
    gemm_bf16_tc_kernel<BN, STAGES><<<grid, THREADS, L::TOTAL, stream>>>(ta, tb, M, N, K, chunk_tiles * BK, ep);
    return cudaGetLastError();
}

}  // namespace tc
}  // namespace mq
