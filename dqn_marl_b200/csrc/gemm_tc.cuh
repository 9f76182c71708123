// bf16 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ->
// tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16) with the fp32 accumulator in TMEM ->
// tcgen05.ld -> fused epilogue -> global.   Hand-written PTX, no CUTLASS.
//
//   gemm_bf16_tc_kernel   C[M][N] = A[M][K] * B[N][K]^T     A, B bf16, both K-major (row-major with K contiguous)
//       CONV = true: A is never materialised.  It is the im2col view of an NHWC activation [B][11][11][C] of a 3x3 / pad 1
//       convolution: the K-block of tap (di, dj) is ONE 4-D TMA box {BK channels, 11, 11, 1 sample} whose origin is shifted
//       by (dj, di); the out-of-bounds part of the box is zero-filled by TMA = the padding.  One sample (121 rows, padded to
//       the 128-row MMA tile) per CTA.  `flip` mirrors the taps (dgrad).
//   gemm_bf16_tn_kernel   C[M][N] = At[K][M]^T * Bt[K][N]   both operands MN-major (K is the row index): the weight-gradient
//       shapes (K = batch rows) run without any transpose.  CONV = true: At is the shifted 4-D box view again (one sample =
//       one 128-row K-block whose last 7 rows stay zero), Bt = dY rows of the sample.
//
// This is the throughput path of the Q-network (DQNNetwork, Louvre_Evacuation/agents/dqn_agent.py:15-61): the
// contractions with >= 16k rows (conv2/conv3 as GEMMs over im2col rows, fc1) run here; the fp32 FFMA kernels of
// gemm_f32.cuh remain the 1e-5 parity path.  Warp roles per CTA (192 threads): warp 0 = TMA producer (one lane),
// warp 1 = TMEM allocator + MMA issuer (one lane), warps 2..5 = epilogue (each owns 32 TMEM lanes).
// One 128 x BN output tile per CTA, optional split-K over gridDim.z (fp32 partials).
#pragma once
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mq {
namespace tc {

constexpr int BM = 128, UMMA_K = 16, THREADS = 192;
constexpr int PIXELS = 121;      // 11 x 11 window of the Q-network's convolutions

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
    float* colsum_partial;       // conv weight gradient only: [splits][N] column sums of dY (= the bias gradient), or null
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
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// shared-memory matrix descriptor of a tile written by TMA with a 128B (layout 2) or 64B (layout 4) swizzle.
//   K-major  tile [rows][BK]: swizzle atoms of 8 rows, `sbo` bytes apart (8 * row bytes); lbo unused
//   MN-major tile [k rows][64 mn]: atoms of 8 k-rows x 128 B, `sbo` = 1024 B apart along K, `lbo` = distance between
//   the 64-wide MN slabs
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t sbo = 1024u, uint32_t lbo = 0u, uint64_t layout = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address
    d |= (uint64_t)((lbo & 0x3FFFFu) >> 4) << 16;          // leading byte offset
    d |= (uint64_t)((sbo & 0x3FFFFu) >> 4) << 32;          // stride byte offset
    d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
    d |= layout << 61;                                     // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    return d;
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, M = 128, N = n; mn_major sets both transpose bits
__host__ __device__ constexpr uint32_t make_idesc(int n, bool mn_major = false) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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

// ---- epilogue of one 128 x BN accumulator tile: TMEM -> registers -> fused ops -> global -----------------------------
// Executed by the four epilogue warps; q = TMEM lane quadrant of the warp.  `row` = global output row of this thread
// (valid iff row_ok), `n0` = first output column of the tile.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const Epilogue& ep, uint32_t tmem_base, int q, bool row_ok, long long row, int n0,
                                              long long M, int N, int split) {
    // fast path: no per-element masks, full 32-column chunks, 16-byte aligned rows -> vector stores
    const bool plain = !ep.mask_bf16 && !ep.mask_f32 && !ep.drop && (ep.ldc % 8 == 0);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        if (!row_ok) continue;
        const int nb = n0 + c0;
        if (nb >= N) continue;
        if (ep.partial) {
            float* dst = ep.partial + ((size_t)split * M + row) * N + nb;
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
            if (nb + 32 <= N) {         // nb is a multiple of 32: 16-byte aligned
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + nb + j));
                    v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) if (nb + j < N) v[j] += __ldg(ep.bias + nb + j);
            }
        }
        if (ep.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        const size_t o = (size_t)row * ep.ldc + nb;
        const bool vec = nb + 32 <= N && (ep.ldc % 8 == 0);
        if (!plain) {
            if (ep.mask_bf16 && vec) {
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
                for (int j = 0; j < 32; ++j) if (nb + j < N) v[j] = ep.drop[(size_t)row * N + nb + j] ? v[j] * ep.drop_scale : 0.f;
            }
        }
        if (vec) {
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

// ---- bf16 epilogue through shared memory + TMA store ----------------------------------------------------------------
// Direct stores from the TMEM register layout (one row per thread) hit 32 different sectors per instruction with 16 bytes
// each; the L2 write path, not the tensor pipe, then bounds the kernel.  Here a warp stages its 32 rows x 64 columns in a
// 4 KB SWIZZLE_128B tile (conflict-free 16-byte shared stores) and one lane issues a bulk tensor store.  Rows / columns
// outside the output map are clipped by TMA (conv: the map is 3-D {Cout, 121, batch}, so rows >= 121 of a sample vanish).
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// One 128 x BN accumulator: bias / relu / relu-mask, bf16, staged per warp, stored by TMA.
//   mask_row  this thread's row of ep.mask_bf16 at column n0 (null: no mask or row out of range)
//   (c_col, c_row, c_z)  coordinates of the warp's first element in the output map; c_z < 0: the map is 2-D
template <int BN>
__device__ __forceinline__ void epilogue_tile_tma(const Epilogue& ep, const CUtensorMap* tmap_out, uint32_t tmem_acc, int q, int lane,
                                                  const __nv_bfloat16* mask_row, int n0, int c_row, int c_z, unsigned char* stage) {
    static_assert(BN % 64 == 0, "TMA-store epilogue works on 64-column groups");
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 64) {
        uint32_t ra[32], rb[32];
        tmem_ld32_nowait(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, ra);
        tmem_ld32_nowait(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c0 + 32), rb);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t packed[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(h ? rb[j] : ra[j]);
            const int nb = c0 + 32 * h;
            if (ep.bias) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + nb + j));
                    v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                }
            }
            if (ep.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (mask_row) {
                const uint4* mk = reinterpret_cast<const uint4*>(mask_row + nb);
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    const uint4 w = __ldg(mk + j / 8);
                    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&ws[k]);
                        if (!(__low2float(b2) > 0.f)) v[j + 2 * k] = 0.f;
                        if (!(__high2float(b2) > 0.f)) v[j + 2 * k + 1] = 0.f;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j], v[j + 1]);
                packed[16 * h + j / 2] = *reinterpret_cast<const uint32_t*>(&p2);
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the previous store has read the tile
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            if (c_z >= 0) tma_store_3d(tmap_out, stage, n0 + c0, c_row, c_z);
            else tma_store_2d(tmap_out, stage, n0 + c0, c_row);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
}

// ---- 32-column bf16 output whose rows are contiguous in memory (ldc == 32: conv1's activation, conv2's data gradient) ----
// A warp's 32 rows x 64 bytes are ONE contiguous 2 KB block of the output.  Stored straight from the TMEM register layout
// (one row per thread) every instruction writes 16 bytes into each of 32 different sectors: four partial-sector writes per
// sector, and the L2 write path bounds the kernel (conv1 forward: 92 us for 127 MB at B = 16384).  Here the warp transposes
// through a 2 KB shared-memory tile (XOR-swizzled 16-byte chunks: conflict-free both ways) and writes 512 contiguous bytes
// per instruction.  rows_ok = number of valid rows of this warp (clipped at the end of the matrix); mask_row as above.
__host__ __device__ inline bool rows32_eligible(const Epilogue& ep) {
    return ep.out_bf16 && !ep.out_f32 && !ep.mask_f32 && !ep.drop && !ep.partial && ep.ldc == 32;
}
__device__ __forceinline__ void epilogue_rows32_bf16(const Epilogue& ep, uint32_t tmem_acc, int q, int lane, const __nv_bfloat16* mask_row,
                                                     long long row0, int rows_ok, unsigned char* stage) {
    uint32_t r[32];
    tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16), r);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (ep.bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + j));
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
    }
    if (ep.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (mask_row) {
        const uint4* mk = reinterpret_cast<const uint4*>(mask_row);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            const uint4 w = __ldg(mk + j / 8);
            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&ws[k]);
                if (!(__low2float(b2) > 0.f)) v[j + 2 * k] = 0.f;
                if (!(__high2float(b2) > 0.f)) v[j + 2 * k + 1] = 0.f;
            }
        }
    }
    __syncwarp();                                          // the previous tile's reads of the staging block are done
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * c], v[8 * c + 1]), p1 = __floats2bfloat162_rn(v[8 * c + 2], v[8 * c + 3]);
        const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * c + 4], v[8 * c + 5]), p3 = __floats2bfloat162_rn(v[8 * c + 6], v[8 * c + 7]);
        *reinterpret_cast<uint4*>(stage + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
            make_uint4(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1),
                       *reinterpret_cast<const uint32_t*>(&p2), *reinterpret_cast<const uint32_t*>(&p3));
    }
    __syncwarp();
    unsigned char* dst = reinterpret_cast<unsigned char*>(ep.out_bf16 + row0 * 32);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int lin = k * 32 + lane, row = lin >> 2, chunk = lin & 3;
        if (row < rows_ok)
            *reinterpret_cast<uint4*>(dst + lin * 16) = *reinterpret_cast<const uint4*>(stage + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
    }
}

// common prologue: barriers, tensor-map prefetch, TMEM allocation
struct Pipe {
    uint64_t *full_bar, *empty_bar, *tmem_full_bar;
    uint32_t* tmem_ptr;
};
template <int STAGES>
__device__ __forceinline__ uint32_t pipe_init(Pipe& pp, unsigned char* bars, const CUtensorMap* ta, const CUtensorMap* tb, uint32_t tmem_cols) {
    pp.full_bar = (uint64_t*)bars;
    pp.empty_bar = pp.full_bar + STAGES;
    pp.tmem_full_bar = pp.empty_bar + STAGES;
    pp.tmem_ptr = (uint32_t*)(pp.tmem_full_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&pp.full_bar[s], 1); mbar_init(&pp.empty_bar[s], 1); }
        mbar_init(pp.tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(ta) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(tb) : "memory");
    }
    if (warp == 1) {      // TMEM allocation by one full warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(pp.tmem_ptr)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *pp.tmem_ptr;
}
__device__ __forceinline__ void pipe_fini(uint32_t tmem_base, uint32_t tmem_cols) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

// =====================================================================================================================
// K-major kernel.  BK = 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B, the 32-channel activations of conv2).
// =====================================================================================================================
struct ConvArgs { int cblocks; int flip; };     // K-blocks per tap (= Cin / BK); mirrored taps

template <int BN, int STAGES, int BK, int MT = 1>
struct SmemLayout {
    static constexpr int A_BYTES = MT * BM * BK * 2, B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// MT = 128-row tiles of A per CTA (1 or 2): with MT = 2 one B tile feeds two accumulators, 256 x BN outputs per CTA — the
// operand bytes pulled from L2 per flop drop from 1/64 (128 x 128) to 1/128 (256 x 256), which is what bounds the fc1 GEMMs.
template <int BN, int STAGES, int BK, bool CONV, int MT = 1>
__global__ void __launch_bounds__(THREADS)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, long long M, int N, int K, int k_chunk, int tma_store, ConvArgs cv,
                    Epilogue ep) {
    extern __shared__ unsigned char smem_raw[];
    using L = SmemLayout<BN, STAGES, BK, MT>;
    static_assert(MT == 1 || (!CONV && MT * BN <= 512), "two row tiles: plain GEMM, accumulators within the 512 TMEM columns");
    constexpr uint64_t LAYOUT = BK == 64 ? 2 : 4;
    constexpr uint32_t SBO = 8 * BK * 2;
    unsigned char* tiles = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);     // swizzle atoms: 1024 B alignment
    constexpr uint32_t TMEM_COLS = MT * BN < 32 ? 32 : MT * BN;
    Pipe pp;
    const uint32_t tmem_base = pipe_init<STAGES>(pp, tiles + STAGES * L::STAGE_BYTES, &tmap_a, &tmap_b, TMEM_COLS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.y, n0 = blockIdx.x * BN;       // CONV: m_tile = sample
    const int k_begin = blockIdx.z * k_chunk;
    const int k_end = min(K, k_begin + k_chunk);
    const int num_kb = (k_end - k_begin + BK - 1) / BK;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t phase = 1;          // fresh barriers: the first pass over the ring does not wait
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&pp.empty_bar[s], phase);
                unsigned char* a_dst = tiles + s * L::STAGE_BYTES;
                unsigned char* b_dst = a_dst + L::A_BYTES;
                const int k = k_begin + kb * BK;
                if (CONV) {
                    mbar_expect_tx(&pp.full_bar[s], PIXELS * BK * 2 + L::B_BYTES);
                    const int kbg = k / BK, tap = kbg / cv.cblocks, cb = kbg - tap * cv.cblocks;
                    int di = tap / 3 - 1, dj = tap % 3 - 1;
                    if (cv.flip) { di = -di; dj = -dj; }
                    tma_load_4d(a_dst, &tmap_a, &pp.full_bar[s], cb * BK, dj, di, m_tile);
                } else {
                    mbar_expect_tx(&pp.full_bar[s], L::STAGE_BYTES);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) tma_load_2d(a_dst + mt * (BM * BK * 2), &tmap_a, &pp.full_bar[s], k, (m_tile * MT + mt) * BM);
                }
                tma_load_2d(b_dst, &tmap_b, &pp.full_bar[s], k, n0);
                if (s == STAGES - 1) phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            const uint64_t desc0 = make_smem_desc(smem_u32(tiles), SBO, 0, LAYOUT);       // advanced by (byte offset >> 4), never rebuilt
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&pp.full_bar[s], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t adesc = desc0 + (uint64_t)((uint32_t)(s * L::STAGE_BYTES) >> 4);
                const uint64_t bdesc = adesc + (uint64_t)(L::A_BYTES >> 4);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
                        umma_bf16(tmem_base + (uint32_t)(mt * BN), adesc + (uint64_t)((mt * BM * BK * 2 + k * UMMA_K * 2) >> 4),
                                  bdesc + (uint64_t)((k * UMMA_K * 2) >> 4), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                umma_commit(&pp.empty_bar[s]);          // frees the stage once these MMAs have read it
                if (s == STAGES - 1) phase ^= 1;
            }
            umma_commit(pp.tmem_full_bar);              // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
        const int q = warp & 3;
        mbar_wait(pp.tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int r = q * 32 + lane;
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
            const int mtile = m_tile * MT + mt;
            const uint32_t acc = tmem_base + (uint32_t)(mt * BN);
            const long long row = CONV ? (long long)m_tile * PIXELS + r : (long long)mtile * BM + r;
            const bool ok = (CONV ? r < PIXELS : true) && row < M && num_kb > 0;
            bool done = false;
            if constexpr (BN % 64 == 0) {
                if (tma_store) {      // all MMAs have completed: the pipeline stages are free, stage 0 becomes the staging area
                    const __nv_bfloat16* mrow = (ep.mask_bf16 && ok) ? ep.mask_bf16 + (size_t)row * ep.ldc + n0 : nullptr;
                    epilogue_tile_tma<BN>(ep, &tmap_out, acc, q, lane, mrow, n0, CONV ? q * 32 : mtile * BM + q * 32, CONV ? m_tile : -1,
                                          tiles + q * 4096);
                    done = true;
                }
            }
            if (!done) epilogue_tile<BN>(ep, acc, q, ok, row, n0, M, N, blockIdx.z);
        }
        if (tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    pipe_fini(tmem_base, TMEM_COLS);
}

// =====================================================================================================================
// CTA-pair kernel (cta_group::2): two CTAs of a cluster (consecutive 128-row tiles) compute a 256 x BN tile with ONE MMA
// stream issued by the leader.  Each CTA loads its own 128 rows of A and only HALF of the B tile (BN / 2 rows); the tensor
// cores of the pair exchange the halves, so every SM reads 4 KB + BN/2 * 32 B of operands per MMA instead of 4 KB + BN * 32 B
// and pulls half of the B bytes from L2.  The peer's TMA loads complete on the LEADER's full barrier (barrier address with
// the CTA-rank bit cleared), tcgen05.commit multicasts to the empty / accumulator barriers of both CTAs, each CTA drains its
// own 128 rows of D from its own TMEM.  Mechanics established by scripts/probes/cta_pair_probe.cu.
// =====================================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c_inner), "r"(c_outer) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {      // arrives on this barrier in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(THREADS)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_out, long long M, int N, int K, int k_chunk, int tma_store, Epilogue ep) {
    extern __shared__ unsigned char smem_raw[];
    constexpr int BK = 64;
    constexpr int A_BYTES = BM * BK * 2, B_BYTES = (BN / 2) * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    static_assert(BN % 32 == 0 && BN <= 256, "pair tile: N a multiple of 32 (16 per CTA half would do for the MMA, 32 for the loads)");
    unsigned char* tiles = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(tiles + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint32_t* tmem_ptr = (uint32_t*)(tmem_full + 1);
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                 // the peer's barriers exist before anything is signalled on them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    const int m_tile = blockIdx.x, n0 = blockIdx.y * BN;      // the pair = two consecutive row tiles = cluster (2, 1, 1) along x
    const int k_begin = blockIdx.z * k_chunk;
    const int k_end = min(K, k_begin + k_chunk);
    const int num_kb = (k_end - k_begin + BK - 1) / BK;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t phase = 1;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&empty_bar[s], phase);                                   // own copy: the commit arrives in both CTAs
                if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);      // the leader's barrier counts the bytes of both
                unsigned char* a_dst = tiles + s * STAGE_BYTES;
                const int k = k_begin + kb * BK;
                tma_load_2d_pair(a_dst, &tmap_a, &full_bar[s], k, m_tile * BM);
                tma_load_2d_pair(a_dst + A_BYTES, &tmap_b, &full_bar[s], k, n0 + (int)rank * (BN / 2));
                if (s == STAGES - 1) phase ^= 1;
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // instruction descriptor of the pair: M = 256, N = BN
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            const uint64_t desc0 = make_smem_desc(smem_u32(tiles), 1024u, 0, 2);
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&full_bar[s], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t adesc = desc0 + (uint64_t)((uint32_t)(s * STAGE_BYTES) >> 4);
                const uint64_t bdesc = adesc + (uint64_t)(A_BYTES >> 4);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                    umma_bf16_pair(tmem_base, adesc + (uint64_t)((k * UMMA_K * 2) >> 4), bdesc + (uint64_t)((k * UMMA_K * 2) >> 4), idesc,
                                   (kb > 0 || k > 0) ? 1u : 0u);
                umma_commit_pair(&empty_bar[s]);
                if (s == STAGES - 1) phase ^= 1;
            }
            umma_commit_pair(tmem_full);
        }
    } else {
        const int q = warp & 3;
        if (num_kb > 0) mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int r = q * 32 + lane;
        const long long row = (long long)m_tile * BM + r;
        const bool ok = row < M && num_kb > 0;
        bool done = false;
        if constexpr (BN % 64 == 0) {
            if (tma_store && num_kb > 0) {
                const __nv_bfloat16* mrow = (ep.mask_bf16 && ok) ? ep.mask_bf16 + (size_t)row * ep.ldc + n0 : nullptr;
                epilogue_tile_tma<BN>(ep, &tmap_out, tmem_base, q, lane, mrow, n0, m_tile * BM + q * 32, -1, tiles + q * 4096);
                if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                done = true;
            }
        }
        if (!done) epilogue_tile<BN>(ep, tmem_base, q, ok, row, n0, M, N, blockIdx.z);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                 // nobody leaves (or frees TMEM) while the pair's MMAs may still read its shared memory
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// =====================================================================================================================
// Persistent implicit-GEMM convolution: the whole weight matrix Wk[BN][9*Cin] stays in shared memory (<= 147 KB for the
// Q-network's layers), every CTA walks over samples (one 121-row tile each): the A ring streams the shifted boxes, the
// accumulator is double-buffered in TMEM so the epilogue of sample t overlaps the MMAs of sample t+1.
// =====================================================================================================================
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// EPI = number of epilogue warps: 4 (one per TMEM lane quadrant) or 8 (two sets of four; set 0 drains accumulator buffer 0 =
// the even samples of this CTA, set 1 buffer 1 = the odd ones, so two epilogues are in flight — the epilogue's latency per
// sample (TMEM load, pack, staging, TMA store: 1.4 - 2.3 us) and not the MMAs (0.3 - 1.2 us) bounds these kernels)
#ifdef MQ_CONV_TRACE
__device__ long long g_conv_trace[256 * 8];
#define TRACE_T0() const long long _t0 = clock64()
#define TRACE_ADD(slot) g_conv_trace[blockIdx.x * 8 + (slot)] += clock64() - _t0
#else
#define TRACE_T0()
#define TRACE_ADD(slot)
#endif
template <int BN, int BK, int STAGES, int EPI>
__global__ void __launch_bounds__(64 + 32 * EPI, 1)
conv_bf16_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                            const __grid_constant__ CUtensorMap tmap_out, long long batch, int nkb, int tma_store, int n_total, ConvArgs cv,
                            Epilogue ep) {
    extern __shared__ unsigned char smem_raw[];
    constexpr uint64_t LAYOUT = BK == 64 ? 2 : 4;
    constexpr uint32_t SBO = 8 * BK * 2;
    // One A stage = the three vertical taps of one column offset dj: a {BK, 11, 13, 1} box whose origin is (dj, -1) holds image
    // rows -1 .. 11 (143 tile rows, zero outside the image); tap di starts at tile row (di + 1) * 11, a row offset of the
    // shared-memory descriptor (the hardware swizzle works on absolute addresses: scripts/probes/tc_desc_probe.cu).  Three
    // boxes per sample and channel block instead of nine: TMA's per-row rate was the limit of the nine-box form.
    constexpr int A_ROWS = 13 * 11, ROW_BYTES = BK * 2, A_BYTES = 160 * ROW_BYTES, W_BYTES = BN * BK * 2;
    constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    const int n_base = blockIdx.y * BN;             // output-channel slice of this CTA (gridDim.y = Cout / BN)
    unsigned char* wtile = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* ring = wtile + (size_t)nkb * W_BYTES;
    unsigned char* stage = ring + STAGES * A_BYTES;            // EPI x 4 KB epilogue staging tiles (tma_store), EPI x 2 KB (32-column rows)
    const bool rows32 = BN == 32 && n_total == 32 && rows32_eligible(ep);
    uint64_t* full_bar = (uint64_t*)(stage + (tma_store ? EPI * 4096 : (rows32 ? EPI * 2048 : 0)));
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;           // [2]
    uint64_t* w_full = tmem_empty + 2;
    uint32_t* tmem_ptr = (uint32_t*)(w_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(w_full, (uint32_t)nkb * W_BYTES);
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(wtile + (size_t)kb * W_BYTES, &tmap_w, w_full, kb * BK, n_base);
            long long g = 0;
            for (long long sample = blockIdx.x; sample < batch; sample += gridDim.x) {
                for (int grp = 0; grp < 3 * cv.cblocks; ++grp, ++g) {          // (column offset, channel block)
                    const int s = (int)(g % STAGES);
                    { TRACE_T0(); mbar_wait(&empty_bar[s], (uint32_t)(((g / STAGES) & 1) ^ 1)); TRACE_ADD(0); }
                    mbar_expect_tx(&full_bar[s], A_ROWS * ROW_BYTES);
                    const int u = grp / cv.cblocks, cb = grp - u * cv.cblocks;
                    const int dj = cv.flip ? 1 - u : u - 1;
                    tma_load_4d(ring + s * A_BYTES, &tmap_a, &full_bar[s], cb * BK, dj, -1, (int)sample);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            const uint64_t a_desc0 = make_smem_desc(smem_u32(ring), SBO, 0, LAYOUT);
            const uint64_t b_desc0 = make_smem_desc(smem_u32(wtile), SBO, 0, LAYOUT);
            constexpr uint32_t tap_off[3] = {0u, (11u * ROW_BYTES) >> 4, (22u * ROW_BYTES) >> 4};
            mbar_wait(w_full, 0);
            long long g = 0;
            int it = 0;
            for (long long sample = blockIdx.x; sample < batch; sample += gridDim.x, ++it) {
                const int buf = it & 1;
                { TRACE_T0(); mbar_wait(&tmem_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1)); TRACE_ADD(1); }   // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                for (int grp = 0; grp < 3 * cv.cblocks; ++grp, ++g) {
                    const int s = (int)(g % STAGES);
                    { TRACE_T0(); mbar_wait(&full_bar[s], (uint32_t)((g / STAGES) & 1)); TRACE_ADD(2); }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int u = grp / cv.cblocks, cb = grp - u * cv.cblocks;
                    // The single issuing thread must not spend more than the ~64 cycles an MMA takes on preparing the next
                    // one: descriptors are advanced by adding (byte offset >> 4) to their address field, nothing is rebuilt.
                    const uint64_t a_desc = a_desc0 + (uint64_t)((uint32_t)(s * A_BYTES) >> 4);
#pragma unroll
                    for (int t = 0; t < 3; ++t) {                  // vertical tap di = t - 1 (mirrored when flip)
                        const int kb = (t * 3 + u) * cv.cblocks + cb;
                        const uint64_t a_t = a_desc + (uint64_t)(cv.flip ? tap_off[2 - t] : tap_off[t]);
                        const uint64_t b_t = b_desc0 + (uint64_t)((uint32_t)(kb * W_BYTES) >> 4);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_bf16(d_tmem, a_t + (uint64_t)((k * UMMA_K * 2) >> 4), b_t + (uint64_t)((k * UMMA_K * 2) >> 4), idesc,
                                      (grp > 0 || t > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        constexpr int HW = BN;                          // columns per epilogue warp
        const int n0 = n_base;                          // first output column of this warp
        const int eset = (warp - 2) >> 2;               // EPI == 8: warps 2-5 take the even samples of this CTA, 6-9 the odd ones
        int it = 0;
        for (long long sample = blockIdx.x; sample < batch; sample += gridDim.x, ++it) {
            const int buf = it & 1;
            if (EPI == 8 && buf != eset) continue;
            if (ep.mask_bf16 && r < PIXELS) {               // pull this row of the relu mask towards the SM while the MMAs run
                const __nv_bfloat16* mp = ep.mask_bf16 + ((size_t)sample * PIXELS + r) * n_total + n0;
#pragma unroll
                for (int b = 0; b < HW * 2; b += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"((const char*)mp + b));
            }
#ifdef MQ_CONV_TRACE
            const long long _e0 = clock64();
#endif
            mbar_wait(&tmem_full[buf], (uint32_t)((it >> 1) & 1));
#ifdef MQ_CONV_TRACE
            const long long _e1 = clock64();
            if (warp == 2 && lane == 0) g_conv_trace[blockIdx.x * 8 + 3] += _e1 - _e0;
#endif
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_base + (uint32_t)(buf * BN + (n0 - n_base));
            bool done = false;
            if constexpr (HW % 64 == 0) {
                if (tma_store) {
                    const __nv_bfloat16* mrow = (ep.mask_bf16 && r < PIXELS) ? ep.mask_bf16 + ((size_t)sample * PIXELS + r) * n_total + n0 : nullptr;
                    epilogue_tile_tma<HW>(ep, &tmap_out, acc, q, lane, mrow, n0, q * 32, (int)sample, stage + (warp - 2) * 4096);
                    done = true;
                }
            }
            if (!done && rows32) {
                const __nv_bfloat16* mrow = (ep.mask_bf16 && r < PIXELS) ? ep.mask_bf16 + ((size_t)sample * PIXELS + r) * 32 : nullptr;
                epilogue_rows32_bf16(ep, acc, q, lane, mrow, sample * PIXELS + q * 32, max(0, min(32, PIXELS - q * 32)), stage + (warp - 2) * 2048);
                done = true;
            }
            if (!done) epilogue_tile<HW>(ep, acc, q, r < PIXELS, sample * PIXELS + r, n0, batch * PIXELS, n_total, 0);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
#ifdef MQ_CONV_TRACE
            if (warp == 2 && lane == 0) { g_conv_trace[blockIdx.x * 8 + 4] += clock64() - _e1; g_conv_trace[blockIdx.x * 8 + 5] += 1; }
#endif
        }
        if (tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // stores complete before the CTA exits
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// =====================================================================================================================
// Persistent GEMM with a RESIDENT B operand (N == BN, K <= a few 64-wide blocks): for tall, skinny products such as conv1
// (495,616 x 32 x 64) a 128-row tile is 4 MMAs of work, and a CTA per tile spends its life in prologue and epilogue.  Here
// one CTA per SM keeps B in shared memory, streams the A tiles through a deep ring (the kernel is an HBM stream) and
// double-buffers the accumulator so that the epilogue of tile t overlaps the loads and MMAs of tile t+1.
// =====================================================================================================================
template <int BN, int STAGES>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_resident_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, long long M, int N,
                          int nkb, Epilogue ep) {
    extern __shared__ unsigned char smem_raw[];
    constexpr int BK = 64;
    constexpr int A_BYTES = BM * BK * 2, W_BYTES = BN * BK * 2;
    constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    unsigned char* wtile = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* ring = wtile + (((size_t)nkb * W_BYTES + 1023) & ~(size_t)1023);
    uint64_t* full_bar = (uint64_t*)(ring + STAGES * A_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;           // [2]
    uint64_t* w_full = tmem_empty + 2;
    uint32_t* tmem_ptr = (uint32_t*)(w_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;
    const long long m_tiles = (M + BM - 1) / BM;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(w_full, (uint32_t)nkb * W_BYTES);
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(wtile + (size_t)kb * W_BYTES, &tmap_w, w_full, kb * BK, 0);
            long long g = 0;
            for (long long tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = (int)(g % STAGES);
                    mbar_wait(&empty_bar[s], (uint32_t)(((g / STAGES) & 1) ^ 1));
                    mbar_expect_tx(&full_bar[s], A_BYTES);
                    tma_load_2d(ring + s * A_BYTES, &tmap_a, &full_bar[s], kb * BK, (int)(tile * BM));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            const uint64_t a_desc0 = make_smem_desc(smem_u32(ring), 1024u, 0, 2);
            const uint64_t b_desc0 = make_smem_desc(smem_u32(wtile), 1024u, 0, 2);
            mbar_wait(w_full, 0);
            long long g = 0;
            int it = 0;
            for (long long tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                mbar_wait(&tmem_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = (int)(g % STAGES);
                    mbar_wait(&full_bar[s], (uint32_t)((g / STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t a_d = a_desc0 + (uint64_t)((uint32_t)(s * A_BYTES) >> 4);
                    const uint64_t b_d = b_desc0 + (uint64_t)((uint32_t)(kb * W_BYTES) >> 4);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16(d_tmem, a_d + (uint64_t)((k * UMMA_K * 2) >> 4), b_d + (uint64_t)((k * UMMA_K * 2) >> 4), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else {
        const int q = warp & 3;
        int it = 0;
        for (long long tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(&tmem_full[buf], (uint32_t)((it >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long row = tile * BM + q * 32 + lane;
            epilogue_tile<BN>(ep, tmem_base + (uint32_t)(buf * BN), q, row < M, row, 0, M, N, 0);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// OBS variant (conv1 without an im2col buffer): the A tile is not loaded but BUILT in shared memory by eight extra warps, one
// thread per tile row, straight from the fp32 observation [B][11][11][6]: row = (sample, pixel), column k = tap * 6 + c for
// k < 54, 1.0 at k = 54 (the ones column of qnet_bf16.cuh), zero above; 16-byte chunk c of row r goes to chunk c ^ (r & 7)
// (the SWIZZLE_128B pattern the K-major descriptors expect).  Generic-proxy writes, then fence.proxy.async before the arrive.
// observation staging of the conv1 builders: the whole samples a 128-row tile touches (at most three: 128 > 121), fp32
constexpr int CONV1_STAGE_FLOATS = 3 * PIXELS * 6 + 2;       // 8720 bytes, a multiple of 16
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
constexpr int CONV1_THREADS = THREADS + 256 + 128;           // producer, MMA, 4 epilogue warps, 8 builder warps, 4 more epilogue warps
template <int BN, int STAGES>
__global__ void __launch_bounds__(CONV1_THREADS, 1)
conv1_obs_resident_kernel(const float* __restrict__ obs, const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a1,
                          int store_a1, long long M, int N, Epilogue ep) {
    constexpr int nkb = 1;
    extern __shared__ unsigned char smem_raw[];
    constexpr int BK = 64;
    constexpr int A_BYTES = BM * BK * 2, W_BYTES = BN * BK * 2;
    constexpr int NACC = 4;                       // accumulator ring in TMEM (4 x BN columns)
    constexpr uint32_t TMEM_COLS = NACC * BN < 32 ? 32 : NACC * BN;
    unsigned char* wtile = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* ring = wtile + (((size_t)nkb * W_BYTES + 1023) & ~(size_t)1023);
    float* obs_stage = (float*)(ring + STAGES * A_BYTES);              // [2 builder groups][2 buffers][CONV1_STAGE_FLOATS]
    unsigned char* epi_stage = (unsigned char*)(obs_stage + 4 * CONV1_STAGE_FLOATS);     // [8 epilogue warps][2 KB]
    uint64_t* full_bar = (uint64_t*)(epi_stage + 8 * 2048);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;       // [NACC]
    uint64_t* tmem_empty = tmem_full + NACC;        // [NACC]
    uint64_t* w_full = tmem_empty + NACC;
    uint32_t* tmem_ptr = (uint32_t*)(w_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 4); mbar_init(&empty_bar[s], 1); }      // four builder warps fill a stage
        for (int b = 0; b < NACC; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;
    const long long m_tiles = (M + BM - 1) / BM;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(w_full, (uint32_t)nkb * W_BYTES);
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(wtile + (size_t)kb * W_BYTES, &tmap_w, w_full, kb * BK, 0);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            const uint64_t a_desc0 = make_smem_desc(smem_u32(ring), 1024u, 0, 2);
            const uint64_t b_desc0 = make_smem_desc(smem_u32(wtile), 1024u, 0, 2);
            mbar_wait(w_full, 0);
            long long g = 0;
            int it = 0;
            for (long long tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
                const int buf = it % NACC;
                mbar_wait(&tmem_empty[buf], (uint32_t)(((it / NACC) & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = (int)(g % STAGES);
                    mbar_wait(&full_bar[s], (uint32_t)((g / STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t a_d = a_desc0 + (uint64_t)((uint32_t)(s * A_BYTES) >> 4);
                    const uint64_t b_d = b_desc0 + (uint64_t)((uint32_t)(kb * W_BYTES) >> 4);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16(d_tmem, a_d + (uint64_t)((k * UMMA_K * 2) >> 4), b_d + (uint64_t)((k * UMMA_K * 2) >> 4), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tmem_full[buf]);
            }
        }
    } else if (warp >= 6 && warp < 14) {
        // ===== builders: warps 6..13, thread j builds row j of the tile =====
        // two groups of four warps take alternate tiles.  The observation of the samples a tile touches is first brought to
        // shared memory with coalesced asynchronous copies (the tile AFTER the one being built is in flight meanwhile); the 27
        // values of a row's 3x3x6 neighbourhood are then shared-memory reads.  Reading them straight from global memory (27
        // scattered 8-byte loads per thread) left the kernel waiting on L2 round trips: 30 % of the HBM rate.
        const int j = ((warp - 6) & 3) * 32 + lane, bgrp = (warp - 6) >> 2;
        constexpr int OBS_F = PIXELS * 6;
        float* const stg = obs_stage + bgrp * 2 * CONV1_STAGE_FLOATS;
        const long long batch = M / PIXELS;
        auto prefetch = [&](long long tile, int buf) {
            const long long b_lo = tile * BM / PIXELS;
            long long b_hi = (tile * BM + BM - 1) / PIXELS;
            if (b_hi > batch - 1) b_hi = batch - 1;
            const int n8 = (int)(b_hi - b_lo + 1) * (OBS_F / 2);              // 8-byte pieces (a sample = 2904 bytes, 8-byte aligned)
            const float* src = obs + b_lo * OBS_F;
            float* dst = stg + buf * CONV1_STAGE_FLOATS;
            for (int c = j; c < n8; c += 128) cp_async8(dst + 2 * c, src + 2 * c);
        };
        {
            const long long first = (long long)blockIdx.x + (long long)bgrp * gridDim.x;
            if (first < m_tiles) prefetch(first, 0);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        long long g = bgrp;
        int kown = 0;
        for (long long tile = (long long)blockIdx.x + g * gridDim.x; tile < m_tiles; tile += 2LL * gridDim.x, g += 2, ++kown) {
            const int buf = kown & 1;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            // every copy of this tile has landed, and every thread of the group is done reading the other buffer (previous tile)
            asm volatile("bar.sync %0, 128;" ::"r"(3 + bgrp) : "memory");
            if (tile + 2LL * gridDim.x < m_tiles) prefetch(tile + 2LL * gridDim.x, buf ^ 1);
            asm volatile("cp.async.commit_group;" ::: "memory");
            const int s = (int)(g % STAGES);
            mbar_wait(&empty_bar[s], (uint32_t)(((g / STAGES) & 1) ^ 1));
            const long long r = tile * BM + j;
            float v[64];
#pragma unroll
            for (int k = 0; k < 64; ++k) v[k] = 0.f;
            if (r < M) {
                const long long b = r / PIXELS;
                const int q = (int)(r - b * PIXELS), i = q / 11, jx = q - i * 11;
                const float* src = stg + buf * CONV1_STAGE_FLOATS + (int)(b - tile * BM / PIXELS) * OBS_F;
                // all 27 loads are issued unconditionally (clamped coordinates); taps outside the image are zeroed afterwards
                float2 x[9][3];
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int ii = min(max(i + t / 3 - 1, 0), 10), jj = min(max(jx + t % 3 - 1, 0), 10);
                    const float2* p2 = reinterpret_cast<const float2*>(src + (ii * 11 + jj) * 6);
                    x[t][0] = p2[0]; x[t][1] = p2[1]; x[t][2] = p2[2];
                }
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int ii = i + t / 3 - 1, jj = jx + t % 3 - 1;
                    const bool in = (unsigned)ii < 11u && (unsigned)jj < 11u;
                    v[t * 6] = in ? x[t][0].x : 0.f; v[t * 6 + 1] = in ? x[t][0].y : 0.f; v[t * 6 + 2] = in ? x[t][1].x : 0.f;
                    v[t * 6 + 3] = in ? x[t][1].y : 0.f; v[t * 6 + 4] = in ? x[t][2].x : 0.f; v[t * 6 + 5] = in ? x[t][2].y : 0.f;
                }
                v[54] = 1.f;
            }
            unsigned char* rowp = ring + s * A_BYTES + j * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * c], v[8 * c + 1]), p1 = __floats2bfloat162_rn(v[8 * c + 2], v[8 * c + 3]);
                const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * c + 4], v[8 * c + 5]), p3 = __floats2bfloat162_rn(v[8 * c + 6], v[8 * c + 7]);
                *reinterpret_cast<uint4*>(rowp + ((c ^ (j & 7)) << 4)) =
                    make_uint4(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1),
                               *reinterpret_cast<const uint32_t*>(&p2), *reinterpret_cast<const uint32_t*>(&p3));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
            if (store_a1) {
                // The tile just built IS a 128 x 64 block of the im2col matrix A1 in the layout TMA writes: when the weight
                // gradient will need A1, the group's first lane sends the stage out as one bulk tensor store (rows >= M are
                // clipped).  The stage is only overwritten STAGES tiles later, and not before this store has read it.
                asm volatile("bar.sync %0, 128;" ::"r"(1 + bgrp) : "memory");                 // all four warps of the group have written
                if (j == 0) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");            // (previous store of this group)
                    tma_store_2d(&tmap_a1, ring + s * A_BYTES, 0, (int)(tile * BM));
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (store_a1 && j == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
        // ===== epilogue: two sets of four warps (2..5 and 14..17), set e drains accumulator buffer e = every other tile.  One
        // set (TMEM load, bias, ReLU, pack, stores: ~500 cycles per tile) was as slow as the builders. =====
        const int q = warp & 3, eset = warp >= 14 ? 1 : 0;
        int it = 0;
        for (long long tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
            const int buf = it % NACC;
            if ((it & 1) != eset) continue;
            mbar_wait(&tmem_full[buf], (uint32_t)((it / NACC) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long row = tile * BM + q * 32 + lane;
            if (BN == 32 && rows32_eligible(ep)) {
                const long long row0 = tile * BM + q * 32;
                const int rows_ok = (int)max(0LL, min(32LL, M - row0));
                const __nv_bfloat16* mrow = (ep.mask_bf16 && row < M) ? ep.mask_bf16 + row * 32 : nullptr;
                epilogue_rows32_bf16(ep, tmem_base + (uint32_t)(buf * BN), q, lane, mrow, row0, rows_ok, epi_stage + ((warp >= 14 ? warp - 10 : warp - 2) << 11));
            } else {
                epilogue_tile<BN>(ep, tmem_base + (uint32_t)(buf * BN), q, row < M, row, 0, M, N, 0);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

// =====================================================================================================================
// MN-major ("TN") kernel: C[M][N] = sum_k At[k][m] * Bt[k][n].  A stage holds BKR k-rows: 2 A slabs and BN/64 B slabs of
// [BKR][64] bf16 (128-byte rows, SWIZZLE_128B).  CONV: one stage = one sample (121 of the 128 rows are loaded, the rest
// stay zero); the A slab of output rows m0 + 64 j is tap (m / Cin), channels (m % Cin) .. +64 of the shifted activation.
// =====================================================================================================================
template <int BN, int STAGES, int BKR, int AW>
struct SmemLayoutTN {
    static constexpr int A_SLAB = BKR * AW * 2, A_SLABS = BM / AW;        // AW = 64: 128 B rows (SWIZZLE_128B); 32: 64 B rows (SWIZZLE_64B)
    static constexpr int B_SLAB = BKR * 128, B_SLABS = BN / 64;
    static constexpr int A_BYTES = A_SLABS * A_SLAB, B_BYTES = B_SLABS * B_SLAB;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN, int STAGES, int BKR, bool CONV, int AW>
__global__ void __launch_bounds__(THREADS)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M, int N, int K,
                    int k_chunk, int cin, Epilogue ep) {
    extern __shared__ unsigned char smem_raw[];
    using L = SmemLayoutTN<BN, STAGES, BKR, AW>;
    static_assert(BN % 64 == 0 && BKR % UMMA_K == 0 && (AW == 64 || AW == 32), "tile shape");
    constexpr uint32_t A_ROW = AW * 2;                  // bytes per k-row of an A slab
    constexpr uint64_t A_LAYOUT = AW == 64 ? 2 : 4;
    unsigned char* tiles = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t TMEM_COLS = BN;
    if (CONV) {     // rows 121..BKR-1 of every slab are never written by TMA: they must read as zero
        constexpr int PAD = BKR - PIXELS;
        for (int i = threadIdx.x; i < STAGES * (L::A_SLABS * PAD * (int)A_ROW + L::B_SLABS * PAD * 128) / 16; i += THREADS) {
            const int per_stage = (L::A_SLABS * PAD * (int)A_ROW + L::B_SLABS * PAD * 128) / 16;
            const int st = i / per_stage;
            int o = i - st * per_stage;
            unsigned char* base = tiles + (size_t)st * L::STAGE_BYTES;
            unsigned char* dst;
            if (o < L::A_SLABS * PAD * (int)A_ROW / 16) {
                const int sl = o / (PAD * (int)A_ROW / 16); o -= sl * (PAD * (int)A_ROW / 16);
                dst = base + sl * L::A_SLAB + PIXELS * A_ROW + o * 16;
            } else {
                o -= L::A_SLABS * PAD * (int)A_ROW / 16;
                const int sl = o / (PAD * 128 / 16); o -= sl * (PAD * 128 / 16);
                dst = base + L::A_BYTES + sl * L::B_SLAB + PIXELS * 128 + o * 16;
            }
            *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // Bias gradient for free: when the last row tile has an A slab that starts exactly at row M (9 Cin is not a multiple of
    // 128), that slab is never loaded; filled once with ones in its first column (all 121 pixel rows), its first D row
    // becomes sum_k dY[k][n], the column sums of the B operand.
    const int slabs_here = min(L::A_SLABS, (M - m0 + AW - 1) / AW);
    const bool ones_row = CONV && ep.colsum_partial != nullptr && slabs_here < L::A_SLABS && m0 + slabs_here * AW == M;
    if (ones_row) {
        constexpr int CHUNKS_PER_ROW = (int)A_ROW / 16;
        for (int i = threadIdx.x; i < STAGES * BKR * CHUNKS_PER_ROW; i += THREADS) {
            const int st = i / (BKR * CHUNKS_PER_ROW), o = i - st * (BKR * CHUNKS_PER_ROW);
            const int k = o / CHUNKS_PER_ROW, pc = o - k * CHUNKS_PER_ROW;
            const int logical = pc ^ (AW == 64 ? (k & 7) : ((k >> 1) & 3));          // TMA's swizzle: chunk index xor row bits
            unsigned char* dst = tiles + (size_t)st * L::STAGE_BYTES + slabs_here * L::A_SLAB + k * A_ROW + pc * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4((logical == 0 && k < PIXELS) ? 0x00003F80u : 0u, 0, 0, 0);     // bf16 1.0 in element 0
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    Pipe pp;
    const uint32_t tmem_base = pipe_init<STAGES>(pp, tiles + STAGES * L::STAGE_BYTES, &tmap_a, &tmap_b, TMEM_COLS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // K is counted in rows (2-D) or samples (CONV)
    const int k_begin = blockIdx.z * k_chunk;
    const int k_end = min(K, k_begin + k_chunk);
    const int num_kb = CONV ? (k_end - k_begin) : (k_end - k_begin + BKR - 1) / BKR;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t phase = 1;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&pp.empty_bar[s], phase);
                unsigned char* a_dst = tiles + s * L::STAGE_BYTES;
                unsigned char* b_dst = a_dst + L::A_BYTES;
                if (CONV) {
                    const int sample = k_begin + kb;
                    const int a_slabs = min(L::A_SLABS, (M - m0 + AW - 1) / AW);
                    mbar_expect_tx(&pp.full_bar[s], a_slabs * PIXELS * A_ROW + L::B_SLABS * PIXELS * 128);
                    for (int j = 0; j < a_slabs; ++j) {
                        const int m = m0 + AW * j, tap = m / cin, c0 = m - tap * cin;
                        tma_load_4d(a_dst + j * L::A_SLAB, &tmap_a, &pp.full_bar[s], c0, tap % 3 - 1, tap / 3 - 1, sample);
                    }
#pragma unroll
                    for (int j = 0; j < L::B_SLABS; ++j) tma_load_2d(b_dst + j * L::B_SLAB, &tmap_b, &pp.full_bar[s], n0 + 64 * j, sample * PIXELS);
                } else {
                    const int k = k_begin + kb * BKR;
                    const int a_slabs = min(L::A_SLABS, (M - m0 + AW - 1) / AW);       // slabs past M are not loaded (their D rows are not stored)
                    mbar_expect_tx(&pp.full_bar[s], a_slabs * L::A_SLAB + L::B_BYTES);
                    for (int j = 0; j < a_slabs; ++j) tma_load_2d(a_dst + j * L::A_SLAB, &tmap_a, &pp.full_bar[s], m0 + AW * j, k);
#pragma unroll
                    for (int j = 0; j < L::B_SLABS; ++j) tma_load_2d(b_dst + j * L::B_SLAB, &tmap_b, &pp.full_bar[s], n0 + 64 * j, k);
                }
                if (s == STAGES - 1) phase ^= 1;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN, true);
            const uint64_t a_desc0 = make_smem_desc(smem_u32(tiles), 8 * A_ROW, (uint32_t)L::A_SLAB, A_LAYOUT);      // advanced, never rebuilt
            const uint64_t b_desc0 = make_smem_desc(smem_u32(tiles) + L::A_BYTES, 1024u, (uint32_t)L::B_SLAB, 2);
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&pp.full_bar[s], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t st_off = (uint64_t)((uint32_t)(s * L::STAGE_BYTES) >> 4);
#pragma unroll
                for (int k = 0; k < BKR / UMMA_K; ++k)
                    umma_bf16(tmem_base, a_desc0 + st_off + (uint64_t)((k * UMMA_K * A_ROW) >> 4), b_desc0 + st_off + (uint64_t)((k * UMMA_K * 128) >> 4),
                              idesc, (kb > 0 || k > 0) ? 1u : 0u);
                umma_commit(&pp.empty_bar[s]);
                if (s == STAGES - 1) phase ^= 1;
            }
            umma_commit(pp.tmem_full_bar);
        }
    } else {
        const int q = warp & 3;
        mbar_wait(pp.tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long row = m0 + q * 32 + lane;
        epilogue_tile<BN>(ep, tmem_base, q, row < M && num_kb > 0, row, n0, M, N, blockIdx.z);
        if (ones_row && q == (slabs_here * AW) / 32) {          // the warp that owns the ones row stores it (warp-uniform branch)
            const bool mine = lane == (slabs_here * AW) % 32;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
                if (mine) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + c0 + j < N) ep.colsum_partial[(size_t)blockIdx.z * N + n0 + c0 + j] = num_kb > 0 ? __uint_as_float(r[j]) : 0.f;
                }
            }
        }
    }
    pipe_fini(tmem_base, TMEM_COLS);
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

// 2D bf16 row-major matrix [rows][cols] (cols contiguous, leading dimension ld elements), box = [box_rows][box_cols];
// box_cols = 64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B
inline bool make_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols = 64) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// NHWC activation [B][11][11][C] bf16 as a 4-D tensor (C, x, y, b); box = {box_c channels, 11, 11, 1}: one sample's window,
// shifted by the tap through the box origin, zero-filled outside the 11 x 11 image
inline bool make_tmap_act(CUtensorMap* out, const void* base, uint64_t batch, uint64_t C, uint32_t box_c, uint32_t box_y = 11) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[4] = {C, 11, 11, batch};
    cuuint64_t strides[3] = {C * 2, 11 * C * 2, 121 * C * 2};
    cuuint32_t box[4] = {box_c, 11, box_y, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// bf16 output maps of the TMA-store epilogue, box = 32 rows x 64 columns (SWIZZLE_128B)
inline bool make_tmap_out2d(CUtensorMap* out, void* base, uint64_t rows, uint64_t cols, uint64_t ld) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, 32};
    cuuint32_t estr[2] = {1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
inline bool make_tmap_out_conv(CUtensorMap* out, void* base, uint64_t batch, uint64_t Cout) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {Cout, (cuuint64_t)PIXELS, batch};
    cuuint64_t strides[2] = {Cout * 2, (cuuint64_t)PIXELS * Cout * 2};
    cuuint32_t box[3] = {64, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// the TMA-store epilogue applies to a bf16-only output with at most bias / relu / a bf16 relu-mask
inline bool tma_store_eligible(const Epilogue& ep, int N, int BN) {
    return BN % 64 == 0 && N % BN == 0 && ep.out_bf16 && !ep.out_f32 && !ep.mask_f32 && !ep.drop && ep.ldc % 8 == 0;
}

// The dynamic shared-memory limit is a per-device function attribute: remember per device what has been raised already
// (one process normally drives one GPU, but nothing here may depend on that).
struct SmemMemo { int bytes[64]; };
template <typename KernelT>
inline cudaError_t ensure_smem(SmemMemo& memo, KernelT kernel, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const bool tracked = dev >= 0 && dev < 64;
    if (tracked && bytes <= memo.bytes[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && tracked) memo.bytes[dev] = bytes;
    return e;
}

// C[M][N] = A[M][K] B[N][K]^T
template <int BN, int STAGES, int MT = 1, int BK = 64>
inline cudaError_t launch(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K, Epilogue ep,
                          int* splits_inout, cudaStream_t stream) {
    int splits = splits_inout ? *splits_inout : 1;
    CUtensorMap ta, tb;
    if (!make_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK) || !make_tmap(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN, BK))
        return cudaErrorInvalidValue;
    using L = SmemLayout<BN, STAGES, BK, MT>;
    static SmemMemo memo{};
    if (cudaError_t e = ensure_smem(memo, gemm_bf16_tc_kernel<BN, STAGES, BK, false, MT>, L::TOTAL); e != cudaSuccess) return e;
    const int k_tiles = (K + BK - 1) / BK;
    if (splits < 1) splits = 1;
    if (splits > k_tiles) splits = k_tiles;
    const int chunk_tiles = (k_tiles + splits - 1) / splits;
    splits = (k_tiles + chunk_tiles - 1) / chunk_tiles;
    if (splits_inout) *splits_inout = splits;
    if (splits > 1 && !ep.partial) return cudaErrorInvalidValue;
    if (splits == 1) ep.partial = nullptr;
    CUtensorMap to = ta;
    const int tma_store = (splits == 1 && tma_store_eligible(ep, N, BN) && make_tmap_out2d(&to, ep.out_bf16, (uint64_t)M, (uint64_t)N, (uint64_t)ep.ldc)) ? 1 : 0;
    dim3 grid((N + BN - 1) / BN, (M + MT * BM - 1) / (MT * BM), splits);
    gemm_bf16_tc_kernel<BN, STAGES, BK, false, MT><<<grid, THREADS, L::TOTAL, stream>>>(ta, tb, to, M, N, K, chunk_tiles * BK, tma_store,
                                                                                   ConvArgs{1, 0}, ep);
    return cudaGetLastError();
}

// C[M][N] = A[M][K] B[N][K]^T for N == BN and K a small multiple of 64: persistent kernel with B resident in shared memory
template <int BN, int STAGES>
inline cudaError_t launch_resident(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, long long M, int N, int K, Epilogue ep,
                                   int n_sms, cudaStream_t stream) {
    constexpr int BK = 64;
    if (N != BN || K % BK != 0 || M <= 0) return cudaErrorInvalidValue;
    const int nkb = K / BK;
    CUtensorMap ta, tw;
    if (!make_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM) || !make_tmap(&tw, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN))
        return cudaErrorInvalidValue;
    const int smem = ((nkb * BN * BK * 2 + 1023) & ~1023) + STAGES * BM * BK * 2 + 1024 + 256;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    static SmemMemo memo{};
    if (cudaError_t e = ensure_smem(memo, gemm_bf16_resident_kernel<BN, STAGES>, smem); e != cudaSuccess) return e;
    ep.partial = nullptr;
    const long long m_tiles = (M + BM - 1) / BM;
    const unsigned grid = (unsigned)(m_tiles < n_sms ? m_tiles : n_sms);
    gemm_bf16_resident_kernel<BN, STAGES><<<grid, THREADS, smem, stream>>>(ta, tw, M, N, nkb, ep);
    return cudaGetLastError();
}

// conv1 forward straight from the fp32 observation (no im2col buffer): Y[B*121][32] = relu(im2col(obs) W1c^T + b)
template <int STAGES>
// a1_out (optional): also write the im2col rows A1 [batch*121][64] bf16 (the conv1 weight gradient reads them)
inline cudaError_t launch_conv1_obs(const float* obs, const __nv_bfloat16* W1c, long long batch, Epilogue ep, int n_sms, cudaStream_t stream,
                                    __nv_bfloat16* a1_out = nullptr) {
    constexpr int BN = 32, BK = 64;
    if (batch <= 0) return cudaErrorInvalidValue;
    CUtensorMap tw, ta1;
    if (!make_tmap(&tw, W1c, (uint64_t)BN, (uint64_t)BK, (uint64_t)BK, BN)) return cudaErrorInvalidValue;
    ta1 = tw;
    if (a1_out && !make_tmap(&ta1, a1_out, (uint64_t)batch * PIXELS, (uint64_t)BK, (uint64_t)BK, BM)) return cudaErrorInvalidValue;
    const int smem = ((BN * BK * 2 + 1023) & ~1023) + STAGES * BM * BK * 2 + 4 * CONV1_STAGE_FLOATS * 4 + 8 * 2048 + 1024 + 256;
    static SmemMemo memo{};
    if (cudaError_t e = ensure_smem(memo, conv1_obs_resident_kernel<BN, STAGES>, smem); e != cudaSuccess) return e;
    ep.partial = nullptr;
    const long long M = batch * PIXELS, m_tiles = (M + BM - 1) / BM;
    const unsigned grid = (unsigned)(m_tiles < n_sms ? m_tiles : n_sms);
    conv1_obs_resident_kernel<BN, STAGES><<<grid, CONV1_THREADS, smem, stream>>>(obs, tw, ta1, a1_out ? 1 : 0, M, BN, ep);
    return cudaGetLastError();
}

// CTA-pair form of launch(): 256 x BN outputs per pair of CTAs (cluster 2 x 1 x 1; grid.x walks over the row tiles)
template <int BN, int STAGES>
inline cudaError_t launch_pair(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K, Epilogue ep,
                               int* splits_inout, cudaStream_t stream) {
    constexpr int BK = 64;
    int splits = splits_inout ? *splits_inout : 1;
    CUtensorMap ta, tb;
    if (!make_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM) || !make_tmap(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN / 2))
        return cudaErrorInvalidValue;
    constexpr int SMEM = STAGES * (BM * BK * 2 + (BN / 2) * BK * 2) + 1024 + 256;
    static SmemMemo memo{};
    if (cudaError_t e = ensure_smem(memo, gemm_bf16_pair_kernel<BN, STAGES>, SMEM); e != cudaSuccess) return e;
    const int k_tiles = (K + BK - 1) / BK;
    if (splits < 1) splits = 1;
    if (splits > k_tiles) splits = k_tiles;
    const int chunk_tiles = (k_tiles + splits - 1) / splits;
    splits = (k_tiles + chunk_tiles - 1) / chunk_tiles;
    if (splits_inout) *splits_inout = splits;
    if (splits > 1 && !ep.partial) return cudaErrorInvalidValue;
    if (splits == 1) ep.partial = nullptr;
    CUtensorMap to = ta;
    const int tma_store = (splits == 1 && tma_store_eligible(ep, N, BN) && make_tmap_out2d(&to, ep.out_bf16, (uint64_t)M, (uint64_t)N, (uint64_t)ep.ldc)) ? 1 : 0;
    const int m_tiles = (M + BM - 1) / BM;
    dim3 grid((m_tiles + 1) & ~1, (N + BN - 1) / BN, splits);       // an even number of row tiles: the odd one out pairs with an empty tile
    // the pair is a cluster of two CTAs along the row tiles (launch attribute, so that the pairing is visible at the call site)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;      // cta_group::2 pairs are adjacent in x
    cfg.attrs = at; cfg.numAttrs = 1;
    const int k_chunk = chunk_tiles * BK;
    const long long Mll = M;
    return cudaLaunchKernelEx(&cfg, gemm_bf16_pair_kernel<BN, STAGES>, ta, tb, to, Mll, N, K, k_chunk, tma_store, ep);
}

// 3x3 / pad 1 convolution over 11 x 11 windows as an implicit GEMM: Y[B*121][Cout] = im2col(X [B][11][11][Cin]) * Wk[Cout][9*Cin]^T
// (taps in (kh, kw, c) order; flip = mirrored taps = the data-gradient convolution).  BK = 64 needs Cin % 64 == 0, BK = 32 Cin % 32 == 0.
template <int BN, int STAGES, int BK>
inline cudaError_t launch_conv(const __nv_bfloat16* X, const __nv_bfloat16* Wk, long long batch, int Cin, int Cout, int flip, Epilogue ep,
                               cudaStream_t stream) {
    if (Cin % BK != 0 || batch <= 0) return cudaErrorInvalidValue;
    CUtensorMap ta, tb;
    if (!make_tmap_act(&ta, X, (uint64_t)batch, (uint64_t)Cin, BK) ||
        !make_tmap(&tb, Wk, (uint64_t)Cout, (uint64_t)9 * Cin, (uint64_t)9 * Cin, BN, BK))
        return cudaErrorInvalidValue;
    using L = SmemLayout<BN, STAGES, BK>;
    static SmemMemo memo{};
    if (cudaError_t e = ensure_smem(memo, gemm_bf16_tc_kernel<BN, STAGES, BK, true>, L::TOTAL); e != cudaSuccess) return e;
    ep.partial = nullptr;
    CUtensorMap to = ta;
    const int tma_store = (tma_store_eligible(ep, Cout, BN) && ep.ldc == Cout && make_tmap_out_conv(&to, ep.out_bf16, (uint64_t)batch, (uint64_t)Cout)) ? 1 : 0;
    dim3 grid((Cout + BN - 1) / BN, (unsigned)batch, 1);
    gemm_bf16_tc_kernel<BN, STAGES, BK, true><<<grid, THREADS, L::TOTAL, stream>>>(ta, tb, to, batch * PIXELS, Cout, 9 * Cin, 9 * Cin, tma_store,
                                                                                 ConvArgs{Cin / BK, flip}, ep);
    return cudaGetLastError();
}

// Persistent form of launch_conv (Cout == BN): weights resident in shared memory, one CTA per SM walking over the samples.
template <int BN, int BK, int STAGES, int EPI = 4>
inline cudaError_t launch_conv_persistent(const __nv_bfloat16* X, const __nv_bfloat16* Wk, long long batch, int Cin, int Cout, int flip,
                                          Epilogue ep, int n_sms, cudaStream_t stream) {
    if (Cin % BK != 0 || batch <= 0 || Cout % BN != 0) return cudaErrorInvalidValue;
    CUtensorMap ta, tw;
    if (!make_tmap_act(&ta, X, (uint64_t)batch, (uint64_t)Cin, BK, 13) ||
        !make_tmap(&tw, Wk, (uint64_t)Cout, (uint64_t)9 * Cin, (uint64_t)9 * Cin, BN, BK))
        return cudaErrorInvalidValue;
    const int nkb = 9 * Cin / BK;
    // bf16 output without per-element fp32 masks: epilogue through shared memory + TMA store
    static_assert(EPI == 4 || EPI == 8, "4 or 8 epilogue warps");
    constexpr int HWE = BN;
    static_assert(HWE >= 32, "an epilogue warp handles at least 32 columns");
    CUtensorMap to = ta;
    const int tma_store = (HWE % 64 == 0 && tma_store_eligible(ep, Cout, BN) && ep.ldc == Cout && make_tmap_out_conv(&to, ep.out_bf16, (uint64_t)batch, (uint64_t)Cout)) ? 1 : 0;
    const int smem = nkb * BN * BK * 2 + STAGES * 160 * BK * 2 + (tma_store ? EPI * 4096 : (BN == 32 ? EPI * 2048 : 0)) + 1024 + 256;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    static SmemMemo memo{};
    if (cudaError_t e = ensure_smem(memo, conv_bf16_persistent_kernel<BN, BK, STAGES, EPI>, smem); e != cudaSuccess) return e;
    ep.partial = nullptr;
    const int slices = Cout / BN;                       // CTAs of different output-channel slices walk over the same samples
    const int per_slice = n_sms / slices < 1 ? 1 : n_sms / slices;
    dim3 grid((unsigned)(batch < per_slice ? batch : per_slice), (unsigned)slices, 1);
    conv_bf16_persistent_kernel<BN, BK, STAGES, EPI><<<grid, 64 + 32 * EPI, smem, stream>>>(ta, tw, to, batch, nkb, tma_store, Cout,
                                                                                          ConvArgs{Cin / BK, flip}, ep);
    return cudaGetLastError();
}

// C[M][N] = At[K][M]^T Bt[K][N]   (M % 8 == 0 and N % 8 == 0: TMA row pitch)
template <int BN, int STAGES, int AW = 64>
inline cudaError_t launch_tn(const __nv_bfloat16* At, int lda, const __nv_bfloat16* Bt, int ldb, int M, int N, int K, Epilogue ep,
                             int* splits_inout, cudaStream_t stream) {
    constexpr int BKR = 64;
    int splits = splits_inout ? *splits_inout : 1;
    CUtensorMap ta, tb;
    if (!make_tmap(&ta, At, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BKR, AW) || !make_tmap(&tb, Bt, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BKR))
        return cudaErrorInvalidValue;
    using L = SmemLayoutTN<BN, STAGES, BKR, AW>;
    static SmemMemo memo{};
    if (cudaError_t e = ensure_smem(memo, gemm_bf16_tn_kernel<BN, STAGES, BKR, false, AW>, L::TOTAL); e != cudaSuccess) return e;
    const int k_tiles = (K + BKR - 1) / BKR;
    if (splits < 1) splits = 1;
    if (splits > k_tiles) splits = k_tiles;
    const int chunk_tiles = (k_tiles + splits - 1) / splits;
    splits = (k_tiles + chunk_tiles - 1) / chunk_tiles;
    if (splits_inout) *splits_inout = splits;
    if (splits > 1 && !ep.partial) return cudaErrorInvalidValue;
    if (splits == 1) ep.partial = nullptr;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
    gemm_bf16_tn_kernel<BN, STAGES, BKR, false, AW><<<grid, THREADS, L::TOTAL, stream>>>(ta, tb, M, N, K, chunk_tiles * BKR, 0, ep);
    return cudaGetLastError();
}

// weight gradient of the 3x3 convolution: dW[9*Cin][Cout] = im2col(X)^T dY, X [B][11][11][Cin], dY [B*121][Cout]; split over samples.
// AW = width of an A slab: 64 (Cin % 64 == 0) or 32 (Cin % 32 == 0, 64-byte swizzle)
template <int BN, int STAGES, int AW>
// ep.colsum_partial != null asks for the column sums of dY as well ([splits][Cout], to be summed over the splits by the caller);
// *colsum_fused tells whether the shape allowed it (a spare A slab starting at row 9 Cin).
inline cudaError_t launch_conv_wgrad(const __nv_bfloat16* X, const __nv_bfloat16* dY, long long batch, int Cin, int Cout, Epilogue ep,
                                     int* splits_inout, cudaStream_t stream, bool* colsum_fused = nullptr) {
    constexpr int BKR = 128;
    if (Cin % AW != 0 || Cout % 8 != 0 || batch <= 0) return cudaErrorInvalidValue;
    int splits = splits_inout ? *splits_inout : 1;
    CUtensorMap ta, tb;
    if (!make_tmap_act(&ta, X, (uint64_t)batch, (uint64_t)Cin, AW) ||
        !make_tmap(&tb, dY, (uint64_t)batch * PIXELS, (uint64_t)Cout, (uint64_t)Cout, PIXELS))
        return cudaErrorInvalidValue;
    using L = SmemLayoutTN<BN, STAGES, BKR, AW>;
    static SmemMemo memo{};
    if (cudaError_t e = ensure_smem(memo, gemm_bf16_tn_kernel<BN, STAGES, BKR, true, AW>, L::TOTAL); e != cudaSuccess) return e;
    const int M = 9 * Cin;
    if (splits < 1) splits = 1;
    if (splits > batch) splits = (int)batch;
    const int chunk = (int)((batch + splits - 1) / splits);
    splits = (int)((batch + chunk - 1) / chunk);
    if (splits_inout) *splits_inout = splits;
    if (splits > 1 && !ep.partial) return cudaErrorInvalidValue;
    if (splits == 1) ep.partial = nullptr;
    const bool spare_slab = (M % BM) != 0 && (M % BM) % AW == 0;
    if (!spare_slab) ep.colsum_partial = nullptr;
    if (colsum_fused) *colsum_fused = ep.colsum_partial != nullptr;
    dim3 grid((Cout + BN - 1) / BN, (M + BM - 1) / BM, splits);
    gemm_bf16_tn_kernel<BN, STAGES, BKR, true, AW><<<grid, THREADS, L::TOTAL, stream>>>(ta, tb, M, Cout, (int)batch, chunk, Cin, ep);
    return cudaGetLastError();
}

}  // namespace tc
}  // namespace mq
