// Probe (not product code): the mechanics of a CTA pair (cta_group::2) MMA on sm_100a, as groundwork for 2-SM tiles.
// A cluster of two CTAs computes C[256][128] = A[256][64] * B[128][64]^T (bf16, K-major, SWIZZLE_128B):
//   * each CTA loads ITS 128 rows of A and ITS 64 rows of B with TMA; the peer's loads complete_tx on the LEADER's
//     mbarrier (cp.async.bulk.tensor ... .cta_group::2 with the barrier address's CTA-rank bit cleared);
//   * the leader's single thread issues tcgen05.mma.cta_group::2 with M = 256, N = 128: each SM reads its own A rows and
//     its half of B from its own shared memory, each SM's TMEM receives its 128 rows of D;
//   * tcgen05.commit.cta_group::2 ... multicast::cluster arrives on the "done" barrier of both CTAs;
//   * each CTA drains its TMEM half.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I include -o /tmp/cta_pair_probe scripts/probes/cta_pair_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../dqn_marl_b200/csrc/gemm_tc.cuh"

using namespace mq::tc;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are counted on the mbarrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
    const uint32_t bar_leader = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(bar_leader), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192) pair_kernel(const __grid_constant__ CUtensorMap ta,
                                                                              const __grid_constant__ CUtensorMap tb, float* C) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* tiles = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* a_smem = tiles;                    // 128 rows x 128 B
    unsigned char* b_smem = tiles + 128 * 128;        // 64 rows x 128 B
    uint64_t* full = (uint64_t*)(b_smem + 64 * 128);
    uint64_t* done = full + 1;
    uint32_t* tmem_ptr = (uint32_t*)(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    if (threadIdx.x == 0) {
        mbar_init(full, 1); mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                // both CTAs' barriers are initialised before anybody signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;
    if (warp == 0 && lane == 0) {
        if (rank == 0) mbar_expect_tx(full, 2 * (128 * 128 + 64 * 128));       // the leader's barrier counts both CTAs' bytes
        tma_load_2d_pair(a_smem, &ta, full, 0, (int)rank * 128);
        tma_load_2d_pair(b_smem, &tb, full, 0, (int)rank * 64);
    } else if (warp == 1 && lane == 0 && rank == 0) {
        mbar_wait(full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // instruction descriptor: M = 256 (the pair), N = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint64_t a0 = make_smem_desc(smem_u32(a_smem) & 0x00FFFFFFu, 1024u, 0, 2), b0 = make_smem_desc(smem_u32(b_smem) & 0x00FFFFFFu, 1024u, 0, 2);
        for (int k = 0; k < 4; ++k) umma_bf16_pair(tmem_base, a0 + 2 * k, b0 + 2 * k, idesc, k > 0 ? 1u : 0u);
        umma_commit_pair(done);
    } else if (warp >= 2) {
        const int q = warp & 3;
        mbar_wait(done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            for (int j = 0; j < 32; ++j) C[((size_t)rank * 128 + q * 32 + lane) * 128 + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
}

int main() {
    const int M = 256, N = 128, K = 64;
    std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) hA[r * K + k] = __float2bfloat16((float)((r * 7 + k * 3) % 17 - 8));
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hB[n * K + k] = __float2bfloat16((float)((n * 5 + k * 11) % 13 - 6));
    __nv_bfloat16 *dA, *dB; float* dC;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dC, M * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xFF, M * N * 4);
    CUtensorMap ta, tb;
    if (!make_tmap(&ta, dA, M, K, K, 128) || !make_tmap(&tb, dB, N, K, K, 64)) { printf("tensor map failed\n"); return 1; }
    const int smem = 128 * 128 + 64 * 128 + 1024 + 256;
    cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    pair_kernel<<<2, 192, smem>>>(ta, tb, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> hC(M * N);
    cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float ref = 0.f;
            for (int k = 0; k < K; ++k) ref += __bfloat162float(hA[m * K + k]) * __bfloat162float(hB[n * K + k]);
            if (hC[m * N + n] != ref) { if (bad < 5) printf("C[%d][%d] = %g, expected %g\n", m, n, hC[m * N + n], ref); ++bad; }
        }
    printf("cta_group::2 pair MMA 256x128x64: %d wrong of %d\n", bad, M * N);
    return bad != 0;
}
