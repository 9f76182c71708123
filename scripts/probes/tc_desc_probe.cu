// Probe (not product code): does a K-major SWIZZLE_128B UMMA smem descriptor whose start address is offset by a number of
// ROWS that is not a multiple of 8 (r0 * 128 B) read rows r0 .. r0+127 of a tile that TMA wrote with its 1024-B-aligned
// swizzle pattern?  Tried with base_offset = 0 and base_offset = (start >> 7) & 7.  This decides whether one padded
// activation tile in shared memory can feed all nine taps of a 3x3 convolution through shifted descriptors.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o /tmp/tc_desc_probe scripts/probes/tc_desc_probe.cu -lcuda && /tmp/tc_desc_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../dqn_marl_b200/csrc/gemm_tc.cuh"

using namespace mq::tc;

__global__ void __launch_bounds__(192) probe_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, int r0,
                                                     int use_base_offset, float* C) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* tiles = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* a_smem = tiles;                    // 256 rows x 128 B
    unsigned char* b_smem = tiles + 256 * 128;        // 64 rows x 128 B
    uint64_t* bar = (uint64_t*)(b_smem + 64 * 128);
    uint64_t* done = bar + 1;
    uint32_t* tmem_ptr = (uint32_t*)(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1); mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;
    if (warp == 0 && lane == 0) {
        mbar_expect_tx(bar, 256 * 128 + 64 * 128);
        tma_load_2d(a_smem, &ta, bar, 0, 0);
        tma_load_2d(b_smem, &tb, bar, 0, 0);
    } else if (warp == 1 && lane == 0) {
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(a_smem) + (uint32_t)r0 * 128u, b_addr = smem_u32(b_smem);
        for (int k = 0; k < 4; ++k) {
            uint64_t adesc = make_smem_desc(a_addr + k * 32, 1024u, 0, 2);
            if (use_base_offset) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;
            const uint64_t bdesc = make_smem_desc(b_addr + k * 32, 1024u, 0, 2);
            umma_bf16(tmem_base, adesc, bdesc, make_idesc(64), k > 0 ? 1u : 0u);
        }
        umma_commit(done);
    } else if (warp >= 2) {
        const int q = warp & 3;
        mbar_wait(done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            for (int j = 0; j < 32; ++j) C[(q * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
}

int main() {
    const int R = 256, K = 64;
    std::vector<__nv_bfloat16> hA(R * K), hB(64 * K);
    for (int r = 0; r < R; ++r) for (int k = 0; k < K; ++k) hA[r * K + k] = __float2bfloat16((float)((r * 7 + k * 3) % 251 - 125));
    for (int n = 0; n < 64; ++n) for (int k = 0; k < K; ++k) hB[n * K + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    __nv_bfloat16 *dA, *dB; float* dC;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dC, 128 * 64 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap ta, tb;
    if (!make_tmap(&ta, dA, R, K, K, 256) || !make_tmap(&tb, dB, 64, K, K, 64)) { printf("tensor map failed\n"); return 1; }
    const int smem = 256 * 128 + 64 * 128 + 1024 + 256;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> hC(128 * 64);
    const int offs[] = {0, 1, 2, 3, 5, 8, 11, 12, 13, 16, 25, 26};
    for (int mode = 0; mode < 2; ++mode) {
        for (int r0 : offs) {
            cudaMemset(dC, 0xFF, 128 * 64 * 4);
            probe_kernel<<<1, 192, smem>>>(ta, tb, r0, mode, dC);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d r0 %d: %s\n", mode, r0, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0, shifted_rows_ok = 0;
            for (int m = 0; m < 128; ++m) {
                bool row_ok = true;
                for (int n = 0; n < 64; ++n) if (hC[m * 64 + n] != __bfloat162float(hA[(r0 + m) * K + n])) { row_ok = false; ++bad; }
                shifted_rows_ok += row_ok;
            }
            printf("base_offset %s  row offset %2d: %4d wrong elements, %3d / 128 rows exact\n", mode ? "(addr>>7)&7" : "0          ", r0, bad, shifted_rows_ok);
        }
    }
    return 0;
}
