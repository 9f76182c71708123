// Probe (not product code): what does the tensor pipe sustain for cta_group::1 bf16 MMAs of 128 x N x 16 when nothing else is
// in the way?  One thread per CTA issues a long run of tcgen05.mma on operand tiles that already sit in shared memory
// (K-major, SWIZZLE_128B, random bits), commits once and waits.  Reported per N: cycles per MMA and the whole-chip TFLOP/s
// with one CTA on every SM.  This is the ceiling the GEMM / convolution kernels can be compared with (the cuBLAS figure in
// MEASURED_PEAKS.json uses 2-SM MMAs and is not reachable with cta_group::1 if this probe says so).
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o /tmp/mma_rate_probe scripts/probes/mma_rate_probe.cu -lcuda && /tmp/mma_rate_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../dqn_marl_b200/csrc/gemm_tc.cuh"

using namespace mq::tc;

template <int N>
__global__ void __launch_bounds__(128) rate_kernel(int n_mma, int stages, long long* cycles) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* tiles = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int A_BYTES = 128 * 128, B_BYTES = N * 128, STAGE = A_BYTES + B_BYTES;
    __shared__ uint64_t done;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    // small-magnitude random bf16 values (exponent bits kept near 1.0 so nothing overflows)
    uint32_t x = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x;
    for (int i = threadIdx.x; i < stages * STAGE / 4; i += blockDim.x) {
        x = x * 1664525u + 1013904223u;
        ((uint32_t*)tiles)[i] = 0x3F003F00u | (x & 0x807F807Fu);
    }
    if (threadIdx.x == 0) { mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_ptr;
    if (threadIdx.x == 32) {
        constexpr uint32_t idesc = make_idesc(N);
        const uint64_t a0 = make_smem_desc(smem_u32(tiles), 1024u, 0, 2);
        const uint64_t b0 = a0 + (uint64_t)(A_BYTES >> 4);
        const long long t0 = clock64();
        int s = 0;
        for (int i = 0; i < n_mma; i += 4) {
            const uint64_t off = (uint64_t)((uint32_t)(s * STAGE) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + (uint32_t)((i & 4) ? N % 512 : 0) % 512, a0 + off + 2 * k, b0 + off + 2 * k, idesc, i > 0 ? 1u : 0u);
            if (++s == stages) s = 0;
        }
        umma_commit(&done);
        mbar_wait(&done, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

template <int N>
void run(int grid, int n_mma, int stages) {
    const int smem = stages * (128 * 128 + N * 128) + 2048;
    cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* d; cudaMalloc(&d, grid * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) rate_kernel<N><<<grid, 128, smem>>>(n_mma, stages, d);
    cudaEventRecord(e0);
    rate_kernel<N><<<grid, 128, smem>>>(n_mma, stages, d);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid); cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost);
    double mean = 0; for (long long c : h) mean += (double)c; mean /= grid;
    const double flops = 2.0 * 128 * N * 16 * (double)n_mma * grid;
    printf("N=%3d grid=%3d stages=%d: %7.1f cycles / MMA (%.0f flop/clk/SM), kernel %.3f ms incl. fill -> >= %.0f TFLOP/s whole chip\n", N, grid, stages,
           mean / n_mma, 2.0 * 128 * N * 16 / (mean / n_mma), ms, flops / (ms * 1e-3) / 1e12);
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        run<32>(grid, 1 << 16, 3); run<64>(grid, 1 << 16, 3); run<128>(grid, 1 << 16, 3); run<256>(grid, 1 << 16, 3);
    }
    run<128>(148, 1 << 16, 1); run<64>(148, 1 << 16, 1);
    return 0;
}
