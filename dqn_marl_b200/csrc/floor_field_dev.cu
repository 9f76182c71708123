// Static floor field on the device, for batches of layouts.  Same result as mq_floor_field (floor_field.cpp) /
// Map.Init_Potential (Louvre_Evacuation/envs/map.py:127-148), bit for bit: the reference's Dijkstra computes the least fixed
// point of  d[v] = min_u fl(d[u] + c(u, v)),  d[exit] = 1,  over the 8-connected valid cells (c = 1.0 on the axes, 1.4 on the
// diagonals, fl = float64 rounding).  float64 addition is monotone, so that fixed point is unique and ANY monotone relaxation
// order reaches it; here every cell of every layout relaxes against its eight neighbours in place, sweep after sweep, until a
// sweep changes nothing.  Values only ever decrease towards the fixed point, aligned 8-byte loads / stores are single
// transactions, so the in-place (chaotic) sweeps need no ordering.  Init-time code: the call synchronises its stream to read
// the "changed" flag.
#include <cuda_runtime.h>
#include <cstdint>
#include "common.h"

namespace mq {

__global__ void ff_init_kernel(int G, long long total, const int32_t* __restrict__ exits, int max_exits, const int32_t* __restrict__ n_exits,
                               int stride, double* __restrict__ d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int lay = (int)(idx / G), c = (int)(idx - (long long)lay * G);
    double v = __longlong_as_double(0x7FF0000000000000LL);       // +inf
    const int ne = n_exits[lay];
    for (int k = 0; k < ne; ++k) {
        const int ex = exits[((size_t)lay * max_exits + k) * 2], ey = exits[((size_t)lay * max_exits + k) * 2 + 1];
        if (ex * stride + ey == c) v = 1.0;                      // map.py:131
    }
    d[idx] = v;
}

__global__ void ff_sweep_kernel(int L, int W, int G, long long total, const uint8_t* __restrict__ wall, double* d, int* changed) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int stride = W + 2;
    const long long lay0 = idx / G * G;
    const int c = (int)(idx - lay0);
    const int x = c / stride, y = c - x * stride;
    // Map.Check_Valid on the pre-search grid (map.py:85-92): only cells inside 1..L x 1..W that are not walls are ever assigned
    if (x < 1 || x > L || y < 1 || y > W || wall[idx]) return;
    const volatile double* dv = d;
    double best = dv[idx];
    const double before = best;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int dx = (i == 0 || i == 4 || i == 7) ? 1 : ((i == 2 || i == 5 || i == 6) ? -1 : 0);
        const int dy = (i == 3 || i == 6 || i == 7) ? 1 : ((i == 1 || i == 4 || i == 5) ? -1 : 0);
        const double du = dv[lay0 + (x + dx) * stride + (y + dy)];       // neighbours of a valid cell are inside the (L+2) x (W+2) array
        const double cand = du + (i < 4 ? 1.0 : 1.4);                    // current_dist + cost (map.py:137-141); inf stays inf
        if (cand < best) best = cand;
    }
    if (best < before) { d[idx] = best; *changed = 1; }
}

__global__ void ff_add_kernel(long long total, const double* __restrict__ add_term, double* __restrict__ d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const double v = d[idx];
    if (v != __longlong_as_double(0x7FF0000000000000LL)) d[idx] = v + add_term[idx];      // map.py:145-147
}

// dp5 / cellinfo of every layout from its floor field (what dqn_marl_b200/layout.py Layout.build derives on the host):
//   cellinfo bit0 Check_Valid (map.py:85-92), bit1 obs channel 3 = invalid or in barrier_list (evacuation_env.py:109),
//   bit2 obs channel 4 = the exit cell (:113), bit3 checkSavefy = Chebyshev distance <= 1 of any exit (map.py:109-112);
//   dp5[c][d] = (space[c] - space[c + MoveTO[d]]) * 5.0 for valid pairs (people.py:270,288), -inf otherwise.
__global__ void layout_tables_kernel(int L, int W, int G, long long total, const double* __restrict__ space, const uint8_t* __restrict__ barrier,
                                     const int32_t* __restrict__ exits, int max_exits, const int32_t* __restrict__ n_exits,
                                     const int32_t* __restrict__ obs_exit, double* __restrict__ dp5, uint8_t* __restrict__ cellinfo) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int stride = W + 2;
    const int lay = (int)(idx / G);
    const long long lay0 = (long long)lay * G;
    const int c = (int)(idx - lay0);
    const int x = c / stride, y = c - x * stride;
    const double INF = __longlong_as_double(0x7FF0000000000000LL);
    const bool interior = x >= 1 && x <= L && y >= 1 && y <= W;
    const double sc = space[idx];
    const bool valid = interior && sc != INF;
    unsigned info = valid ? 1u : 0u;
    if (!valid || barrier[idx]) info |= 2u;
    if (x == obs_exit[lay * 2] && y == obs_exit[lay * 2 + 1]) info |= 4u;
    const int ne = n_exits[lay];
    for (int k = 0; k < ne; ++k) {
        const int ex = exits[((size_t)lay * max_exits + k) * 2], ey = exits[((size_t)lay * max_exits + k) * 2 + 1];
        if (abs(x - ex) <= 1 && abs(y - ey) <= 1) info |= 8u;
    }
    cellinfo[idx] = (uint8_t)info;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int dx = (i == 0 || i == 4 || i == 7) ? 1 : ((i == 2 || i == 5 || i == 6) ? -1 : 0);
        const int dy = (i == 3 || i == 6 || i == 7) ? 1 : ((i == 1 || i == 4 || i == 5) ? -1 : 0);
        double v = -INF;
        if (valid) {
            const int nx = x + dx, ny = y + dy;
            const double sn = space[lay0 + nx * stride + ny];
            if (nx >= 1 && nx <= L && ny >= 1 && ny <= W && sn != INF) v = (sc - sn) * 5.0;
        }
        dp5[idx * 8 + i] = v;
    }
}

}  // namespace mq

extern "C" int mq_layout_tables_device(int32_t L, int32_t W, int32_t n_layouts, const double* space, const uint8_t* barrier, const int32_t* exits,
                                       int32_t max_exits, const int32_t* n_exits, const int32_t* obs_exit, double* dp5_out, uint8_t* cellinfo_out,
                                       void* stream) {
    MQ_REQUIRE(L > 0 && W > 0 && n_layouts > 0 && space && barrier && exits && max_exits > 0 && n_exits && obs_exit && dp5_out && cellinfo_out,
               "mq_layout_tables_device: bad argument");
    MQ_ON_DEVICE_OF(space);
    const int G = (L + 2) * (W + 2);
    const long long total = (long long)G * n_layouts;
    mq::layout_tables_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(L, W, G, total, space, barrier, exits, max_exits, n_exits,
                                                                                              obs_exit, dp5_out, cellinfo_out);
    MQ_CUDA(cudaGetLastError());
    return MQ_OK;
}

extern "C" int mq_floor_field_device(int32_t L, int32_t W, int32_t n_layouts, const uint8_t* wall, const int32_t* exits, int32_t max_exits,
                                     const int32_t* n_exits, const double* add_term, double* space_out, int32_t* sweeps_out, void* stream) {
    MQ_REQUIRE(L > 0 && W > 0 && n_layouts > 0 && wall && exits && max_exits > 0 && n_exits && space_out, "mq_floor_field_device: bad argument");
    MQ_ON_DEVICE_OF(wall);
    cudaStream_t s = (cudaStream_t)stream;
    const int stride = W + 2, G = (L + 2) * stride;
    const long long total = (long long)G * n_layouts;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    int* changed = nullptr;
    MQ_CUDA(cudaMalloc(&changed, sizeof(int)));
    mq::ff_init_kernel<<<blocks, 256, 0, s>>>(G, total, exits, max_exits, n_exits, stride, space_out);
    int sweeps = 0, h_changed = 1;
    const int max_sweeps = 4 * (G + 8);            // a shortest path visits every cell at most once: G sweeps always suffice
    while (h_changed && sweeps < max_sweeps) {
        cudaMemsetAsync(changed, 0, sizeof(int), s);
        for (int k = 0; k < 16; ++k) mq::ff_sweep_kernel<<<blocks, 256, 0, s>>>(L, W, G, total, wall, space_out, changed);
        sweeps += 16;
        cudaError_t e = cudaMemcpyAsync(&h_changed, changed, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { cudaFree(changed); return mq::fail(MQ_ERR_CUDA, "mq_floor_field_device: %s", cudaGetErrorString(e)); }
    }
    cudaFree(changed);
    if (h_changed) return mq::fail(MQ_ERR_CUDA, "mq_floor_field_device: no fixed point after %d sweeps", sweeps);
    if (add_term) mq::ff_add_kernel<<<blocks, 256, 0, s>>>(total, add_term, space_out);
    if (sweeps_out) *sweeps_out = sweeps;
    MQ_CUDA(cudaGetLastError());
    return MQ_OK;
}
