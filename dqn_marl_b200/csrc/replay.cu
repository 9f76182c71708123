// Device-resident replay ring — push (scatter) and uniform sampling without replacement (gather).
//
// Replaces DQNAgent.memory = deque(maxlen=memory_size) (Louvre_Evacuation/agents/dqn_agent.py:88-89),
// remember() (:97-99) and the `random.sample` + zip + np.array + FloatTensor + .to(device) block of learn()
// (:132-140).  Storage is fp32 SoA in HBM (5817 B per transition instead of the reference's 11.6 KB of
// pickled float64), never leaves the GPU, and is sampled by a keyed permutation so every index of a batch
// is computed independently by its own warp (no host RNG, no host gather, no H2D copy).
//
// Both kernels are pure HBM copies: one warp moves one 2904 B observation row, 128-bit loads and (phase permitting)
// 128-bit stores, all loads of a row issued before the stores (copy_row).
#include <cstdint>
#include <new>
#include "common.h"
#include "philox.cuh"

namespace mq {

// One warp moves one 2904-byte observation row.  Rows are 726 floats: 8-byte, not 16-byte aligned — a row starts either on a
// 16-byte boundary or 8 bytes after one.  Its 724-float body is read with 128-bit loads from the first 16-byte boundary of the
// SOURCE (181 uint4, all issued before the first store) and the odd 8 bytes — the head of a row that starts off a boundary, the
// tail of one that starts on it — by lane 0.  When source and destination have the same phase the body is stored with 128-bit
// stores too; otherwise (a sampled slot and its batch row differ in parity) each 16 bytes leave as two 64-bit stores.
constexpr int ROW_BODY4 = (MQ_OBS_SIZE - 2) / 4;        // 181 uint4
constexpr int ROW_ITERS = (ROW_BODY4 + 31) / 32;        // 6 loads per lane
static_assert(MQ_OBS_SIZE % 4 == 2, "row = 16-byte body + one float2");

__device__ __forceinline__ void copy_row(const float* __restrict__ src, float* __restrict__ dst, int lane) {
    const bool s8 = (reinterpret_cast<uintptr_t>(src) & 8u) != 0, d8 = (reinterpret_cast<uintptr_t>(dst) & 8u) != 0;
    const int off = s8 ? 2 : 0;                          // floats in front of the 16-byte aligned body (source side)
    const int odd = s8 ? 0 : MQ_OBS_SIZE - 2;            // float offset of the odd float2
    const uint4* sb = reinterpret_cast<const uint4*>(src + off);
    uint4 v[ROW_ITERS];
    float2 e = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < ROW_ITERS; ++k) {
        const int i = lane + 32 * k;
        if (i < ROW_BODY4) v[k] = __ldg(sb + i);
    }
    if (lane == 0) e = __ldg(reinterpret_cast<const float2*>(src + odd));
    if (s8 == d8) {
        uint4* db = reinterpret_cast<uint4*>(dst + off);
#pragma unroll
        for (int k = 0; k < ROW_ITERS; ++k) {
            const int i = lane + 32 * k;
            if (i < ROW_BODY4) db[i] = v[k];
        }
    } else {
        uint2* db = reinterpret_cast<uint2*>(dst + off);
#pragma unroll
        for (int k = 0; k < ROW_ITERS; ++k) {
            const int i = lane + 32 * k;
            if (i < ROW_BODY4) { db[2 * i] = make_uint2(v[k].x, v[k].y); db[2 * i + 1] = make_uint2(v[k].z, v[k].w); }
        }
    }
    if (lane == 0) *reinterpret_cast<float2*>(dst + odd) = e;
}

struct RingView {
    float* state; float* next_state; int32_t* action; float* reward; uint8_t* done;
    long long capacity;
};

// grid: one warp per (transition, {state,next_state}) pair
__global__ void __launch_bounds__(256)
replay_push_kernel(RingView ring, long long cursor, const float* __restrict__ state, const int32_t* __restrict__ action,
                   const double* __restrict__ reward, const float* __restrict__ next_state,
                   const uint8_t* __restrict__ done, long long n) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= 2 * n) return;
    const long long k = w >> 1;
    const long long slot = (cursor + k) % ring.capacity;
    if (w & 1) {
        copy_row(next_state + k * MQ_OBS_SIZE, ring.next_state + slot * MQ_OBS_SIZE, lane);
    } else {
        copy_row(state + k * MQ_OBS_SIZE, ring.state + slot * MQ_OBS_SIZE, lane);
        if (lane == 0) {
            ring.action[slot] = action[k];
            ring.reward[slot] = (float)reward[k];     // torch.FloatTensor(rewards) (dqn_agent.py:138)
            ring.done[slot] = done[k];
        }
    }
}

__global__ void __launch_bounds__(256)
replay_sample_kernel(RingView ring, long long size, long long oldest, uint4 round_keys, const long long* __restrict__ inject,
                     float* __restrict__ state, long long* __restrict__ action, float* __restrict__ reward,
                     float* __restrict__ next_state, uint8_t* __restrict__ done, long long* __restrict__ idx_out,
                     long long B) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= 2 * B) return;
    const long long k = w >> 1;
    // logical index 0 = oldest transition (deque order); without replacement: keyed permutation of [0, size)
    const long long logical = inject ? inject[k] : (long long)feistel_index((uint64_t)k, (uint64_t)size, round_keys);
    const long long slot = (oldest + logical) % ring.capacity;
    if (w & 1) {
        copy_row(ring.next_state + slot * MQ_OBS_SIZE, next_state + k * MQ_OBS_SIZE, lane);
    } else {
        copy_row(ring.state + slot * MQ_OBS_SIZE, state + k * MQ_OBS_SIZE, lane);
        if (lane == 0) {
            action[k] = (long long)ring.action[slot];       // torch.LongTensor(actions) (dqn_agent.py:137)
            reward[k] = ring.reward[slot];
            done[k] = ring.done[slot];
            if (idx_out) idx_out[k] = logical;
        }
    }
}

}  // namespace mq

struct mq_replay {
    mq::RingView ring;
    int device = 0;
    long long size = 0;      // len(deque)
    long long cursor = 0;    // next physical slot to write
    int64_t launches = 0;
};

extern "C" int mq_replay_create(mq_replay** out, int64_t capacity, int32_t device, const mq_replay_store* store) {
    MQ_REQUIRE(out && store && capacity > 0, "mq_replay_create: bad argument");
    MQ_REQUIRE(store->state && store->next_state && store->action && store->reward && store->done,
               "mq_replay_create: storage buffers missing");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return mq::fail(MQ_ERR_CUDA, "mq_replay_create: no CUDA device (this build has no CPU fallback)");
    MQ_REQUIRE(device >= 0 && device < ndev, "mq_replay_create: device %d outside 0..%d", device, ndev - 1);
    mq_replay* rb = new (std::nothrow) mq_replay();
    if (!rb) return mq::fail(MQ_ERR_ALLOC, "mq_replay_create: out of host memory");
    rb->ring = {store->state, store->next_state, store->action, store->reward, store->done, (long long)capacity};
    rb->device = device;
    *out = rb;
    return MQ_OK;
}

extern "C" int mq_replay_destroy(mq_replay* rb) { delete rb; return MQ_OK; }
extern "C" int64_t mq_replay_size(const mq_replay* rb) { return rb ? rb->size : 0; }
extern "C" int64_t mq_replay_cursor(const mq_replay* rb) { return rb ? rb->cursor : 0; }
extern "C" int64_t mq_replay_launch_count(const mq_replay* rb) { return rb ? rb->launches : 0; }

extern "C" int mq_replay_restore(mq_replay* rb, int64_t size, int64_t cursor) {
    MQ_REQUIRE(rb, "mq_replay_restore: null handle");
    MQ_REQUIRE(size >= 0 && size <= rb->ring.capacity && cursor >= 0 && cursor < rb->ring.capacity,
               "mq_replay_restore: size %lld / cursor %lld outside a ring of %lld", (long long)size, (long long)cursor, rb->ring.capacity);
    MQ_REQUIRE(size == rb->ring.capacity || cursor == size % rb->ring.capacity,
               "mq_replay_restore: a ring that is not full writes at slot size (size %lld, cursor %lld)", (long long)size, (long long)cursor);
    rb->size = size; rb->cursor = cursor;
    return MQ_OK;
}

extern "C" int mq_replay_push(mq_replay* rb, const float* state, const int32_t* action, const double* reward,
                              const float* next_state, const uint8_t* done, int64_t n, void* stream) {
    MQ_REQUIRE(rb && state && action && reward && next_state && done, "mq_replay_push: null argument");
    MQ_REQUIRE(n >= 0 && n <= rb->ring.capacity, "mq_replay_push: n=%lld exceeds the capacity %lld", (long long)n, rb->ring.capacity);
    if (n == 0) return MQ_OK;
    MQ_ON_DEVICE(rb->device);
    const long long warps = 2 * n;
    const int blocks = (int)((warps * 32 + 255) / 256);
    mq::replay_push_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rb->ring, rb->cursor, state, action, reward, next_state,
                                                                     done, n);
    MQ_CUDA(cudaGetLastError());
    rb->cursor = (rb->cursor + n) % rb->ring.capacity;
    rb->size = rb->size + n > rb->ring.capacity ? rb->ring.capacity : rb->size + n;
    rb->launches += 1;
    return MQ_OK;
}

extern "C" int mq_replay_sample(mq_replay* rb, int64_t B, uint64_t seed, uint64_t draw_id, const int64_t* inject_idx,
                                float* state, int64_t* action, float* reward, float* next_state, uint8_t* done,
                                int64_t* idx_out, void* stream) {
    MQ_REQUIRE(rb && state && action && reward && next_state && done, "mq_replay_sample: null argument");
    // random.sample raises ValueError when the population is smaller than the sample (dqn_agent.py:132)
    MQ_REQUIRE(B > 0 && B <= rb->size, "mq_replay_sample: sample larger than population (B=%lld, size=%lld)", (long long)B,
               rb->size);
    MQ_ON_DEVICE(rb->device);
    const uint4 rk = mq::philox4x32(0u, (uint32_t)draw_id, 0u, mq::STREAM_SAMPLE, seed);
    const long long oldest = rb->size < rb->ring.capacity ? 0 : rb->cursor;
    const long long warps = 2 * B;
    const int blocks = (int)((warps * 32 + 255) / 256);
    mq::replay_sample_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rb->ring, rb->size, oldest, rk, (const long long*)inject_idx,
                                                                       state, (long long*)action, reward, next_state, done,
                                                                       (long long*)idx_out, B);
    MQ_CUDA(cudaGetLastError());
    rb->launches += 1;
    return MQ_OK;
}
