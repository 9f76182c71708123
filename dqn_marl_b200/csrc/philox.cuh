// Keyed random draws on the device: Philox4x32-10, bit-identical to oracle/keyed_draws.py and
// oracle/env_oracle.c.  Every random decision of the reference is addressed by
// (seed; env, tick|episode, person, stream) instead of a position in a global MT19937 stream,
// which is what makes the env step parallel and replayable (DESIGN.md "Random draws").
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mq {

enum : uint32_t {
    STREAM_NOISE0 = 0,    // people.py:290  random.uniform(-.1,.1): stream = dir/2, words 2*(dir%2), +1
    STREAM_HEALTH = 4,    // people.py:69-75 np.random.uniform: words 0,1; people.py:239 shuffle priority: word 2
    STREAM_SPAWN = 16,    // people.py:186-190 randint: stream = 16 + attempt/2, words 2*(attempt%2), +1
    STREAM_AGENT = 32,    // dqn_agent.py:103-104: words 0,1 = u, word 2 = action
    STREAM_SAMPLE = 48    // dqn_agent.py:132: 4 words = Feistel round keys
};

__host__ __device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     uint64_t seed) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Same function with the ten round keys precomputed (k_r = seed words + r * Weyl constants): the env kernels keep them in the
// kernel-parameter constant bank, which turns the key schedule (20 integer adds per call) into constant operands of the xors.
struct PhiloxKeys { uint32_t k[20]; };
__host__ __device__ inline PhiloxKeys philox_keys(uint64_t seed) {
    PhiloxKeys pk;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { pk.k[2 * r] = k0; pk.k[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return pk;
}
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& pk) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ pk.k[2 * r], n2 = h0 ^ c3 ^ pk.k[2 * r + 1];
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// CPython random.random() / NumPy legacy random_sample: ((a>>5)*2^26 + (b>>6)) / 2^53, exact.
__host__ __device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

// k-th element of the keyed permutation of [0, size): 4-round balanced Feistel + cycle walking
// (oracle/keyed_draws.py sample_index).  Sampling WITHOUT replacement, one index per thread.
__host__ __device__ __forceinline__ uint64_t feistel_index(uint64_t k, uint64_t size, uint4 rk) {
    uint32_t bits = 2;
    while (bits < 64 && ((size - 1) >> bits) != 0) ++bits;
    const uint32_t half = (bits + 1) >> 1;
    const uint64_t mask = (1ull << half) - 1ull;
    const uint32_t keys[4] = {rk.x, rk.y, rk.z, rk.w};
    uint64_t x = k;
    for (;;) {
        uint64_t l = x >> half, r = x & mask;
#pragma unroll
        for (int rnd = 0; rnd < 4; ++rnd) {
            uint64_t nl = r;
            r = l ^ ((uint64_t)mix32((uint32_t)r ^ keys[rnd]) & mask);
            l = nl;
        }
        x = (l << half) | r;
        if (x < size) return x;
    }
}

}  // namespace mq
