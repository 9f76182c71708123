#include <algorithm>
#include "common.h"

namespace mq {

char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace mq

extern "C" const char* mq_last_error(void) { return mq::err_buf(); }
extern "C" int mq_abi_version(void) { return MQ_ABI_VERSION; }

// ---------------------------------------------------------------------------------------------------------------------
// Host side of the compact observation wire format (include/marl_b200.h: mq_env_set_obs_wire): dense f32 windows from the
// 544-byte records, with a few host threads.  Pure unpacking: the values were computed by the env kernel.
// ---------------------------------------------------------------------------------------------------------------------
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

// one window (136 words) -> 726 floats at o (16-byte aligned or not)
#if defined(__SSE2__)
// 12 floats of a PAIR of cells for the 64 combinations of their plane bits (b1, b3, b4 of cell c in bits 0, 2, 4; of cell
// c + 1 in bits 1, 3, 5), channel-2 slots left zero: [0 b1 . b3 | b4 0 0 b1' | . b3' b4' 0]
struct PairLut {
    alignas(16) float v[64][12];
    PairLut() {
        for (int i = 0; i < 64; ++i) {
            const float a1 = (float)(i & 1), b1 = (float)((i >> 1) & 1), a3 = (float)((i >> 2) & 1), b3 = (float)((i >> 3) & 1);
            const float a4 = (float)((i >> 4) & 1), b4 = (float)((i >> 5) & 1);
            const float t[12] = {0.f, a1, 0.f, a3, a4, 0.f, 0.f, b1, 0.f, b3, b4, 0.f};
            std::memcpy(v[i], t, sizeof(t));
        }
    }
};
static const PairLut g_pair_lut;
static inline void expand_window(const uint32_t* rec, float* o) {
    constexpr int CELLS = MQ_OBS_WIN * MQ_OBS_WIN;
    int c = 0;
    for (int word = 0; word < 4; ++word) {
        uint32_t p1 = rec[121 + word], p3 = rec[125 + word], p4 = rec[129 + word];
        const int end = word == 3 ? CELLS - 1 : 32 * (word + 1);         // cell 120 is handled after the loop
        for (; c < end; c += 2) {
            const float* l = g_pair_lut.v[(p1 & 3u) | ((p3 & 3u) << 2) | ((p4 & 3u) << 4)];
            const __m128 v2 = _mm_castsi128_ps(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(rec + c)));     // [v2(c), v2(c+1), 0, 0]
            const __m128 x0 = _mm_or_ps(_mm_load_ps(l), _mm_shuffle_ps(v2, v2, _MM_SHUFFLE(2, 0, 2, 2)));       // lane 2 = v2(c)
            const __m128 x2 = _mm_or_ps(_mm_load_ps(l + 8), _mm_shuffle_ps(v2, v2, _MM_SHUFFLE(2, 2, 2, 1)));   // lane 0 = v2(c+1)
            _mm_storeu_ps(o + 6 * c, x0);
            _mm_storeu_ps(o + 6 * c + 4, _mm_load_ps(l + 4));
            _mm_storeu_ps(o + 6 * c + 8, x2);
            p1 >>= 2; p3 >>= 2; p4 >>= 2;
        }
        if (word == 3) {                                                 // c == 120: bit 24 of the fourth plane words
            float v;
            std::memcpy(&v, rec + 120, 4);
            const float t[6] = {0.f, (float)(p1 & 1u), v, (float)(p3 & 1u), (float)(p4 & 1u), 0.f};
            std::memcpy(o + 6 * 120, t, sizeof(t));
        }
    }
    o[60 * MQ_OBS_CH + 5] = 1.f;                                         // evacuation_env.py:116-117 (i == 5 and j == 5)
}
#else
static inline void expand_window(const uint32_t* rec, float* o) {
    constexpr int CELLS = MQ_OBS_WIN * MQ_OBS_WIN;
    for (int c = 0; c < CELLS; ++c) {
        const int word = c >> 5, bit = c & 31;
        float v2;
        std::memcpy(&v2, rec + c, 4);
        // channel 0 == space / inf == 0 (quirk Q1); channel 5 only at the centre cell
        const float t[6] = {0.f, (float)((rec[121 + word] >> bit) & 1u), v2, (float)((rec[125 + word] >> bit) & 1u),
                            (float)((rec[129 + word] >> bit) & 1u), c == 60 ? 1.f : 0.f};
        std::memcpy(o + 6 * c, t, sizeof(t));
    }
}
#endif

// ---- AVX-512 form (runtime-dispatched; the translation unit itself is compiled for baseline x86-64) -----------------------
// Eight cells = 48 output floats = three 64-byte vectors.  The channel-2 values of the eight cells are spread to their slots
// (6 i + 2) by one permute per vector; the bits of the three planes are deposited to their slots (6 i + 1 / + 3 / + 4) with PDEP
// into a 48-bit mask whose three 16-bit pieces blend 1.0f into the vectors.  About 20 instructions per eight cells against
// about 60 for the pair-LUT form: the expansion was compute-bound (5 GB/s of output per host thread, linear in the threads),
// which is what capped the host-buffer rate of an 8-GPU box (4 threads per rank).
#if defined(__x86_64__) && defined(__GNUC__)
#define MQ_WIRE_AVX512 1
#include <immintrin.h>
namespace {
struct Avx512Tables {
    alignas(64) int32_t idx[3][16];
    uint16_t ch2[3];
    uint64_t m1 = 0, m3 = 0, m4 = 0;
    Avx512Tables() {
        for (int k = 0; k < 3; ++k) {
            ch2[k] = 0;
            for (int j = 0; j < 16; ++j) {
                const int f = 16 * k + j, cell = f / 6, ch = f % 6;
                idx[k][j] = ch == 2 ? cell : 0;
                if (ch == 2) ch2[k] |= (uint16_t)(1u << j);
            }
        }
        for (int i = 0; i < 8; ++i) { m1 |= 1ull << (6 * i + 1); m3 |= 1ull << (6 * i + 3); m4 |= 1ull << (6 * i + 4); }
    }
};
const Avx512Tables g_t512;
}  // namespace

__attribute__((target("avx512f,avx512vl,bmi2")))
static inline void expand_window_avx512(const uint32_t* rec, float* o) {
    const uint8_t* b1 = reinterpret_cast<const uint8_t*>(rec + 121);
    const uint8_t* b3 = reinterpret_cast<const uint8_t*>(rec + 125);
    const uint8_t* b4 = reinterpret_cast<const uint8_t*>(rec + 129);
    const __m512i i0 = _mm512_load_si512(g_t512.idx[0]), i1 = _mm512_load_si512(g_t512.idx[1]), i2 = _mm512_load_si512(g_t512.idx[2]);
    const __mmask16 c0 = g_t512.ch2[0], c1 = g_t512.ch2[1], c2 = g_t512.ch2[2];
    const __m512 ones = _mm512_set1_ps(1.f);
    for (int g = 0; g < 15; ++g) {                                       // cells 8 g .. 8 g + 7
        const __m512 v = _mm512_castps256_ps512(_mm256_loadu_ps(reinterpret_cast<const float*>(rec + 8 * g)));
        const uint64_t m = _pdep_u64(b1[g], g_t512.m1) | _pdep_u64(b3[g], g_t512.m3) | _pdep_u64(b4[g], g_t512.m4);
        float* q = o + 48 * g;
        _mm512_storeu_ps(q, _mm512_mask_mov_ps(_mm512_maskz_permutexvar_ps(c0, i0, v), (__mmask16)m, ones));
        _mm512_storeu_ps(q + 16, _mm512_mask_mov_ps(_mm512_maskz_permutexvar_ps(c1, i1, v), (__mmask16)(m >> 16), ones));
        _mm512_storeu_ps(q + 32, _mm512_mask_mov_ps(_mm512_maskz_permutexvar_ps(c2, i2, v), (__mmask16)(m >> 32), ones));
    }
    float v;                                                             // cell 120: bit 0 of byte 15 of every plane
    std::memcpy(&v, rec + 120, 4);
    const float t[6] = {0.f, (float)(b1[15] & 1u), v, (float)(b3[15] & 1u), (float)(b4[15] & 1u), 0.f};
    std::memcpy(o + 6 * 120, t, sizeof(t));
    o[60 * MQ_OBS_CH + 5] = 1.f;                                         // evacuation_env.py:116-117 (i == 5 and j == 5)
}

// as expand_range below, with 64-byte streaming stores
__attribute__((target("avx512f,avx512vl,bmi2")))
static void expand_range_avx512(const uint32_t* wire, int64_t w0, int64_t w1, float* obs) {
    constexpr int BLOCK = 8;
    alignas(64) float stage[BLOCK * MQ_OBS_SIZE + 16];
    for (int64_t w = w0; w < w1; w += BLOCK) {
        const int n = (int)std::min<int64_t>(BLOCK, w1 - w);
        float* dst = obs + w * MQ_OBS_SIZE;
        const size_t total = (size_t)n * MQ_OBS_SIZE;                    // floats
        const size_t mis = ((uintptr_t)dst & 63u) / 4;                   // stage + mis has the 64-byte alignment of dst
        for (int k = 0; k < n; ++k) expand_window_avx512(wire + (w + k) * MQ_OBS_WIRE_WORDS, stage + mis + (size_t)k * MQ_OBS_SIZE);
        size_t k = 0;
        const size_t head = mis ? 16 - mis : 0;
        for (; k < head && k < total; ++k) dst[k] = stage[mis + k];
        for (; k + 16 <= total; k += 16) _mm512_stream_ps(dst + k, _mm512_load_ps(stage + mis + k));
        for (; k < total; ++k) dst[k] = stage[mis + k];
    }
    _mm_sfence();
}

static bool wire_use_avx512() {
    static const bool use = [] {
        const char* v = getenv("MQ_WIRE_ISA");                           // A/B knob: "sse2" forces the baseline form
        if (v && std::strcmp(v, "sse2") == 0) return false;
        return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("bmi2");
    }();
    return use;
}
#endif

// Windows [w0, w1): expanded eight at a time into a cache-resident staging block, which then leaves for the (pinned, never
// re-read by this thread) destination with 16-byte streaming stores — no read-for-ownership of 2.9 KB per window, i.e. half
// the memory traffic of plain stores.  The destination range is only 8-byte aligned (726 floats per window): the first and
// last partial 16 bytes of the range go out as plain stores.
static void expand_range(const uint32_t* wire, int64_t w0, int64_t w1, float* obs) {
#ifdef MQ_WIRE_AVX512
    if (((uintptr_t)obs & 3u) == 0 && wire_use_avx512()) { expand_range_avx512(wire, w0, w1, obs); return; }
#endif
    constexpr int BLOCK = 8;
    alignas(64) float stage[BLOCK * MQ_OBS_SIZE + 4];
    for (int64_t w = w0; w < w1; w += BLOCK) {
        const int n = (int)std::min<int64_t>(BLOCK, w1 - w);
        float* dst = obs + w * MQ_OBS_SIZE;
        const size_t total = (size_t)n * MQ_OBS_SIZE;                    // floats
#if defined(__SSE2__)
        // stage so that stage + mis has the alignment of dst: element k of the block sits at stage[mis + k]
        const size_t mis = ((uintptr_t)dst & 15u) / 4;                   // 0 or 2 (8-byte aligned rows)
        for (int k = 0; k < n; ++k) expand_window(wire + (w + k) * MQ_OBS_WIRE_WORDS, stage + mis + (size_t)k * MQ_OBS_SIZE);
        size_t k = 0;
        const size_t head = mis ? 4 - mis : 0;                           // floats up to the first 16-byte boundary of dst
        for (; k < head && k < total; ++k) dst[k] = stage[mis + k];
        for (; k + 4 <= total; k += 4)
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + k), _mm_load_si128(reinterpret_cast<const __m128i*>(stage + mis + k)));
        for (; k < total; ++k) dst[k] = stage[mis + k];
#else
        for (int k = 0; k < n; ++k) expand_window(wire + (w + k) * MQ_OBS_WIRE_WORDS, dst + (size_t)k * MQ_OBS_SIZE);
#endif
    }
#if defined(__SSE2__)
    _mm_sfence();
#endif
}

// persistent workers: an expansion is a sub-millisecond job issued every env step, thread creation per call would cost as
// much as the work
namespace {
class ExpandPool {
public:
    static ExpandPool& get() { static ExpandPool p; return p; }
    void run(const uint32_t* wire, int64_t n_windows, float* obs, int n_threads) {
        std::unique_lock<std::mutex> call(call_mu_);                     // one expansion at a time per process
        grow(n_threads - 1);
        const int64_t per = ((n_windows + n_threads - 1) / n_threads + 7) & ~(int64_t)7;
        {
            std::lock_guard<std::mutex> lk(mu_);
            wire_ = wire; obs_ = obs; n_ = n_windows; per_ = per; parts_ = n_threads; next_ = 1; pending_ = n_threads - 1;
            ++epoch_;
        }
        cv_.notify_all();
        expand_range(wire, 0, std::min<int64_t>(n_windows, per), obs);   // the caller takes part 0
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }
private:
    ExpandPool() = default;
    ~ExpandPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void grow(int n) {
        while ((int)workers_.size() < n) workers_.emplace_back([this] { loop(); });
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            int part;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || (epoch_ != seen && next_ < parts_); });
                if (stop_) return;
                part = next_++;
                if (next_ >= parts_) seen = epoch_;
            }
            const int64_t a = (int64_t)part * per_, b = std::min<int64_t>(n_, a + per_);
            if (a < b) expand_range(wire_, a, b, obs_);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::mutex call_mu_, mu_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    const uint32_t* wire_ = nullptr; float* obs_ = nullptr;
    int64_t n_ = 0, per_ = 0;
    int parts_ = 0, next_ = 0, pending_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};
}  // namespace

extern "C" int mq_obs_wire_expand(const uint32_t* wire, int64_t n_windows, float* obs_out, int32_t n_threads) {
    MQ_REQUIRE(wire && obs_out && n_windows >= 0, "mq_obs_wire_expand: bad argument");
    if (n_threads <= 0) {
        const int64_t hw = std::max<int64_t>(1, (int64_t)std::thread::hardware_concurrency());
        n_threads = (int32_t)std::min<int64_t>(hw, std::max<int64_t>(1, n_windows / 1024));
    }
    n_threads = std::min<int32_t>(n_threads, 256);
    if (n_threads == 1 || n_windows < 16 * n_threads) { expand_range(wire, 0, n_windows, obs_out); return MQ_OK; }
    ExpandPool::get().run(wire, n_windows, obs_out, n_threads);
    return MQ_OK;
}
