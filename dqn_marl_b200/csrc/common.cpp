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

// Windows [w0, w1): expanded eight at a time into a cache-resident staging block, which then leaves for the (pinned, never
// re-read by this thread) destination with 16-byte streaming stores — no read-for-ownership of 2.9 KB per window, i.e. half
// the memory traffic of plain stores.  The destination range is only 8-byte aligned (726 floats per window): the first and
// last partial 16 bytes of the range go out as plain stores.
static void expand_range(const uint32_t* wire, int64_t w0, int64_t w1, float* obs) {
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
