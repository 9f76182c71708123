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
#include <cstring>
#include <thread>
#include <vector>

static void expand_range(const uint32_t* wire, int64_t w0, int64_t w1, float* obs) {
    // (b1, b3, b4) as floats for the 8 combinations of the three plane bits of a cell
    static const float lut[8][3] = {{0.f, 0.f, 0.f}, {1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {1.f, 1.f, 0.f},
                                    {0.f, 0.f, 1.f}, {1.f, 0.f, 1.f}, {0.f, 1.f, 1.f}, {1.f, 1.f, 1.f}};
    constexpr int CELLS = MQ_OBS_WIN * MQ_OBS_WIN;
    for (int64_t w = w0; w < w1; ++w) {
        const uint32_t* rec = wire + w * MQ_OBS_WIRE_WORDS;
        float* o = obs + w * MQ_OBS_SIZE;
        float v2[CELLS];
        std::memcpy(v2, rec, sizeof(v2));
        int c = 0;
        for (int word = 0; word < 4; ++word) {
            uint32_t p1 = rec[121 + word], p3 = rec[125 + word], p4 = rec[129 + word];
            const int end = word == 3 ? CELLS : 32 * (word + 1);
            for (; c + 1 < end; c += 2) {                                   // two cells = 12 floats = three 16-byte stores
                const float* a = lut[(p1 & 1u) | ((p3 & 1u) << 1) | ((p4 & 1u) << 2)];
                const float* b = lut[((p1 >> 1) & 1u) | (((p3 >> 1) & 1u) << 1) | (((p4 >> 1) & 1u) << 2)];
                const float t[12] = {0.f, a[0], v2[c], a[1], a[2], 0.f, 0.f, b[0], v2[c + 1], b[1], b[2], 0.f};
                std::memcpy(o + 6 * c, t, sizeof(t));
                p1 >>= 2; p3 >>= 2; p4 >>= 2;
            }
            if (c < end) {                                                   // cell 120, the odd one out
                const float* a = lut[(p1 & 1u) | ((p3 & 1u) << 1) | ((p4 & 1u) << 2)];
                const float t[6] = {0.f, a[0], v2[c], a[1], a[2], 0.f};
                std::memcpy(o + 6 * c, t, sizeof(t));
                ++c;
            }
        }
        o[60 * MQ_OBS_CH + 5] = 1.f;                                         // evacuation_env.py:116-117 (i == 5 and j == 5)
    }
}

extern "C" int mq_obs_wire_expand(const uint32_t* wire, int64_t n_windows, float* obs_out, int32_t n_threads) {
    MQ_REQUIRE(wire && obs_out && n_windows >= 0, "mq_obs_wire_expand: bad argument");
    if (n_threads <= 0) {
        const int64_t hw = std::max<int64_t>(1, (int64_t)std::thread::hardware_concurrency());
        n_threads = (int32_t)std::min<int64_t>(hw, std::max<int64_t>(1, n_windows / 1024));
    }
    if (n_threads == 1 || n_windows < 2 * n_threads) { expand_range(wire, 0, n_windows, obs_out); return MQ_OK; }
    std::vector<std::thread> pool;
    const int64_t per = (n_windows + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        const int64_t a = t * per, b = std::min<int64_t>(n_windows, a + per);
        if (a < b) pool.emplace_back(expand_range, wire, a, b, obs_out);
    }
    for (auto& th : pool) th.join();
    return MQ_OK;
}
