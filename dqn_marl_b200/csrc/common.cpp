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
    for (int64_t w = w0; w < w1; ++w) {
        const uint32_t* rec = wire + w * MQ_OBS_WIRE_WORDS;
        float* o = obs + w * MQ_OBS_SIZE;
        for (int c = 0; c < MQ_OBS_WIN * MQ_OBS_WIN; ++c) {
            const int word = c >> 5, bit = c & 31;
            float v2;
            std::memcpy(&v2, rec + c, sizeof(float));
            o[0] = 0.f;                                                   // channel 0 == space / inf (quirk Q1)
            o[1] = (float)((rec[121 + word] >> bit) & 1u);
            o[2] = v2;
            o[3] = (float)((rec[125 + word] >> bit) & 1u);
            o[4] = (float)((rec[129 + word] >> bit) & 1u);
            o[5] = c == 60 ? 1.f : 0.f;                                   // evacuation_env.py:116-117 (i == 5 and j == 5)
            o += MQ_OBS_CH;
        }
    }
}

extern "C" int mq_obs_wire_expand(const uint32_t* wire, int64_t n_windows, float* obs_out, int32_t n_threads) {
    MQ_REQUIRE(wire && obs_out && n_windows >= 0, "mq_obs_wire_expand: bad argument");
    if (n_threads <= 0) n_threads = (int32_t)std::min<int64_t>(16, std::max<int64_t>(1, n_windows / 2048));
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
