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
