// Shared host-side plumbing of libmarl_b200.so: status codes, thread-local error text, CUDA checks.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/marl_b200.h"

namespace mq {

char* err_buf();                       // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);

#define MQ_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return mq::fail(MQ_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,  \
                            cudaGetErrorString(_e));                                           \
    } while (0)

#define MQ_REQUIRE(cond, ...)                                    \
    do {                                                         \
        if (!(cond)) return mq::fail(MQ_ERR_ARG, __VA_ARGS__);   \
    } while (0)

}  // namespace mq
