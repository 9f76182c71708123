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

// Every handle is tied to one CUDA device; a caller whose CURRENT device is another one (an agent built on cuda:1 while
// cuda:0 is current, the reference's `DQNAgent(..., torch.device('cuda:1'), cfg)`) must still launch on the handle's device:
// stream handle 0 means "the current device's default stream".  The guard switches for the duration of the call only.
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        int cur = -1;
        err = cudaGetDevice(&cur);
        if (err == cudaSuccess && cur != device) {
            err = cudaSetDevice(device);
            if (err == cudaSuccess) prev = cur;
        }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// device that owns a device pointer (entry points without a handle), -1 if it is not device memory
inline int pointer_device(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}

#define MQ_ON_DEVICE(dev)                                                                                     \
    mq::DeviceGuard _mq_guard(dev);                                                                           \
    if (_mq_guard.err != cudaSuccess)                                                                         \
        return mq::fail(MQ_ERR_CUDA, "cannot switch to CUDA device %d: %s", (int)(dev), cudaGetErrorString(_mq_guard.err))

#define MQ_ON_DEVICE_OF(ptr)                                                                                  \
    const int _mq_pdev = mq::pointer_device(ptr);                                                             \
    if (_mq_pdev < 0) return mq::fail(MQ_ERR_ARG, "%s is not a CUDA device pointer", #ptr);                    \
    MQ_ON_DEVICE(_mq_pdev)

}  // namespace mq
