// Shared helpers for libfwi_b200.so (error plumbing, small device utilities).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include "../../include/fwi_b200.h"

namespace fwi {

void set_error(const char* fmt, ...);

#define FWI_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            ::fwi::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,              \
                             cudaGetErrorString(e__));                                         \
            return (e__ == cudaErrorMemoryAllocation) ? FWI_ENOMEM : FWI_ECUDA;                \
        }                                                                                      \
    } while (0)

#define FWI_REQUIRE(cond, ...)                                                                 \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            ::fwi::set_error(__VA_ARGS__);                                                     \
            return FWI_EINVAL;                                                                 \
        }                                                                                      \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace fwi
