// Shared helpers for libnbm_b200 (sm_100a).  Error plumbing for the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/nbm_b200.h"

namespace nbm {

void set_error(const char *fmt, ...);   // thread-local message, see capi.cu

inline int cuda_fail(cudaError_t e, const char *what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return NBM_ERR_CUDA;
}

#define NBM_CUDA(call)                                          \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return ::nbm::cuda_fail(e__, #call); \
    } while (0)

#define NBM_REQUIRE(cond, ...)                                  \
    do {                                                        \
        if (!(cond)) {                                          \
            ::nbm::set_error(__VA_ARGS__);                      \
            return NBM_ERR_INVALID;                             \
        }                                                       \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Order-preserving float <-> uint32 map, so float min/max can use integer atomics.
__host__ __device__ inline uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b;
    memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float ordered_to_float(uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

}  // namespace nbm
