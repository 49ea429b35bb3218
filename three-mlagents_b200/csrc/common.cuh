// common.cuh — error plumbing shared by all translation units of libtmla.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tmla.h"

void tmla_set_error(const char *fmt, ...);

#define TMLA_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            tmla_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return TMLA_ECUDA;                                                               \
        }                                                                                    \
    } while (0)

#define TMLA_REQUIRE(cond, msg)                                  \
    do {                                                         \
        if (!(cond)) {                                           \
            tmla_set_error("%s: %s", __func__, msg);             \
            return TMLA_EINVAL;                                  \
        }                                                        \
    } while (0)

#define TMLA_LAUNCH_CHECK() TMLA_CUDA(cudaGetLastError())

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// streaming (evict-first) 128-bit store for write-once rollout rows
__device__ __forceinline__ void st_stream_f4(float4 *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
