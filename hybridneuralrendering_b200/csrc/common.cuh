// Shared helpers for libhnr (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define HNR_OK 0
#define HNR_ERR_CUDA (-1)
#define HNR_ERR_ARG (-2)
#define HNR_ERR_UNSUPPORTED (-3)

extern "C" void hnr_set_error(const char* msg);

#define HNR_CHECK_ARG(cond, msg)                                  \
    do {                                                          \
        if (!(cond)) {                                            \
            hnr_set_error(msg);                                   \
            return HNR_ERR_ARG;                                   \
        }                                                         \
    } while (0)

#define HNR_CHECK_LAUNCH(name)                                                                  \
    do {                                                                                        \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess) {                                                               \
            char buf__[256];                                                                    \
            snprintf(buf__, sizeof(buf__), "%s: %s", name, cudaGetErrorString(e__));            \
            hnr_set_error(buf__);                                                               \
            return HNR_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define HNR_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            char buf__[256];                                                                    \
            snprintf(buf__, sizeof(buf__), "%s: %s", #call, cudaGetErrorString(e__));           \
            hnr_set_error(buf__);                                                               \
            return HNR_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

static inline int64_t hnr_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

#define HNR_NUM_SMS 148

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : 0.01f * x; }

// activation codes shared with the host side
enum { HNR_ACT_NONE = 0, HNR_ACT_LRELU = 1, HNR_ACT_SIGMOID = 2, HNR_ACT_COLOR = 3 };

__device__ __forceinline__ float apply_act(float x, int act) {
    switch (act) {
        case HNR_ACT_LRELU: return lrelu(x);
        case HNR_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
        case HNR_ACT_COLOR: return (1.f / (1.f + expf(-x))) * 1.002f - 0.001f;
        default: return x;
    }
}

// derivative of the activation expressed through its OUTPUT y
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
    switch (act) {
        case HNR_ACT_LRELU: return y > 0.f ? 1.f : 0.01f;
        case HNR_ACT_SIGMOID: return y * (1.f - y);
        case HNR_ACT_COLOR: {
            float s = (y + 0.001f) / 1.002f;
            return 1.002f * s * (1.f - s);
        }
        default: return 1.f;
    }
}
