// Fused dense Adam step for the neural-point tables.  The reference updates every point row every step
// with torch.optim.Adam (models/mvs_points_volumetric_model.py:94-104: one optimiser over xyz / embedding / conf / dir / colour,
// dense semantics: rows with a zero gradient still decay their moments and move).  torch's for-each implementation makes
// several passes over the 39*N floats; this kernel reads p, g, m, v once and writes p, m, v once (28 bytes per element).
// Arithmetic follows torch.optim.Adam (amsgrad off, maximize off):
//     m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
// with optional L2 weight decay folded into g first.
#include "common.cuh"
#include "hnr.h"

namespace {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float b1, float b2, float step_size,
                                                   float inv_bc2_sqrt, float eps, float wd) {
    const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = i4 * 4;
    if (i + 4 <= n) {
        float4 pp = reinterpret_cast<float4*>(p)[i4], gg = reinterpret_cast<const float4*>(g)[i4];
        float4 mm = reinterpret_cast<float4*>(m)[i4], vv = reinterpret_cast<float4*>(v)[i4];
        float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = G[k] + wd * P[k];
            M[k] = M[k] + (1.f - b1) * (gk - M[k]);                 // torch: exp_avg.lerp_(grad, 1 - beta1)
            V[k] = b2 * V[k] + (1.f - b2) * gk * gk;                // torch: exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
            const float denom = sqrtf(V[k]) * inv_bc2_sqrt + eps;
            P[k] = P[k] - step_size * (M[k] / denom);
        }
        reinterpret_cast<float4*>(p)[i4] = pp;
        reinterpret_cast<float4*>(m)[i4] = mm;
        reinterpret_cast<float4*>(v)[i4] = vv;
    } else {
        for (int64_t j = i; j < n; ++j) {
            const float gk = g[j] + wd * p[j];
            const float mj = m[j] + (1.f - b1) * (gk - m[j]);
            const float vj = b2 * v[j] + (1.f - b2) * gk * gk;
            m[j] = mj; v[j] = vj;
            p[j] = p[j] - step_size * (mj / (sqrtf(vj) * inv_bc2_sqrt + eps));
        }
    }
}

// ---- multi-tensor form: one launch for every tensor of a parameter group ----
constexpr int ADAM_MAX_T = 64;
struct AdamMulti {
    float* p[ADAM_MAX_T];
    const float* g[ADAM_MAX_T];
    float* m[ADAM_MAX_T];
    float* v[ADAM_MAX_T];
    int64_t n[ADAM_MAX_T];
    int32_t blk0[ADAM_MAX_T + 1];       // first block of tensor t (a block covers 1024 elements)
    int32_t nt;
};

__global__ void __launch_bounds__(256) adam_multi_kernel(const __grid_constant__ AdamMulti A, float b1, float b2, float step_size,
                                                         float inv_bc2_sqrt, float eps, float wd, const int32_t* __restrict__ guard) {
    // range guard of the split-fp16 forward kernels (ops.status_word): a step whose activations saturated must not be applied.
    // The host raises at its next synchronisation point; until then the parameters stay untouched.
    if (guard && *guard != 0) return;
    int t = 0;
    for (int q = 1; q < A.nt; ++q)
        if ((int)blockIdx.x >= A.blk0[q]) t = q;
    const int64_t n = A.n[t];
    float* __restrict__ p = A.p[t];
    const float* __restrict__ g = A.g[t];
    float* __restrict__ m = A.m[t];
    float* __restrict__ v = A.v[t];
    const int64_t i = ((int64_t)(blockIdx.x - A.blk0[t]) * 256 + threadIdx.x) * 4;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec && i + 4 <= n) {
        const int64_t i4 = i >> 2;
        float4 pp = reinterpret_cast<float4*>(p)[i4], gg = reinterpret_cast<const float4*>(g)[i4];
        float4 mm = reinterpret_cast<float4*>(m)[i4], vv = reinterpret_cast<float4*>(v)[i4];
        float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = G[k] + wd * P[k];
            M[k] = M[k] + (1.f - b1) * (gk - M[k]);
            V[k] = b2 * V[k] + (1.f - b2) * gk * gk;
            P[k] = P[k] - step_size * (M[k] / (sqrtf(V[k]) * inv_bc2_sqrt + eps));
        }
        reinterpret_cast<float4*>(p)[i4] = pp;
        reinterpret_cast<float4*>(m)[i4] = mm;
        reinterpret_cast<float4*>(v)[i4] = vv;
    } else {
        for (int64_t j = i; j < n && j < i + 4; ++j) {
            const float gk = g[j] + wd * p[j];
            const float mj = m[j] + (1.f - b1) * (gk - m[j]);
            const float vj = b2 * v[j] + (1.f - b2) * gk * gk;
            m[j] = mj; v[j] = vj;
            p[j] = p[j] - step_size * (mj / (sqrtf(vj) * inv_bc2_sqrt + eps));
        }
    }
}

}  // namespace

// Adam step over nt <= 64 tensors in ONE launch (same arithmetic as hnr_adam_step; all tensors share the hyper-parameters and the
// step count).  p/g/m/v/n: HOST arrays of device pointers / element counts.  guard (optional device word): the launch is a no-op
// when *guard != 0.
extern "C" int hnr_adam_multi(int64_t nt, float* const* p, const float* const* g, float* const* m, float* const* v, const int64_t* n,
                              float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step, const int32_t* guard,
                              void* stream) {
    HNR_CHECK_ARG(nt >= 0 && nt <= ADAM_MAX_T, "adam_multi: at most 64 tensors per launch");
    HNR_CHECK_ARG(step >= 1, "adam_multi: step counts from 1");
    if (nt == 0) return HNR_OK;
    AdamMulti A{};
    int64_t blk = 0;
    for (int t = 0; t < nt; ++t) {
        A.p[t] = p[t]; A.g[t] = g[t]; A.m[t] = m[t]; A.v[t] = v[t]; A.n[t] = n[t];
        A.blk0[t] = (int32_t)blk;
        blk += hnr_cdiv(n[t], 1024);
        HNR_CHECK_ARG(blk < (1ll << 31), "adam_multi: too many elements");
    }
    A.blk0[nt] = (int32_t)blk;
    A.nt = (int32_t)nt;
    if (blk == 0) return HNR_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_multi_kernel<<<(unsigned)blk, 256, 0, (cudaStream_t)stream>>>(A, beta1, beta2, (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), eps,
                                                                      weight_decay, guard);
    HNR_CHECK_LAUNCH("adam_multi");
    return HNR_OK;
}

// one Adam step over n contiguous fp32 elements (p, g, m, v 16-byte aligned); step = 1-based step count AFTER this update
extern "C" int hnr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int64_t step, void* stream) {
    if (n == 0) return HNR_OK;
    HNR_CHECK_ARG(step >= 1, "adam_step: step counts from 1");
    HNR_CHECK_ARG(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
                  "adam_step: buffers must be 16-byte aligned");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1), inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    const int64_t n4 = hnr_cdiv(n, 4);
    adam_kernel<<<(unsigned)hnr_cdiv(n4, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, beta1, beta2, step_size, inv_bc2_sqrt, eps,
                                                                               weight_decay);
    HNR_CHECK_LAUNCH("adam_step");
    return HNR_OK;
}
