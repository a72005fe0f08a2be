// Compositing (SURVEY.md §8a rows C1 + C2): ray_dist prologue + alpha compositing, forward and
// backward.  Replaces neural_points_volumetric_model.py:331-339 and
// models/rendering/diff_ray_marching.py:508-557 (radiance_render + alpha_blend) of the reference.
//
// One warp per ray; lanes stride over the ray's samples in chunks of 32 so every global access is
// a coalesced 128 B (or 512 B for the 4-channel features) line.  The cummax / cumprod / suffix-sum
// are warp scans with a carry between chunks.  HBM-bound: ~29 B per (ray, sample) forward.
#include "common.cuh"

namespace {

__device__ __forceinline__ float warp_incl_scan_max(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = fmaxf(v, t);
    }
    return v;
}
__device__ __forceinline__ float warp_incl_scan_mul(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v *= t;
    }
    return v;
}
// inclusive suffix sum: out[i] = sum_{j>=i} v[j]
__device__ __forceinline__ float warp_incl_suffix_sum(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += t;
    }
    return v;
}

// ray_dist for sample i given running cummax m_i (inclusive) and the next sample's depth.
__device__ __forceinline__ float seg_len(float m_i, float z_next, bool last, float vz, int unit_mode) {
    float d = last ? vz : (fmaxf(m_i, z_next) - m_i);
    bool bad = d < 1e-8f || (unit_mode && d > 2.f * vz);
    return bad ? vz : d;
}

// feats (R,SR,4) [sigma,r,g,b]; valid (R,SR) u8; z (R,SR,*) with stride zs (camera depth of the sample)
// or, if dist_in != nullptr, precomputed segment lengths (ray_march() drop-in entry).
__global__ void __launch_bounds__(256)
composite_fwd_kernel(const float* __restrict__ feats, const uint8_t* __restrict__ valid, const float* __restrict__ z, int zs,
                     const float* __restrict__ dist_in, const float* __restrict__ bg, float vz, int unit_mode, int R, int SR,
                     float* __restrict__ ray_color, float* __restrict__ opacity, float* __restrict__ accT,
                     float* __restrict__ bweight, float* __restrict__ bgT, float* __restrict__ dist_out) {
    const int lane = threadIdx.x & 31;
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= R) return;
    const float4* f4 = reinterpret_cast<const float4*>(feats) + (size_t)ray * SR;
    const size_t base = (size_t)ray * SR;
    float carry_m = -INFINITY, carry_T = 1.f;
    float cr = 0.f, cg = 0.f, cb = 0.f;
    for (int c0 = 0; c0 < SR; c0 += 32) {
        const int i = c0 + lane;
        const bool in = i < SR;
        float v = in ? (float)valid[base + i] : 0.f;
        float d;
        if (dist_in) {
            d = in ? dist_in[base + i] : 0.f;
        } else {
            float zi = in ? z[(base + i) * zs] : -INFINITY;
            float m = fmaxf(warp_incl_scan_max(zi, lane), carry_m);
            float zn = (i + 1 < SR) ? z[(base + i + 1) * zs] : 0.f;
            d = in ? seg_len(m, zn, i == SR - 1, vz, unit_mode) * v : 0.f;
            carry_m = __shfl_sync(0xffffffffu, m, 31);
        }
        float4 f = in ? f4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float s = f.x * v;
        float o = 1.f - expf(-s * d);
        float a = in ? (1.f - o + 1e-10f) : 1.f;
        float Tincl = warp_incl_scan_mul(a, lane) * carry_T;
        float Texcl = __shfl_up_sync(0xffffffffu, Tincl, 1);
        if (lane == 0) Texcl = carry_T;
        carry_T = __shfl_sync(0xffffffffu, Tincl, 31);
        float w = o * Texcl;
        if (in) {
            opacity[base + i] = o;
            accT[base + i] = Texcl;
            bweight[base + i] = w;
            if (dist_out) dist_out[base + i] = d;
            cr += w * f.y; cg += w * f.z; cb += w * f.w;
        }
    }
    cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
    if (lane == 0) {
        float b0 = bg ? bg[0] : 0.f, b1 = bg ? bg[1] : 0.f, b2 = bg ? bg[2] : 0.f;
        ray_color[ray * 3 + 0] = cr + b0 * carry_T;
        ray_color[ray * 3 + 1] = cg + b1 * carry_T;
        ray_color[ray * 3 + 2] = cb + b2 * carry_T;
        bgT[ray] = carry_T;
    }
}

// backward: recompute opacities from (feats, valid, dist) and use the saved transmittance.
// g_color (R,3) required; g_opacity, g_bw, g_accT (R,SR) and g_bgT (R) optional.  Writes g_feats (R,SR,4).
__global__ void __launch_bounds__(256)
composite_bwd_kernel(const float* __restrict__ feats, const uint8_t* __restrict__ valid, const float* __restrict__ dist,
                     const float* __restrict__ accT, const float* __restrict__ bgT, const float* __restrict__ bg,
                     const float* __restrict__ g_color, const float* __restrict__ g_opacity, const float* __restrict__ g_bgT,
                     const float* __restrict__ g_bw, const float* __restrict__ g_accT, int R, int SR, float* __restrict__ g_feats) {
    const int lane = threadIdx.x & 31;
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= R) return;
    const float4* f4 = reinterpret_cast<const float4*>(feats) + (size_t)ray * SR;
    float4* g4 = reinterpret_cast<float4*>(g_feats) + (size_t)ray * SR;
    const size_t base = (size_t)ray * SR;
    const float gr = g_color[ray * 3], gg = g_color[ray * 3 + 1], gb = g_color[ray * 3 + 2];
    float gTend = (bg ? gr * bg[0] + gg * bg[1] + gb * bg[2] : 0.f) + (g_bgT ? g_bgT[ray] : 0.f);
    // S_i = sum_{k>i} gT_k T_k, k running to `end`; walk chunks from the back
    float carry = gTend * bgT[ray];
    const int nchunk = (SR + 31) / 32;
    for (int c = nchunk - 1; c >= 0; --c) {
        const int i = c * 32 + lane;
        const bool in = i < SR;
        float v = in ? (float)valid[base + i] : 0.f;
        float d = in ? dist[base + i] : 0.f;
        float4 f = in ? f4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float T = in ? accT[base + i] : 0.f;
        float s = f.x * v;
        float e = expf(-s * d);
        float o = 1.f - e;
        float a = 1.f - o + 1e-10f;
        float gw = gr * f.y + gg * f.z + gb * f.w + ((g_bw && in) ? g_bw[base + i] : 0.f);
        float gT_T = in ? (o * gw + (g_accT ? g_accT[base + i] : 0.f)) * T : 0.f;
        float incl = warp_incl_suffix_sum(gT_T, lane);
        float after = __shfl_down_sync(0xffffffffu, incl, 1);
        float S = (lane == 31 ? 0.f : after) + carry;   // strictly-after sum (no cancellation)
        carry += __shfl_sync(0xffffffffu, incl, 0);
        if (in) {
            float go = T * gw - S / a + (g_opacity ? g_opacity[base + i] : 0.f);
            float w = o * T;
            g4[i] = make_float4(v * go * d * e, w * gr, w * gg, w * gb);
        }
    }
}

}  // namespace

extern "C" int hnr_composite_fwd(const float* feats, const uint8_t* valid, const float* z, int z_stride, const float* dist_in,
                                 const float* bg, float vsize_z, int unit_mode, int64_t R, int64_t SR, float* ray_color,
                                 float* opacity, float* acc_trans, float* blend_weight, float* bg_trans, float* dist_out,
                                 void* stream) {
    if (R == 0) return HNR_OK;
    HNR_CHECK_ARG(R > 0 && SR > 0, "composite_fwd: bad shape");
    HNR_CHECK_ARG((z != nullptr) != (dist_in != nullptr), "composite_fwd: pass exactly one of z / dist_in");
    const int threads = 256;
    const int64_t blocks = hnr_cdiv(R * 32, threads);
    composite_fwd_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        feats, valid, z, z_stride, dist_in, bg, vsize_z, unit_mode, (int)R, (int)SR, ray_color, opacity, acc_trans, blend_weight,
        bg_trans, dist_out);
    HNR_CHECK_LAUNCH("composite_fwd");
    return HNR_OK;
}

extern "C" int hnr_composite_bwd(const float* feats, const uint8_t* valid, const float* dist, const float* acc_trans,
                                 const float* bg_trans, const float* bg, const float* g_color, const float* g_opacity,
                                 const float* g_bg_trans, const float* g_blend_weight, const float* g_acc_trans, int64_t R,
                                 int64_t SR, float* g_feats, void* stream) {
    HNR_CHECK_ARG(R >= 0 && SR > 0, "composite_bwd: bad shape");
    if (R == 0) return HNR_OK;
    const int threads = 256;
    const int64_t blocks = hnr_cdiv(R * 32, threads);
    composite_bwd_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        feats, valid, dist, acc_trans, bg_trans, bg, g_color, g_opacity, g_bg_trans, g_blend_weight, g_acc_trans, (int)R, (int)SR,
        g_feats);
    HNR_CHECK_LAUNCH("composite_bwd");
    return HNR_OK;
}
