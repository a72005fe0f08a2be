// Point aggregation glue kernels: everything around the
// dense layers -- neighbour gather straight from the point tables (no (1,R,SR,K,C) tensors are
// materialised), inverse-distance weights, positional encodings written as the MLP input rows,
// the weighted K-sum with the density head, projection into the reference views, the
// pyramid lookup ("bilinear-upsample then truncated nearest pixel", evaluated directly on the
// conv pyramid so the (V,45,H,W) tensor never exists), the learned multi-view blend, and all the
// matching backward kernels (scatter-add of point gradients with red.global.add).
//
// Reference behaviour restated: models/aggregators/point_aggregators.py:825-833, :892-1037,
// :1064-1096, :1188-1217, :1422-1424, :1472-1508; models/helpers/networks.py:175-189;
// models/neural_points/neural_points.py:702-733; models/neural_points_volumetric_model.py:248-310.
//
// These are HBM/L2 gather kernels: one warp per neighbour row (lane = embedding channel, so the
// 128 B embedding row is one coalesced transaction) or per sample.
#include "common.cuh"

#include "hnr.h"
#include "img_common.cuh"

namespace {

constexpr int F_EMB = 32;          // point_features_dim
constexpr int NF_FEAT = 3;         // num_feat_freqs
constexpr int NF_DIST = 5;         // dist_xyz_freq
constexpr int NF_VIEW = 4;         // num_viewdir_freqs
constexpr int X0_W = F_EMB + 2 * NF_FEAT * F_EMB + 2 * NF_DIST * 6;   // 284
constexpr int E_W = 7;
constexpr int HID = 256;
constexpr int X5_W = HID + 2 * NF_VIEW * 3;                            // 280
constexpr int AUX_C = 45;

__device__ __forceinline__ void rot3(const float* m, float x, float y, float z, float& ox, float& oy, float& oz) {
    ox = x * m[0] + y * m[3] + z * m[6];
    oy = x * m[1] + y * m[4] + z * m[7];
    oz = x * m[2] + y * m[5] + z * m[8];
}

__device__ __forceinline__ bool slot_masked(const int32_t* pidx, const uint8_t* mask, int64_t i) {
    return mask ? (mask[i] == 0) : (pidx[i] < 0);
}

// ------------------------------------------------------------------ A1/A2: weights (all samples)
__global__ void nbr_weights_kernel(const float* __restrict__ xyz, const float* __restrict__ conf, const int32_t* __restrict__ pidx,
                                   const uint8_t* __restrict__ mask, const float* __restrict__ loc_w, int64_t S, int K,
                                   float* __restrict__ weight, float* __restrict__ confc, uint8_t* __restrict__ valid) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const float lx = loc_w[s * 3], ly = loc_w[s * 3 + 1], lz = loc_w[s * 3 + 2];
    float sum = 0.f;
    bool any = false;
    for (int k = 0; k < K; ++k) {
        int64_t i = s * K + k;
        bool m = slot_masked(pidx, mask, i);
        int64_t g = max(pidx[i], 0);
        float dx = xyz[g * 3] - lx, dy = xyz[g * 3 + 1] - ly, dz = xyz[g * 3 + 2] - lz;
        float u = m ? 0.f : 1.f / fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-6f);
        weight[i] = u;
        sum += u;
        any |= !m;
        if (confc) confc[i] = conf ? fminf(fmaxf(conf[g], 1e-4f), 1.f) : 1.f;
    }
    float inv = 1.f / fmaxf(sum, 1e-8f);
    for (int k = 0; k < K; ++k) weight[s * K + k] *= inv;
    if (valid) valid[s] = any ? 1 : 0;
}

// ------------------------------------------------------------------ A3 prologue: MLP input rows
// one warp per (valid sample, neighbour) row.
__global__ void __launch_bounds__(256)
nbr_features_kernel(const float* __restrict__ xyz, const float* __restrict__ xyz_pers, const float* __restrict__ emb,
                    const float* __restrict__ color, const float* __restrict__ dir, const int32_t* __restrict__ pidx,
                    const int32_t* __restrict__ vlist, const float* __restrict__ loc_w, const float* __restrict__ loc_pers,
                    const float* __restrict__ raydirs, const float* __restrict__ cam, int64_t rows, int K, float* __restrict__ X0,
                    float* __restrict__ E) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const int64_t v = row / K;
    const int k = (int)(row - v * K);
    const int64_t s = vlist[v];
    const int64_t g = max(pidx[s * K + k], 0);
    float* x0 = X0 + row * X0_W;
    // embedding + its encoding: lane = channel
    {
        float e = emb[g * F_EMB + lane];
        x0[lane] = e;
        float* pe = x0 + F_EMB + lane * (2 * NF_FEAT);
#pragma unroll
        for (int f = 0; f < NF_FEAT; ++f) {
            float sn, cs;
            sincosf(e * (float)(1 << f), &sn, &cs);
            pe[2 * f] = sn;
            pe[2 * f + 1] = cs;
        }
    }
    // distance features (agg_dist_pers = 20)
    const float px = xyz[g * 3], py = xyz[g * 3 + 1], pz = xyz[g * 3 + 2];
    float d[6];
    {
        float wx = px - loc_w[s * 3], wy = py - loc_w[s * 3 + 1], wz = pz - loc_w[s * 3 + 2];
        rot3((cam + 12), wx, wy, wz, d[0], d[1], d[2]);
        float qx, qy, qz;
        if (xyz_pers) {
            qx = xyz_pers[g * 3]; qy = xyz_pers[g * 3 + 1]; qz = xyz_pers[g * 3 + 2];
        } else {
            float cx, cy, cz;
            rot3((cam + 3), px - cam[0], py - cam[1], pz - cam[2], cx, cy, cz);
            qx = cx / cz; qy = cy / cz; qz = cz;
        }
        const float sx = loc_pers[s * 3], sy = loc_pers[s * 3 + 1], sz = loc_pers[s * 3 + 2];
        d[3] = qx * qz - sx * sz;
        d[4] = qy * qz - sy * sz;
        d[5] = qz - sz;
    }
    if (lane < 6 * NF_DIST) {
        int c = lane / NF_DIST, f = lane - c * NF_DIST;
        float dv = d[0];
#pragma unroll
        for (int i = 1; i < 6; ++i) dv = (c == i) ? d[i] : dv;
        float sn, cs;
        sincosf(dv * (float)(1 << f), &sn, &cs);
        float* o = x0 + F_EMB + 2 * NF_FEAT * F_EMB + lane * 2;
        o[0] = sn;
        o[1] = cs;
    }
    // block3 extras: colour, dir - view, <dir, view>
    if (lane < E_W) {
        float vx, vy, vz, rx, ry, rz;
        rot3((cam + 12), raydirs[s * 3], raydirs[s * 3 + 1], raydirs[s * 3 + 2], vx, vy, vz);
        rot3((cam + 12), dir[g * 3], dir[g * 3 + 1], dir[g * 3 + 2], rx, ry, rz);
        float out;
        if (lane < 3) out = color[g * 3 + lane];
        else if (lane == 3) out = rx - vx;
        else if (lane == 4) out = ry - vy;
        else if (lane == 5) out = rz - vz;
        else out = rx * vx + ry * vy + rz * vz;
        E[row * E_W + lane] = out;
    }
}

__global__ void __launch_bounds__(256)
nbr_features_bwd_kernel(const float* __restrict__ dX0, const float* __restrict__ dE, const float* __restrict__ emb,
                        const int32_t* __restrict__ pidx, const uint8_t* __restrict__ mask, const int32_t* __restrict__ vlist,
                        const float* __restrict__ raydirs, const float* __restrict__ cam, int64_t rows, int K, float* __restrict__ d_emb,
                        float* __restrict__ d_color, float* __restrict__ d_dir, int ldx) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const int64_t v = row / K;
    const int k = (int)(row - v * K);
    const int64_t s = vlist[v];
    if (slot_masked(pidx, mask, s * K + k)) return;     // masked rows carry exactly zero gradient
    const int64_t g = pidx[s * K + k];
    const float* dx = dX0 + row * ldx;
    if (d_emb) {
        float e = emb[g * F_EMB + lane];
        float acc = dx[lane];
        const float* dpe = dx + F_EMB + lane * (2 * NF_FEAT);
#pragma unroll
        for (int f = 0; f < NF_FEAT; ++f) {
            float fr = (float)(1 << f), sn, cs;
            sincosf(e * fr, &sn, &cs);
            acc += fr * (cs * dpe[2 * f] - sn * dpe[2 * f + 1]);
        }
        atomicAdd(&d_emb[g * F_EMB + lane], acc);
    }
    if (lane < 3) {
        const float* de = dE + row * E_W;
        if (d_color) atomicAdd(&d_color[g * 3 + lane], de[lane]);
        if (d_dir) {
            float vx, vy, vz;
            rot3((cam + 12), raydirs[s * 3], raydirs[s * 3 + 1], raydirs[s * 3 + 2], vx, vy, vz);
            float gx = de[3] + de[6] * vx, gy = de[4] + de[6] * vy, gz = de[5] + de[6] * vz;
            // dirR_j = sum_i dir_i rt[i][j]  ->  d dir_i = sum_j g_j rt[i][j]
            float o = gx * (cam + 12)[lane * 3] + gy * (cam + 12)[lane * 3 + 1] + gz * (cam + 12)[lane * 3 + 2];
            atomicAdd(&d_dir[g * 3 + lane], o);
        }
    }
}

// ------------------------------------------------------------------ A3 epilogue: density head + weighted K-sum
__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// one warp per valid sample; lane owns 8 of the 256 hidden channels
__global__ void __launch_bounds__(256)
alpha_ksum_fwd_kernel(const float* __restrict__ H, const float* __restrict__ weight, const float* __restrict__ confc,
                      const int32_t* __restrict__ vlist, const float* __restrict__ w_alpha, const float* __restrict__ b_alpha,
                      const float* __restrict__ raydirs, const float* __restrict__ cam, int64_t Nv, int K, float* __restrict__ sigma,
                      float* __restrict__ X5, float* __restrict__ alpha_raw) {
    const int lane = threadIdx.x & 31;
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (v >= Nv) return;
    const int64_t s = vlist[v];
    const float4* wa4 = reinterpret_cast<const float4*>(w_alpha) + lane * 2;
    const float4 wa0 = wa4[0], wa1 = wa4[1];
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
    float sg = 0.f;
    for (int k = 0; k < K; ++k) {
        const float4* h4 = reinterpret_cast<const float4*>(H + (v * K + k) * HID) + lane * 2;
        float4 h0 = h4[0], h1 = h4[1];
        float dot = h0.x * wa0.x + h0.y * wa0.y + h0.z * wa0.z + h0.w * wa0.w + h1.x * wa1.x + h1.y * wa1.y + h1.z * wa1.z + h1.w * wa1.w;
        float raw = warp_sum(dot) + b_alpha[0];
        float wc = weight[s * K + k] * (confc ? confc[s * K + k] : 1.f);
        sg += wc * softplus_t(raw - 1.f);
        a0.x += wc * h0.x; a0.y += wc * h0.y; a0.z += wc * h0.z; a0.w += wc * h0.w;
        a1.x += wc * h1.x; a1.y += wc * h1.y; a1.z += wc * h1.z; a1.w += wc * h1.w;
        if (lane == 0 && alpha_raw) alpha_raw[v * K + k] = raw;
    }
    float4* o4 = reinterpret_cast<float4*>(X5 + v * X5_W) + lane * 2;
    o4[0] = a0; o4[1] = a1;
    if (lane == 0) sigma[v] = sg;
    if (lane < 3 * NF_VIEW) {       // viewdir encoding, `ori=True` layout minus the raw 3: all sines then all cosines
        float vx, vy, vz;
        rot3((cam + 12), raydirs[s * 3], raydirs[s * 3 + 1], raydirs[s * 3 + 2], vx, vy, vz);
        int c = lane / NF_VIEW, f = lane - c * NF_VIEW;
        float val = c == 0 ? vx : (c == 1 ? vy : vz);
        float sn, cs;
        sincosf(val * (float)(1 << f), &sn, &cs);
        X5[v * X5_W + HID + lane] = sn;
        X5[v * X5_W + HID + 3 * NF_VIEW + lane] = cs;
    }
}

// grid-stride over samples so the dense-weight gradient is accumulated in registers first
template <int KT>      // KT > 0: compile-time neighbour count (the K-loop unrolls and its 2*K row loads are issued up front)
__global__ void __launch_bounds__(256)
alpha_ksum_bwd_kernel(const float* __restrict__ H, const float* __restrict__ weight, const float* __restrict__ confc,
                      const int32_t* __restrict__ vlist, const float* __restrict__ w_alpha, const float* __restrict__ alpha_raw,
                      const float* __restrict__ d_sigma, const float* __restrict__ dX5, int64_t Nv, int K, float* __restrict__ dH,
                      float* __restrict__ d_wc, float* __restrict__ d_walpha, float* __restrict__ d_balpha) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float4* wa4 = reinterpret_cast<const float4*>(w_alpha) + lane * 2;
    const float4 wa0 = wa4[0], wa1 = wa4[1];
    float4 gw0 = make_float4(0, 0, 0, 0), gw1 = gw0;
    float gb = 0.f;
    for (int64_t v = warp0; v < Nv; v += nwarps) {
        const int64_t s = vlist[v];
        const float ds = d_sigma[v];
        const float4* g4 = reinterpret_cast<const float4*>(dX5 + v * X5_W) + lane * 2;
        const float4 g0 = g4[0], g1 = g4[1];
        const int KK = KT > 0 ? KT : K;
#pragma unroll
        for (int k = 0; k < KK; ++k) {
            const int64_t row = v * KK + k;
            const float4* h4 = reinterpret_cast<const float4*>(H + row * HID) + lane * 2;
            float4 h0 = h4[0], h1 = h4[1];
            float wc = weight[s * KK + k] * (confc ? confc[s * KK + k] : 1.f);
            float raw = alpha_raw[row] - 1.f;
            float sp = softplus_t(raw);
            float sgm = 1.f / (1.f + expf(-raw));
            float draw = wc * ds * sgm;
            float hd = h0.x * g0.x + h0.y * g0.y + h0.z * g0.z + h0.w * g0.w + h1.x * g1.x + h1.y * g1.y + h1.z * g1.z + h1.w * g1.w;
            hd = warp_sum(hd);
            if (lane == 0) d_wc[row] = sp * ds + hd;
            float4 o0, o1;
            o0.x = wc * g0.x + draw * wa0.x; o0.y = wc * g0.y + draw * wa0.y; o0.z = wc * g0.z + draw * wa0.z; o0.w = wc * g0.w + draw * wa0.w;
            o1.x = wc * g1.x + draw * wa1.x; o1.y = wc * g1.y + draw * wa1.y; o1.z = wc * g1.z + draw * wa1.z; o1.w = wc * g1.w + draw * wa1.w;
            float4* d4 = reinterpret_cast<float4*>(dH + row * HID) + lane * 2;
            d4[0] = o0; d4[1] = o1;
            gw0.x += draw * h0.x; gw0.y += draw * h0.y; gw0.z += draw * h0.z; gw0.w += draw * h0.w;
            gw1.x += draw * h1.x; gw1.y += draw * h1.y; gw1.z += draw * h1.z; gw1.w += draw * h1.w;
            gb += draw;
        }
    }
    // CTA-level reduction first: one global atomic per channel per CTA instead of one per warp (the 256 addresses are shared
    // by every warp of the grid)
    __shared__ float red[8][HID + 1];
    {
        float* o = &red[threadIdx.x >> 5][lane * 8];
        o[0] = gw0.x; o[1] = gw0.y; o[2] = gw0.z; o[3] = gw0.w; o[4] = gw1.x; o[5] = gw1.y; o[6] = gw1.z; o[7] = gw1.w;
        if (lane == 0) red[threadIdx.x >> 5][HID] = gb;
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int c = threadIdx.x; c <= HID; c += blockDim.x) {
        float t = 0.f;
        for (int w = 0; w < nw; ++w) t += red[w][c];
        atomicAdd(c < HID ? d_walpha + c : d_balpha, t);
    }
}

// Same backward for the fused training path: the saved layer-3 output is read from its split image (img_common.cuh; lane = one
// 8-column group = one 16-byte piece per plane) and the result is written as the GATED gradient dZ_3 = dH * act'(H) in the
// same format -- the first operand of the fused data-gradient chain (nbr_bwd_f16.cu) and of the weight-gradient kernel.
// (forcing two CTAs per SM with __launch_bounds__(256, 2) spills and is slower: 0.40 -> 0.51 ms)
__global__ void __launch_bounds__(256)
alpha_ksum_bwd_img_kernel(const uint8_t* __restrict__ himg, const float* __restrict__ weight, const float* __restrict__ confc,
                          const int32_t* __restrict__ vlist, const float* __restrict__ w_alpha, const float* __restrict__ alpha_raw,
                          const float* __restrict__ d_sigma, const float* __restrict__ dX5, int64_t Nv, uint8_t* __restrict__ dzimg,
                          float* __restrict__ d_wc, float* __restrict__ d_walpha, float* __restrict__ d_balpha, float* __restrict__ d_confc) {
    constexpr int KK = 8;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float4* wa4 = reinterpret_cast<const float4*>(w_alpha) + lane * 2;
    const float4 wa0 = wa4[0], wa1 = wa4[1];
    const float wa[8] = {wa0.x, wa0.y, wa0.z, wa0.w, wa1.x, wa1.y, wa1.z, wa1.w};
    float gw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) gw[i] = 0.f;
    float gb = 0.f;
    for (int64_t v = warp0; v < Nv; v += nwarps) {
        const int64_t s = vlist[v];
        const float ds = d_sigma[v];
        const float4* g4 = reinterpret_cast<const float4*>(dX5 + v * X5_W) + lane * 2;
        const float4 g0 = g4[0], g1 = g4[1];
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        uint4 hh[KK], hl[KK];
#pragma unroll
        for (int k = 0; k < KK; ++k) {
            const uint8_t* p = himg + img::piece_off(v * KK + k, lane, HID);
            hh[k] = __ldg(reinterpret_cast<const uint4*>(p));
            hl[k] = __ldg(reinterpret_cast<const uint4*>(p + img::plane_bytes(HID)));
        }
#pragma unroll
        for (int k = 0; k < KK; ++k) {
            const int64_t row = v * KK + k;
            float h[8];
            img::join8_bf16(hh[k], hl[k], h);
            const float wc = weight[s * KK + k] * (confc ? confc[s * KK + k] : 1.f);
            const float raw = alpha_raw[row] - 1.f;
            const float sp = softplus_t(raw);
            const float sgm = 1.f / (1.f + expf(-raw));
            const float draw = wc * ds * sgm;
            float hd = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) hd = fmaf(h[i], g[i], hd);
            hd = warp_sum(hd);
            if (lane == 0) {
                d_wc[row] = sp * ds + hd;
                if (d_confc) d_confc[s * KK + k] = (sp * ds + hd) * weight[s * KK + k];      // straight into the (S, K) gradient of conf_coefficient
            }
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                o[i] = (wc * g[i] + draw * wa[i]) * (h[i] > 0.f ? 1.f : 0.01f);
                gw[i] = fmaf(draw, h[i], gw[i]);
            }
            gb += draw;
            uint4 hi, lo;
            img::split8_bf16(o, hi, lo);
            uint8_t* q = dzimg + img::piece_off(row, lane, HID);
            *reinterpret_cast<uint4*>(q) = hi;
            *reinterpret_cast<uint4*>(q + img::plane_bytes(HID)) = lo;
        }
    }
    __shared__ float red[8][HID + 1];
    {
        float* o = &red[threadIdx.x >> 5][lane * 8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = gw[i];
        if (lane == 0) red[threadIdx.x >> 5][HID] = gb;
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int c = threadIdx.x; c <= HID; c += blockDim.x) {
        float t = 0.f;
        for (int w = 0; w < nw; ++w) t += red[w][c];
        atomicAdd(c < HID ? d_walpha + c : d_balpha, t);
    }
}

// gradient of the per-point confidence: through wc = w_hat * clampST(conf) (identity gradient) for
// valid samples, plus the upstream gradient of the returned conf_coefficient for every slot
// (masked slots alias point 0 exactly like the reference's clamp(pidx, 0) gather).
__global__ void conf_bwd_kernel(const float* __restrict__ d_wc, const float* __restrict__ weight, const int32_t* __restrict__ vlist,
                                const int32_t* __restrict__ pidx, int64_t Nv, int K, float* __restrict__ d_conf) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nv * K) return;
    int64_t v = i / K;
    int k = (int)(i - v * K);
    int64_t s = vlist[v];
    float w = weight[s * K + k];
    if (w == 0.f) return;
    atomicAdd(&d_conf[max(pidx[s * K + k], 0)], d_wc[i] * w);
}
__global__ void __launch_bounds__(256) conf_up_bwd_kernel(const float* __restrict__ d_confc, const int32_t* __restrict__ pidx, int64_t n,
                                                          float* __restrict__ d_conf) {
    // masked slots (pidx < 0) all alias point 0 (clamp(pidx, 0), neural_points.py:711): hundreds of thousands of terms for ONE address.
    // They are summed in double per thread (grid-stride) and per block, one atomic per block: neither serialised same-address
    // atomics nor the fp32 rounding of a 300k-term running sum (7e-4 relative at the shipped shapes)
    double gm = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = d_confc[i];
        const int p = pidx[i];
        if (p < 0) gm += (double)g;
        else if (g != 0.f) atomicAdd(&d_conf[p], g);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gm += __shfl_xor_sync(0xffffffffu, gm, o);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gm;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        if (t != 0.0) atomicAdd(&d_conf[0], (float)t);
    }
}

// ------------------------------------------------------------------ P1: projection into the reference views
__global__ void project_views_kernel(const float* __restrict__ loc_w, const float* __restrict__ w2c, const float* __restrict__ Kmat,
                                     const float* __restrict__ campos, const float* __restrict__ campos_n, int V, int64_t S,
                                     float* __restrict__ xy, float* __restrict__ delta) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)V * S) return;
    int v = (int)(i / S);
    int64_t s = i - (int64_t)v * S;
    const float x = loc_w[s * 3], y = loc_w[s * 3 + 1], z = loc_w[s * 3 + 2];
    const float* m = w2c + v * 16;   // row-major world->camera
    float cx = x * m[0] + y * m[1] + z * m[2] + m[3];
    float cy = x * m[4] + y * m[5] + z * m[6] + m[7];
    float cz = x * m[8] + y * m[9] + z * m[10] + m[11];
    float ix = cx * Kmat[0] + cy * Kmat[1] + cz * Kmat[2];
    float iy = cx * Kmat[3] + cy * Kmat[4] + cz * Kmat[5];
    float iz = cx * Kmat[6] + cy * Kmat[7] + cz * Kmat[8];
    float den = iz + 1e-10f;
    xy[i * 2] = ix / den;
    xy[i * 2 + 1] = iy / den;
    if (delta) {
        float ax = x - campos[0], ay = y - campos[1], az = z - campos[2];
        float an = sqrtf(ax * ax + ay * ay + az * az) + 1e-6f;
        float bx = x - campos_n[v * 3], by = y - campos_n[v * 3 + 1], bz = z - campos_n[v * 3 + 2];
        float bn = sqrtf(bx * bx + by * by + bz * bz) + 1e-6f;
        delta[i * 3] = bx / bn - ax / an;
        delta[i * 3 + 1] = by / bn - ay / an;
        delta[i * 3 + 2] = bz / bn - az / an;
    }
}

// ------------------------------------------------------------------ I2: pyramid lookup
struct hnr_pyramid_t {
    const float* lvl[4];   // NHWC: (V,H,W,3), (V,h1,w1,6), (V,h2,w2,12), (V,h3,w3,24)
    float* grad[4];        // matching gradient buffers (grad[0] unused)
    int h[4], w[4], c[4];
};

__device__ __forceinline__ void bilin_setup(int dst, int in_size, int out_size, int& i0, int& i1, float& l0, float& l1) {
    // torch upsample_bilinear2d, align_corners=False
    float scale = (float)in_size / (float)out_size;
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    l1 = src - (float)i0;
    l0 = 1.f - l1;
}

template <bool BWD>
__global__ void __launch_bounds__(256)
image_gather_kernel(hnr_pyramid_t P, const float* __restrict__ xy, const int32_t* __restrict__ vlist, int V, int64_t S, int64_t Nv,
                    float* __restrict__ aux, float* __restrict__ ok, const float* __restrict__ d_aux, int d_ld = AUX_C) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (int64_t)V * Nv) return;
    const int v = (int)(wid / Nv);
    const int64_t n = wid - (int64_t)v * Nv;
    const int64_t s = vlist[n];
    const float fx = xy[((int64_t)v * S + s) * 2], fy = xy[((int64_t)v * S + s) * 2 + 1];
    int px = (int)fx, py = (int)fy;                      // truncation toward zero, as .to(torch.int32)
    const int H = P.h[0], W = P.w[0];
    const bool inb = !(px < 0 || px >= W || py < 0 || py >= H);
    if (!BWD && lane == 0) ok[wid] = inb ? 1.f : 0.f;
    const bool zero = !inb || (px == 0 && py == 0);      // pixel (0,0) is the zeroed "invalid" slot
    for (int ch = lane; ch < AUX_C; ch += 32) {
        int l = ch < 3 ? 0 : (ch < 9 ? 1 : (ch < 21 ? 2 : 3));
        int c = ch - (l == 0 ? 0 : (l == 1 ? 3 : (l == 2 ? 9 : 21)));
        if (!BWD) {
            float val = 0.f;
            if (!zero) {
                if (l == 0) {
                    val = P.lvl[0][(((int64_t)v * H + py) * W + px) * 3 + c];
                } else {
                    int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
                    bilin_setup(py, P.h[l], H, y0, y1, ly0, ly1);
                    bilin_setup(px, P.w[l], W, x0, x1, lx0, lx1);
                    const float* b = P.lvl[l] + (int64_t)v * P.h[l] * P.w[l] * P.c[l];
                    const int C = P.c[l], wl = P.w[l];
                    float v00 = b[((int64_t)y0 * wl + x0) * C + c], v01 = b[((int64_t)y0 * wl + x1) * C + c];
                    float v10 = b[((int64_t)y1 * wl + x0) * C + c], v11 = b[((int64_t)y1 * wl + x1) * C + c];
                    val = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
                }
            }
            aux[wid * AUX_C + ch] = val;
        } else {
            if (zero || l == 0) continue;
            float g = d_aux[wid * d_ld + ch];
            if (g == 0.f) continue;
            int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
            bilin_setup(py, P.h[l], H, y0, y1, ly0, ly1);
            bilin_setup(px, P.w[l], W, x0, x1, lx0, lx1);
            float* b = P.grad[l] + (int64_t)v * P.h[l] * P.w[l] * P.c[l];
            const int C = P.c[l], wl = P.w[l];
            atomicAdd(&b[((int64_t)y0 * wl + x0) * C + c], g * ly0 * lx0);
            atomicAdd(&b[((int64_t)y0 * wl + x1) * C + c], g * ly0 * lx1);
            atomicAdd(&b[((int64_t)y1 * wl + x0) * C + c], g * ly1 * lx0);
            atomicAdd(&b[((int64_t)y1 * wl + x1) * C + c], g * ly1 * lx1);
        }
    }
}

// Forward lookup, vectorised: 16 lanes per (view, sample) pair, two pairs per warp.  A lane owns one 16-byte (level 3, level 2)
// or 8-byte (level 1) channel group of one pyramid level -- the NHWC layout makes the channels of a tap contiguous -- and
// evaluates the 4-tap bilinear interpolation with the same expression per channel as the scalar kernel above:
//   sub-lane 0..5 : level 3 (24 ch) float4 q        -> aux[21 + 4q ..]
//   sub-lane 6..8 : level 2 (12 ch) float4 q-6      -> aux[ 9 + 4(q-6) ..]
//   sub-lane 9..11: level 1 ( 6 ch) float2 q-9      -> aux[ 3 + 2(q-9) ..]
//   sub-lane 12   : level 0 rgb (nearest pixel), validity flag
__global__ void __launch_bounds__(256)
image_gather_fwd_v2_kernel(hnr_pyramid_t P, const float* __restrict__ xy, const int32_t* __restrict__ vlist, int V, int64_t S, int64_t Nv,
                           float* __restrict__ aux, float* __restrict__ ok, int aux_ld, const float* __restrict__ delta) {
    const int q = threadIdx.x & 15;
    const int64_t pair = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (pair >= (int64_t)V * Nv || q > 12) return;
    const int v = (int)(pair / Nv);
    const int64_t n = pair - (int64_t)v * Nv;
    const int64_t s = vlist[n];
    const float2 f = *reinterpret_cast<const float2*>(xy + ((int64_t)v * S + s) * 2);
    const int px = (int)f.x, py = (int)f.y;                  // truncation toward zero, as .to(torch.int32)
    const int H = P.h[0], W = P.w[0];
    const bool inb = !(px < 0 || px >= W || py < 0 || py >= H);
    const bool zero = !inb || (px == 0 && py == 0);          // pixel (0,0) is the zeroed "invalid" slot
    float* o = aux + pair * aux_ld;
    if (q == 12) {
        ok[pair] = inb ? 1.f : 0.f;
        if (delta) {                                         // 48-wide rows: the view-direction difference rides in columns 45..47
            const float* dp = delta + ((int64_t)v * S + s) * 3;
            o[AUX_C] = dp[0]; o[AUX_C + 1] = dp[1]; o[AUX_C + 2] = dp[2];
        }
        float r = 0.f, g = 0.f, b = 0.f;
        if (!zero) {
            const float* p0 = P.lvl[0] + (((int64_t)v * H + py) * W + px) * 3;
            r = p0[0]; g = p0[1]; b = p0[2];
        }
        o[0] = r; o[1] = g; o[2] = b;
        return;
    }
    const int l = q < 6 ? 3 : (q < 9 ? 2 : 1);
    const int grp = q < 6 ? q : (q < 9 ? q - 6 : q - 9);     // float4 (levels 3, 2) or float2 (level 1) index inside the pixel
    const int ch0 = l == 3 ? 21 + 4 * grp : (l == 2 ? 9 + 4 * grp : 3 + 2 * grp);
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
    if (!zero) {
        int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
        const int hl = l == 3 ? P.h[3] : (l == 2 ? P.h[2] : P.h[1]), wl = l == 3 ? P.w[3] : (l == 2 ? P.w[2] : P.w[1]);
        const int C = l == 3 ? 24 : (l == 2 ? 12 : 6);
        const float* lv = l == 3 ? P.lvl[3] : (l == 2 ? P.lvl[2] : P.lvl[1]);
        bilin_setup(py, hl, H, y0, y1, ly0, ly1);
        bilin_setup(px, wl, W, x0, x1, lx0, lx1);
        const float* b = lv + (int64_t)v * hl * wl * C;
        const int o00 = (y0 * wl + x0) * C, o01 = (y0 * wl + x1) * C, o10 = (y1 * wl + x0) * C, o11 = (y1 * wl + x1) * C;
        if (l == 1) {
            const float2 a = *reinterpret_cast<const float2*>(b + o00 + 2 * grp), bq = *reinterpret_cast<const float2*>(b + o01 + 2 * grp);
            const float2 c = *reinterpret_cast<const float2*>(b + o10 + 2 * grp), d = *reinterpret_cast<const float2*>(b + o11 + 2 * grp);
            r0 = ly0 * (lx0 * a.x + lx1 * bq.x) + ly1 * (lx0 * c.x + lx1 * d.x);
            r1 = ly0 * (lx0 * a.y + lx1 * bq.y) + ly1 * (lx0 * c.y + lx1 * d.y);
        } else {
            const float4 a = *reinterpret_cast<const float4*>(b + o00 + 4 * grp), bq = *reinterpret_cast<const float4*>(b + o01 + 4 * grp);
            const float4 c = *reinterpret_cast<const float4*>(b + o10 + 4 * grp), d = *reinterpret_cast<const float4*>(b + o11 + 4 * grp);
            r0 = ly0 * (lx0 * a.x + lx1 * bq.x) + ly1 * (lx0 * c.x + lx1 * d.x);
            r1 = ly0 * (lx0 * a.y + lx1 * bq.y) + ly1 * (lx0 * c.y + lx1 * d.y);
            r2 = ly0 * (lx0 * a.z + lx1 * bq.z) + ly1 * (lx0 * c.z + lx1 * d.z);
            r3 = ly0 * (lx0 * a.w + lx1 * bq.w) + ly1 * (lx0 * c.w + lx1 * d.w);
        }
    }
    o[ch0] = r0; o[ch0 + 1] = r1;
    if (l != 1) { o[ch0 + 2] = r2; o[ch0 + 3] = r3; }
}

// Backward of the lookup with the forward's lane mapping: 16 sub-lanes per (view, sample), sub-lane q < 12 owns one float4 (levels 3, 2)
// or float2 (level 1) channel group of one pyramid level and scatters its gradient to the 4 taps with VECTOR reductions
// (red.global.add.v4.f32 / .v2.f32, sm_90+): 48 reduction instructions per (view, sample) instead of 168 scalar atomics -- the
// kernel is bound by the L2 atomic units.  Level 0 (the rgb image) receives no gradient.
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__global__ void __launch_bounds__(256)
image_gather_bwd_v2_kernel(hnr_pyramid_t P, const float* __restrict__ xy, const int32_t* __restrict__ vlist, int V, int64_t S, int64_t Nv,
                           const float* __restrict__ d_aux, int d_ld) {
    const int q = threadIdx.x & 15;
    const int64_t pair = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (pair >= (int64_t)V * Nv || q > 11) return;
    const int v = (int)(pair / Nv);
    const int64_t n = pair - (int64_t)v * Nv;
    const int64_t s = vlist[n];
    const float2 f = *reinterpret_cast<const float2*>(xy + ((int64_t)v * S + s) * 2);
    const int px = (int)f.x, py = (int)f.y;
    const int H = P.h[0], W = P.w[0];
    const bool inb = !(px < 0 || px >= W || py < 0 || py >= H);
    if (!inb || (px == 0 && py == 0)) return;                // zeroed slot: no gradient
    const int l = q < 6 ? 3 : (q < 9 ? 2 : 1);
    const int grp = q < 6 ? q : (q < 9 ? q - 6 : q - 9);
    const int ch0 = l == 3 ? 21 + 4 * grp : (l == 2 ? 9 + 4 * grp : 3 + 2 * grp);
    const float* gp = d_aux + pair * d_ld + ch0;
    const float g0 = gp[0], g1 = gp[1], g2 = l != 1 ? gp[2] : 0.f, g3 = l != 1 ? gp[3] : 0.f;
    if (g0 == 0.f && g1 == 0.f && g2 == 0.f && g3 == 0.f) return;
    int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
    const int hl = l == 3 ? P.h[3] : (l == 2 ? P.h[2] : P.h[1]), wl = l == 3 ? P.w[3] : (l == 2 ? P.w[2] : P.w[1]);
    const int C = l == 3 ? 24 : (l == 2 ? 12 : 6);
    float* lv = l == 3 ? P.grad[3] : (l == 2 ? P.grad[2] : P.grad[1]);
    bilin_setup(py, hl, H, y0, y1, ly0, ly1);
    bilin_setup(px, wl, W, x0, x1, lx0, lx1);
    float* b = lv + (int64_t)v * hl * wl * C + (l == 1 ? 2 * grp : 4 * grp);
    const int o00 = (y0 * wl + x0) * C, o01 = (y0 * wl + x1) * C, o10 = (y1 * wl + x0) * C, o11 = (y1 * wl + x1) * C;
    const float w00 = ly0 * lx0, w01 = ly0 * lx1, w10 = ly1 * lx0, w11 = ly1 * lx1;
    if (l == 1) {
        red_add_v2(b + o00, g0 * w00, g1 * w00); red_add_v2(b + o01, g0 * w01, g1 * w01);
        red_add_v2(b + o10, g0 * w10, g1 * w10); red_add_v2(b + o11, g0 * w11, g1 * w11);
    } else {
        red_add_v4(b + o00, g0 * w00, g1 * w00, g2 * w00, g3 * w00); red_add_v4(b + o01, g0 * w01, g1 * w01, g2 * w01, g3 * w01);
        red_add_v4(b + o10, g0 * w10, g1 * w10, g2 * w10, g3 * w10); red_add_v4(b + o11, g0 * w11, g1 * w11, g2 * w11, g3 * w11);
    }
}

// ------------------------------------------------------------------ I3: learned multi-view blend
// thread per (sample, channel): merged = keep * sum_v aux_v w_v / (sum_v w_v + 1e-6), w_v = sig_v * ok_v
__global__ void blend_fwd_kernel(const float* __restrict__ aux, const float* __restrict__ sig, const float* __restrict__ ok,
                                 const uint8_t* __restrict__ keep, int V, int64_t Nv, int aux_ld, float* __restrict__ merged, int merged_ld) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nv * merged_ld) return;
    int64_t n = i / merged_ld;
    int c = (int)(i - n * merged_ld);
    if (c >= AUX_C) { merged[i] = 0.f; return; }             // padding columns of a 48-wide row
    float num = 0.f, den = 0.f;
    for (int v = 0; v < V; ++v) {
        float w = sig[(int64_t)v * Nv + n] * ok[(int64_t)v * Nv + n];
        num += aux[((int64_t)v * Nv + n) * aux_ld + c] * w;
        den += w;
    }
    float m = num / (den + 1e-6f);
    if (keep && !keep[n]) m = 0.f;
    merged[i] = m;
}

// warp per sample: d_aux, d_sig
__global__ void __launch_bounds__(256)
blend_bwd_kernel(const float* __restrict__ aux, const float* __restrict__ sig, const float* __restrict__ ok,
                 const uint8_t* __restrict__ keep, const float* __restrict__ d_merged, int V, int64_t Nv, float* __restrict__ d_aux,
                 float* __restrict__ d_sig, int ald, int dld) {
    const int lane = threadIdx.x & 31;
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= Nv) return;
    const bool kept = !(keep && !keep[n]);
    float den = 1e-6f;
    for (int v = 0; v < V; ++v) den += sig[(int64_t)v * Nv + n] * ok[(int64_t)v * Nv + n];
    // merged_raw per channel (two channels per lane: lane, lane+32)
    float dm[2], mr[2];
    for (int j = 0; j < 2; ++j) {
        int c = lane + 32 * j;
        dm[j] = (c < AUX_C && kept) ? d_merged[n * AUX_C + c] : 0.f;
        float num = 0.f;
        if (c < AUX_C)
            for (int v = 0; v < V; ++v) num += aux[((int64_t)v * Nv + n) * ald + c] * sig[(int64_t)v * Nv + n] * ok[(int64_t)v * Nv + n];
        mr[j] = num / den;
    }
    for (int v = 0; v < V; ++v) {
        const int64_t r = (int64_t)v * Nv + n;
        const float w = sig[r] * ok[r];
        float part = 0.f;
        for (int j = 0; j < 2; ++j) {
            int c = lane + 32 * j;
            if (c < AUX_C) {
                float a = aux[r * ald + c];
                d_aux[r * dld + c] = dm[j] * w / den;
                part += dm[j] * (a - mr[j]);
            } else if (c < dld) {
                d_aux[r * dld + c] = 0.f;                  // padding columns of 48-wide rows
            }
        }
        part = warp_sum(part);
        if (lane == 0) d_sig[r] = part / den * ok[r];
    }
}

inline unsigned warp_blocks(int64_t warps, int threads = 256) { return (unsigned)hnr_cdiv(warps * 32, threads); }

}  // namespace

extern "C" int hnr_nbr_weights(const float* xyz, const float* conf, const int32_t* pidx, const uint8_t* mask, const float* loc_w,
                               int64_t S, int64_t K, float* weight, float* confc, uint8_t* valid, void* stream) {
    HNR_CHECK_ARG(S >= 0 && K > 0, "nbr_weights: bad shape");
    if (S == 0) return HNR_OK;
    nbr_weights_kernel<<<(unsigned)hnr_cdiv(S, 256), 256, 0, (cudaStream_t)stream>>>(xyz, conf, pidx, mask, loc_w, S, (int)K, weight,
                                                                                    confc, valid);
    HNR_CHECK_LAUNCH("nbr_weights");
    return HNR_OK;
}

extern "C" int hnr_nbr_features(const float* xyz, const float* xyz_pers, const float* emb, const float* color, const float* dir,
                                const int32_t* pidx, const int32_t* vlist, const float* loc_w, const float* loc_pers,
                                const float* raydirs, const float* cam, int64_t Nv, int64_t K, int64_t emb_dim, float* X0,
                                float* E, void* stream) {
    HNR_CHECK_ARG(emb_dim == F_EMB, "nbr_features: point_features_dim must be 32");
    HNR_CHECK_ARG(Nv >= 0 && K > 0, "nbr_features: bad shape");
    if (Nv == 0) return HNR_OK;
    nbr_features_kernel<<<warp_blocks(Nv * K), 256, 0, (cudaStream_t)stream>>>(xyz, xyz_pers, emb, color, dir, pidx, vlist, loc_w,
                                                                              loc_pers, raydirs, cam, Nv * K, (int)K, X0, E);
    HNR_CHECK_LAUNCH("nbr_features");
    return HNR_OK;
}

extern "C" int hnr_nbr_features_bwd(const float* dX0, const float* dE, const float* emb, const int32_t* pidx, const uint8_t* mask,
                                    const int32_t* vlist, const float* raydirs, const float* cam, int64_t Nv, int64_t K,
                                    float* d_emb, float* d_color, float* d_dir, void* stream) {
    HNR_CHECK_ARG(Nv >= 0 && K > 0, "nbr_features_bwd: bad shape");
    if (Nv == 0) return HNR_OK;
    nbr_features_bwd_kernel<<<warp_blocks(Nv * K), 256, 0, (cudaStream_t)stream>>>(dX0, dE, emb, pidx, mask, vlist, raydirs, cam,
                                                                                  Nv * K, (int)K, d_emb, d_color, d_dir, X0_W);
    HNR_CHECK_LAUNCH("nbr_features_bwd");
    return HNR_OK;
}

// same with an explicit row stride of dX0 (>= 224: only [emb 32 | PE(emb) 192] are read) -- the fused data-gradient chain
// (hnr_nbr_bwd_f16) produces exactly those 224 columns
extern "C" int hnr_nbr_features_bwd_ld(const float* dX0, int64_t ldx, const float* dE, const float* emb, const int32_t* pidx,
                                       const uint8_t* mask, const int32_t* vlist, const float* raydirs, const float* cam, int64_t Nv,
                                       int64_t K, float* d_emb, float* d_color, float* d_dir, void* stream) {
    HNR_CHECK_ARG(Nv >= 0 && K > 0 && ldx >= F_EMB + 2 * NF_FEAT * F_EMB, "nbr_features_bwd_ld: bad shape");
    if (Nv == 0) return HNR_OK;
    nbr_features_bwd_kernel<<<warp_blocks(Nv * K), 256, 0, (cudaStream_t)stream>>>(dX0, dE, emb, pidx, mask, vlist, raydirs, cam,
                                                                                  Nv * K, (int)K, d_emb, d_color, d_dir, (int)ldx);
    HNR_CHECK_LAUNCH("nbr_features_bwd_ld");
    return HNR_OK;
}

extern "C" int hnr_alpha_ksum_fwd(const float* H, const float* weight, const float* confc, const int32_t* vlist, const float* w_alpha,
                                  const float* b_alpha, const float* raydirs, const float* cam, int64_t Nv, int64_t K,
                                  int64_t hidden, float* sigma, float* X5, float* alpha_raw, void* stream) {
    HNR_CHECK_ARG(hidden == HID, "alpha_ksum: shading_feature_num must be 256");
    if (Nv == 0) return HNR_OK;
    alpha_ksum_fwd_kernel<<<warp_blocks(Nv), 256, 0, (cudaStream_t)stream>>>(H, weight, confc, vlist, w_alpha, b_alpha, raydirs, cam,
                                                                            Nv, (int)K, sigma, X5, alpha_raw);
    HNR_CHECK_LAUNCH("alpha_ksum_fwd");
    return HNR_OK;
}

extern "C" int hnr_alpha_ksum_bwd(const float* H, const float* weight, const float* confc, const int32_t* vlist, const float* w_alpha,
                                  const float* alpha_raw, const float* d_sigma, const float* dX5, int64_t Nv, int64_t K,
                                  int64_t hidden, float* dH, float* d_wc, float* d_walpha, float* d_balpha, void* stream) {
    HNR_CHECK_ARG(hidden == HID, "alpha_ksum: shading_feature_num must be 256");
    if (Nv == 0) return HNR_OK;
    int64_t blocks = hnr_cdiv(Nv * 32, 256);
    if (blocks > 8 * HNR_NUM_SMS) blocks = 8 * HNR_NUM_SMS;
    if (K == 8)
        alpha_ksum_bwd_kernel<8><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(H, weight, confc, vlist, w_alpha, alpha_raw, d_sigma, dX5,
                                                                                    Nv, (int)K, dH, d_wc, d_walpha, d_balpha);
    else
        alpha_ksum_bwd_kernel<0><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(H, weight, confc, vlist, w_alpha, alpha_raw, d_sigma, dX5,
                                                                                    Nv, (int)K, dH, d_wc, d_walpha, d_balpha);
    HNR_CHECK_LAUNCH("alpha_ksum_bwd");
    return HNR_OK;
}

// Backward of the density head + weighted K-sum for the fused training path (K == 8): like hnr_alpha_ksum_bwd, but the saved
// layer-3 output comes as a split image (h3img) and the result is the gated gradient dZ_3 = dH * LeakyReLU'(H_3) as a split
// image (dz3img; rows beyond Nv*8 are left to the consumer, hnr_nbr_bwd_f16 zeroes them).
extern "C" int hnr_alpha_ksum_bwd_img(const void* h3img, const float* weight, const float* confc, const int32_t* vlist, const float* w_alpha,
                                      const float* alpha_raw, const float* d_sigma, const float* dX5, int64_t Nv, int64_t K, void* dz3img,
                                      float* d_wc, float* d_walpha, float* d_balpha, float* d_confc, void* stream) {
    HNR_CHECK_ARG(K == 8, "alpha_ksum_bwd_img: K must be 8");
    if (Nv == 0) return HNR_OK;
    int64_t blocks = hnr_cdiv(Nv, 8);
    if (blocks > 4 * HNR_NUM_SMS) blocks = 4 * HNR_NUM_SMS;
    alpha_ksum_bwd_img_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)h3img, weight, confc, vlist, w_alpha, alpha_raw,
                                                                                 d_sigma, dX5, Nv, (uint8_t*)dz3img, d_wc, d_walpha, d_balpha, d_confc);
    HNR_CHECK_LAUNCH("alpha_ksum_bwd_img");
    return HNR_OK;
}

extern "C" int hnr_conf_bwd(const float* d_wc, const float* weight, const int32_t* vlist, const int32_t* pidx, const float* d_confc,
                            int64_t Nv, int64_t S, int64_t K, float* d_conf, void* stream) {
    if (Nv > 0 && d_wc) {
        conf_bwd_kernel<<<(unsigned)hnr_cdiv(Nv * K, 256), 256, 0, (cudaStream_t)stream>>>(d_wc, weight, vlist, pidx, Nv, (int)K, d_conf);
        HNR_CHECK_LAUNCH("conf_bwd");
    }
    if (S > 0 && d_confc) {
        int64_t nb = hnr_cdiv(S * K, 256 * 8);
        if (nb > 2 * HNR_NUM_SMS) nb = 2 * HNR_NUM_SMS;
        conf_up_bwd_kernel<<<(unsigned)(nb < 1 ? 1 : nb), 256, 0, (cudaStream_t)stream>>>(d_confc, pidx, S * K, d_conf);
        HNR_CHECK_LAUNCH("conf_up_bwd");
    }
    return HNR_OK;
}

extern "C" int hnr_project_views(const float* loc_w, const float* w2c, const float* Kmat, const float* campos, const float* campos_n,
                                 int64_t V, int64_t S, float* xy, float* delta, void* stream) {
    if (V * S == 0) return HNR_OK;
    project_views_kernel<<<(unsigned)hnr_cdiv(V * S, 256), 256, 0, (cudaStream_t)stream>>>(loc_w, w2c, Kmat, campos, campos_n, (int)V, S,
                                                                                          xy, delta);
    HNR_CHECK_LAUNCH("project_views");
    return HNR_OK;
}

static int fill_pyramid(hnr_pyramid_t& P, const float* const* lvl, float* const* grad, const int64_t* hw) {
    const int ch[4] = {3, 6, 12, 24};
    for (int l = 0; l < 4; ++l) {
        P.lvl[l] = lvl ? lvl[l] : nullptr;
        P.grad[l] = grad ? grad[l] : nullptr;
        P.h[l] = (int)hw[2 * l];
        P.w[l] = (int)hw[2 * l + 1];
        P.c[l] = ch[l];
    }
    return 0;
}

extern "C" int hnr_image_gather_fwd(const float* const* levels, const int64_t* level_hw, const float* xy, const int32_t* vlist,
                                    int64_t V, int64_t S, int64_t Nv, float* aux, float* ok, int64_t aux_ld, const float* delta,
                                    void* stream) {
    if (V * Nv == 0) return HNR_OK;
    HNR_CHECK_ARG(aux_ld >= AUX_C && (!delta || aux_ld >= AUX_C + 3), "image_gather_fwd: aux_ld must be >= 45 (>= 48 with delta)");
    hnr_pyramid_t P;
    fill_pyramid(P, levels, nullptr, level_hw);
    // vector path: needs the channel groups 16-byte (levels 2, 3) / 8-byte (level 1) aligned, i.e. base pointers from the allocator
    // and the standard 6 / 12 / 24 channel pyramid, and 32-bit tap offsets inside one view
    const bool vec_ok = P.c[1] == 6 && P.c[2] == 12 && P.c[3] == 24 && (reinterpret_cast<uintptr_t>(levels[1]) & 7) == 0 &&
                        (reinterpret_cast<uintptr_t>(levels[2]) & 15) == 0 && (reinterpret_cast<uintptr_t>(levels[3]) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(xy) & 7) == 0 && (int64_t)P.h[1] * P.w[1] * 24 < (1ll << 31);
    if (vec_ok)
        image_gather_fwd_v2_kernel<<<(unsigned)hnr_cdiv(V * Nv * 16, 256), 256, 0, (cudaStream_t)stream>>>(P, xy, vlist, (int)V, S, Nv, aux, ok,
                                                                                                           (int)aux_ld, delta);
    else {
        HNR_CHECK_ARG(aux_ld == AUX_C && !delta, "image_gather_fwd: padded rows need the vector path (standard 6/12/24-channel pyramid)");
        image_gather_kernel<false><<<warp_blocks(V * Nv), 256, 0, (cudaStream_t)stream>>>(P, xy, vlist, (int)V, S, Nv, aux, ok, nullptr);
    }
    HNR_CHECK_LAUNCH("image_gather_fwd");
    return HNR_OK;
}

extern "C" int hnr_image_gather_bwd_ld(float* const* level_grads, const int64_t* level_hw, const float* xy, const int32_t* vlist,
                                       const float* d_aux, int64_t d_ld, int64_t V, int64_t S, int64_t Nv, void* stream) {
    if (V * Nv == 0) return HNR_OK;
    HNR_CHECK_ARG(d_ld >= AUX_C, "image_gather_bwd: d_aux row stride must be >= 45");
    hnr_pyramid_t P;
    fill_pyramid(P, nullptr, level_grads, level_hw);
    // vector reductions need the standard 6 / 12 / 24-channel pyramid with 8- / 16-byte aligned gradient tensors
    const bool vec_ok = P.c[1] == 6 && P.c[2] == 12 && P.c[3] == 24 && (reinterpret_cast<uintptr_t>(level_grads[1]) & 7) == 0 &&
                        (reinterpret_cast<uintptr_t>(level_grads[2]) & 15) == 0 && (reinterpret_cast<uintptr_t>(level_grads[3]) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(xy) & 7) == 0 && (int64_t)P.h[1] * P.w[1] * 24 < (1ll << 31);
    if (vec_ok)
        image_gather_bwd_v2_kernel<<<(unsigned)hnr_cdiv(V * Nv * 16, 256), 256, 0, (cudaStream_t)stream>>>(P, xy, vlist, (int)V, S, Nv, d_aux, (int)d_ld);
    else
        image_gather_kernel<true><<<warp_blocks(V * Nv), 256, 0, (cudaStream_t)stream>>>(P, xy, vlist, (int)V, S, Nv, nullptr, nullptr, d_aux, (int)d_ld);
    HNR_CHECK_LAUNCH("image_gather_bwd");
    return HNR_OK;
}

extern "C" int hnr_image_gather_bwd(float* const* level_grads, const int64_t* level_hw, const float* xy, const int32_t* vlist,
                                    const float* d_aux, int64_t V, int64_t S, int64_t Nv, void* stream) {
    return hnr_image_gather_bwd_ld(level_grads, level_hw, xy, vlist, d_aux, AUX_C, V, S, Nv, stream);
}

extern "C" int hnr_blend_fwd(const float* aux, const float* sig, const float* ok, const uint8_t* keep, int64_t V, int64_t Nv,
                             int64_t aux_ld, float* merged, int64_t merged_ld, void* stream) {
    if (Nv == 0) return HNR_OK;
    HNR_CHECK_ARG(aux_ld >= AUX_C && merged_ld >= AUX_C, "blend_fwd: row strides must be >= 45");
    blend_fwd_kernel<<<(unsigned)hnr_cdiv(Nv * merged_ld, 256), 256, 0, (cudaStream_t)stream>>>(aux, sig, ok, keep, (int)V, Nv, (int)aux_ld, merged,
                                                                                             (int)merged_ld);
    HNR_CHECK_LAUNCH("blend_fwd");
    return HNR_OK;
}

extern "C" int hnr_blend_bwd_ld(const float* aux, int64_t aux_ld, const float* sig, const float* ok, const uint8_t* keep, const float* d_merged,
                                int64_t V, int64_t Nv, float* d_aux, int64_t d_aux_ld, float* d_sig, void* stream) {
    if (Nv == 0) return HNR_OK;
    HNR_CHECK_ARG(aux_ld >= AUX_C && d_aux_ld >= AUX_C && d_aux_ld <= 64, "blend_bwd: row strides must be in [45, 64]");
    blend_bwd_kernel<<<warp_blocks(Nv), 256, 0, (cudaStream_t)stream>>>(aux, sig, ok, keep, d_merged, (int)V, Nv, d_aux, d_sig, (int)aux_ld,
                                                                        (int)d_aux_ld);
    HNR_CHECK_LAUNCH("blend_bwd");
    return HNR_OK;
}

extern "C" int hnr_blend_bwd(const float* aux, const float* sig, const float* ok, const uint8_t* keep, const float* d_merged, int64_t V,
                             int64_t Nv, float* d_aux, float* d_sig, void* stream) {
    return hnr_blend_bwd_ld(aux, AUX_C, sig, ok, keep, d_merged, V, Nv, d_aux, AUX_C, d_sig, stream);
}
