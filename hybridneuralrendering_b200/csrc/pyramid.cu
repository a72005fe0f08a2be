// Feature pyramid of the hybrid image branch: three blocks of [conv3x3 stride 2, LeakyReLU, conv3x3 stride 1,
// LeakyReLU] with 3 -> 6 -> 6, 6 -> 12 -> 12, 12 -> 24 -> 24 channels over the V reference views
// (reference: layers built at models/aggregators/point_aggregators.py:598-630, applied at :1047-1063).  Exact fp32, NHWC
// end to end (the views arrive NHWC and the lookup kernel reads NHWC levels), forward and backward, no cuDNN:
//   * conv_fwd: thread = output pixel, all output channels in registers, weights in shared memory as [tap][ci][co] (broadcast
//     float4 reads), 9 taps x Cin contiguous input floats per pixel; bias + LeakyReLU fused; only the activated outputs are kept
//     (LeakyReLU' is read off their sign in the backward);
//   * conv_bwd_data: thread = input pixel; gathers the gated output gradients of the (stride-dependent) taps that touch it and adds an
//     optional external gradient (the pyramid level is also read by the image gather);
//   * conv_bwd_weight: the reduction over all pixels.  A block stages a strip of gated output-gradient rows and the matching input
//     halo in shared memory; every thread owns a few (tap, ci, co) weights and loops over the strip's pixels; one atomic per weight
//     and block at the end.  Bias gradient = column sums of the same gated gradients.
// These layers hold 1.8 GFLOP per step against ~0.2 GB of traffic: the point of owning them is the ~1.3 ms and ~40 library launches
// (layout conversions, scalePackedTensor, TF32-off cuDNN engines) they cost per training step through cuDNN.
#include "common.cuh"
#include "hnr.h"

namespace {

__device__ __forceinline__ float lrelu_g(float y) { return y > 0.f ? 1.f : 0.01f; }

// ---------------------------------------------------------------------------------------------------------------- forward
template <int CIN, int COUT, int STRIDE>
__global__ void __launch_bounds__(128) conv_fwd_kernel(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ b,
                                                       float* __restrict__ out, int V, int Hi, int Wi, int Ho, int Wo) {
    __shared__ __align__(16) float ws[9 * CIN * COUT];          // [tap][ci][co]
    __shared__ float bs[COUT];
    for (int i = threadIdx.x; i < 9 * CIN * COUT; i += blockDim.x) {
        const int co = i % COUT, ci = (i / COUT) % CIN, tap = i / (COUT * CIN);
        ws[i] = W[(co * CIN + ci) * 9 + tap];                    // torch layout (co, ci, ky, kx)
    }
    if (threadIdx.x < COUT) bs[threadIdx.x] = b[threadIdx.x];
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (int64_t)V * Ho * Wo) return;
    const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho), v = (int)(p / ((int64_t)Wo * Ho));
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = bs[co];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * STRIDE - 1 + ky;
        if (iy < 0 || iy >= Hi) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * STRIDE - 1 + kx;
            if (ix < 0 || ix >= Wi) continue;
            const float* ip = in + (((int64_t)v * Hi + iy) * Wi + ix) * CIN;
            const float* wp = ws + (ky * 3 + kx) * CIN * COUT;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                const float x = __ldg(ip + ci);
#pragma unroll
                for (int co = 0; co < COUT; ++co) acc[co] = fmaf(x, wp[ci * COUT + co], acc[co]);
            }
        }
    }
    float* op = out + p * COUT;
#pragma unroll
    for (int co = 0; co < COUT; ++co) op[co] = acc[co] > 0.f ? acc[co] : 0.01f * acc[co];
}

// ---------------------------------------------------------------------------------------------------------------- data gradient
// dIn[q][ci] = sum over taps / co of dOut[p][co] * act'(Out[p][co]) * W[co][ci][tap]  (+ ext[q][ci]),  q = p * STRIDE - 1 + k
template <int CIN, int COUT, int STRIDE>
__global__ void __launch_bounds__(128) conv_bwd_data_kernel(const float* __restrict__ dOut, const float* __restrict__ Out, const float* __restrict__ W,
                                                            const float* __restrict__ ext, float* __restrict__ dIn, int V, int Hi, int Wi, int Ho, int Wo) {
    __shared__ __align__(16) float ws[9 * COUT * CIN];          // [tap][co][ci]
    for (int i = threadIdx.x; i < 9 * CIN * COUT; i += blockDim.x) {
        const int ci = i % CIN, co = (i / CIN) % COUT, tap = i / (COUT * CIN);
        ws[i] = W[(co * CIN + ci) * 9 + tap];
    }
    __syncthreads();
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (int64_t)V * Hi * Wi) return;
    const int ix = (int)(q % Wi), iy = (int)((q / Wi) % Hi), v = (int)(q / ((int64_t)Wi * Hi));
    float acc[CIN];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) acc[ci] = ext ? ext[q * CIN + ci] : 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int ty = iy + 1 - ky;
        if (ty < 0 || ty % STRIDE != 0) continue;
        const int oy = ty / STRIDE;
        if (oy >= Ho) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int tx = ix + 1 - kx;
            if (tx < 0 || tx % STRIDE != 0) continue;
            const int ox = tx / STRIDE;
            if (ox >= Wo) continue;
            const int64_t p = ((int64_t)v * Ho + oy) * Wo + ox;
            const float* wp = ws + (ky * 3 + kx) * COUT * CIN;
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                const float g = __ldg(dOut + p * COUT + co) * lrelu_g(__ldg(Out + p * COUT + co));
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) acc[ci] = fmaf(g, wp[co * CIN + ci], acc[ci]);
            }
        }
    }
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) dIn[q * CIN + ci] = acc[ci];
}

// ---------------------------------------------------------------------------------------------------------------- weight gradient
// dW[co][ci][ky][kx] = sum over pixels of g[p][co] * in[p * STRIDE - 1 + k][ci],  g = dOut * act'(Out);  db[co] = sum g[p][co].
// A block walks `rb` output rows of one 64-pixel strip.  Per row it stages the gated gradients g[64][COUT] and the three input rows
// of the halo in shared memory (channels zero-padded to multiples of 4).  A thread owns a 4 (co) x 4 (ci) x 3 (kx) register tile of
// one kernel row ky -- per pixel one float4 of g and three float4 of the inputs feed 48 FMAs (1 shared-memory read per 12 FMAs;
// the first version read two operands per FMA and serialised ~2400 same-address global atomics per weight).  The 256 threads form
// NG = 256 / TPG pixel groups (TPG = tiles per kernel row x 3); groups are folded in shared memory, then one global atomic per
// weight and block.
template <int CIN, int COUT, int STRIDE>
__global__ void __launch_bounds__(256) conv_bwd_weight_kernel(const float* __restrict__ dOut, const float* __restrict__ Out, const float* __restrict__ In,
                                                              float* __restrict__ dW, float* __restrict__ db, int V, int Hi, int Wi, int Ho, int Wo, int rb) {
    constexpr int TW = 64;
    constexpr int IW = TW * STRIDE + 2;
    constexpr int CP = (CIN + 3) / 4 * 4, OP = (COUT + 3) / 4 * 4, NCI = CP / 4, NCO = OP / 4;
    constexpr int TPG = NCO * NCI * 3;
    constexpr int NG = 256 / TPG;
    constexpr int NW = 9 * CIN * COUT;
    static_assert(NG >= 1, "tile count");
    __shared__ __align__(16) float g[TW][OP];
    __shared__ __align__(16) float xin[3][IW][CP];
    __shared__ float red[NW + COUT];
    const int strips = (Wo + TW - 1) / TW, rgroups = (Ho + rb - 1) / rb;
    const int sx = blockIdx.x % strips, rg = (blockIdx.x / strips) % rgroups, v = blockIdx.x / (strips * rgroups);
    const int ox0 = sx * TW, ix0 = ox0 * STRIDE - 1;
    const int t = threadIdx.x, grp = t / TPG, u = t - grp * TPG;
    const bool worker = grp < NG;
    const int ky = u / (NCO * NCI), co4 = (u / NCI) % NCO, ci4 = u % NCI;
    float acc[3][4][4];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int d = 0; d < 4; ++d) acc[a][c][d] = 0.f;
    float bsum = 0.f;
    for (int i = t; i < NW + COUT; i += 256) red[i] = 0.f;
    const int oy_end = min(Ho, rg * rb + rb);
    for (int oy = rg * rb; oy < oy_end; ++oy) {
        __syncthreads();
        for (int i = t; i < TW * OP; i += 256) {
            const int px = i / OP, co = i - px * OP, ox = ox0 + px;
            float val = 0.f;
            if (ox < Wo && co < COUT) {
                const int64_t p = ((int64_t)v * Ho + oy) * Wo + ox;
                val = dOut[p * COUT + co] * lrelu_g(Out[p * COUT + co]);
            }
            g[px][co] = val;
        }
        for (int i = t; i < 3 * IW * CP; i += 256) {
            const int ci = i % CP, xx = (i / CP) % IW, kr = i / (CP * IW);
            const int iy = oy * STRIDE - 1 + kr, ix = ix0 + xx;
            xin[kr][xx][ci] = (ci < CIN && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi) ? In[(((int64_t)v * Hi + iy) * Wi + ix) * CIN + ci] : 0.f;
        }
        __syncthreads();
        if (worker) {
            for (int px = grp; px < TW; px += NG) {
                const float4 gv = *reinterpret_cast<const float4*>(&g[px][co4 * 4]);
                const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4 xv = *reinterpret_cast<const float4*>(&xin[ky][px * STRIDE + kx][ci4 * 4]);
                    const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int d = 0; d < 4; ++d) acc[kx][c][d] = fmaf(gg[c], xx[d], acc[kx][c][d]);
                }
            }
        }
        if (t < COUT)
            for (int px = 0; px < TW; ++px) bsum += g[px][t];
    }
    __syncthreads();
    if (worker) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    const int co = co4 * 4 + c, ci = ci4 * 4 + d;
                    if (co < COUT && ci < CIN && acc[kx][c][d] != 0.f) atomicAdd(&red[(co * CIN + ci) * 9 + ky * 3 + kx], acc[kx][c][d]);
                }
    }
    if (t < COUT) red[NW + t] = bsum;
    __syncthreads();
    for (int i = t; i < NW; i += 256)
        if (red[i] != 0.f) atomicAdd(dW + i, red[i]);
    if (t < COUT && red[NW + t] != 0.f) atomicAdd(db + t, red[NW + t]);
}

template <int CIN, int COUT, int STRIDE>
int launch_fwd(const float* in, const float* W, const float* b, float* out, int V, int Hi, int Wi, int Ho, int Wo, cudaStream_t st) {
    const int64_t n = (int64_t)V * Ho * Wo;
    conv_fwd_kernel<CIN, COUT, STRIDE><<<(unsigned)hnr_cdiv(n, 128), 128, 0, st>>>(in, W, b, out, V, Hi, Wi, Ho, Wo);
    return 0;
}
template <int CIN, int COUT, int STRIDE>
int launch_bwd(const float* dOut, const float* Out, const float* In, const float* W, const float* ext, float* dIn, float* dW, float* db, int V,
               int Hi, int Wi, int Ho, int Wo, cudaStream_t st) {
    // rows per block: ~4 blocks per SM overall (few global atomics per weight, enough blocks to fill the machine)
    const int strips = (Wo + 63) / 64;
    int rb = (int)hnr_cdiv((int64_t)strips * Ho * V, 4 * HNR_NUM_SMS);
    if (rb < 1) rb = 1;
    const int rgroups = (Ho + rb - 1) / rb;
    conv_bwd_weight_kernel<CIN, COUT, STRIDE><<<(unsigned)(strips * rgroups * V), 256, 0, st>>>(dOut, Out, In, dW, db, V, Hi, Wi, Ho, Wo, rb);
    if (dIn) {
        const int64_t n = (int64_t)V * Hi * Wi;
        conv_bwd_data_kernel<CIN, COUT, STRIDE><<<(unsigned)hnr_cdiv(n, 128), 128, 0, st>>>(dOut, Out, W, ext, dIn, V, Hi, Wi, Ho, Wo);
    }
    return 0;
}

inline int down(int x) { return (x - 1) / 2 + 1; }      // conv3x3, stride 2, padding 1

}  // namespace

// Forward of the three pyramid blocks.  img (V,H,W,3) NHWC; w[6] / b[6]: torch-layout (Cout,Cin,3,3) weights and biases of
// s1.conv0, s1.conv1, s2.conv0, s2.conv1, s3.conv0, s3.conv1; act[6]: activated outputs of the six convolutions, NHWC
// (act[1], act[3], act[5] are the pyramid levels (V,H/2,W/2,6), (V,H/4,W/4,12), (V,H/8,W/8,24); the others are kept for backward).
extern "C" int hnr_pyramid_fwd(const float* img, const float* const* w, const float* const* b, float* const* act, int64_t V, int64_t H, int64_t W,
                               void* stream) {
    if (V == 0) return HNR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int v = (int)V, h0 = (int)H, w0 = (int)W, h1 = down(h0), w1 = down(w0), h2 = down(h1), w2 = down(w1), h3 = down(h2), w3 = down(w2);
    launch_fwd<3, 6, 2>(img, w[0], b[0], act[0], v, h0, w0, h1, w1, st);
    launch_fwd<6, 6, 1>(act[0], w[1], b[1], act[1], v, h1, w1, h1, w1, st);
    launch_fwd<6, 12, 2>(act[1], w[2], b[2], act[2], v, h1, w1, h2, w2, st);
    launch_fwd<12, 12, 1>(act[2], w[3], b[3], act[3], v, h2, w2, h2, w2, st);
    launch_fwd<12, 24, 2>(act[3], w[4], b[4], act[4], v, h2, w2, h3, w3, st);
    launch_fwd<24, 24, 1>(act[4], w[5], b[5], act[5], v, h3, w3, h3, w3, st);
    HNR_CHECK_LAUNCH("pyramid_fwd");
    return HNR_OK;
}

// Backward: dlev[3] = gradients w.r.t. the three levels (from the image gather, NHWC, may be NULL = zero), act[6] as saved by the
// forward; dw[6] / db[6] accumulate (zero them first); scratch[5]: gradient buffers shaped like act[4], act[3], act[2], act[1],
// act[0] (written).  No gradient flows to the images.
extern "C" int hnr_pyramid_bwd(const float* img, const float* const* w, const float* const* act, const float* const* dlev, float* const* dw,
                               float* const* db, float* const* scratch, int64_t V, int64_t H, int64_t W, void* stream) {
    if (V == 0) return HNR_OK;
    HNR_CHECK_ARG(dlev[2] != nullptr, "pyramid_bwd: the gradient of the coarsest level is required");
    cudaStream_t st = (cudaStream_t)stream;
    const int v = (int)V, h0 = (int)H, w0 = (int)W, h1 = down(h0), w1 = down(w0), h2 = down(h1), w2 = down(w1), h3 = down(h2), w3 = down(w2);
    // block s3: L3 <- a3 <- L2
    launch_bwd<24, 24, 1>(dlev[2], act[5], act[4], w[5], nullptr, scratch[0], dw[5], db[5], v, h3, w3, h3, w3, st);      // d a3
    launch_bwd<12, 24, 2>(scratch[0], act[4], act[3], w[4], dlev[1], scratch[1], dw[4], db[4], v, h2, w2, h3, w3, st);   // d L2 (+ gather)
    // block s2: L2 <- a2 <- L1
    launch_bwd<12, 12, 1>(scratch[1], act[3], act[2], w[3], nullptr, scratch[2], dw[3], db[3], v, h2, w2, h2, w2, st);   // d a2
    launch_bwd<6, 12, 2>(scratch[2], act[2], act[1], w[2], dlev[0], scratch[3], dw[2], db[2], v, h1, w1, h2, w2, st);    // d L1 (+ gather)
    // block s1: L1 <- a1 <- img
    launch_bwd<6, 6, 1>(scratch[3], act[1], act[0], w[1], nullptr, scratch[4], dw[1], db[1], v, h1, w1, h1, w1, st);     // d a1
    launch_bwd<3, 6, 2>(scratch[4], act[0], img, w[0], nullptr, nullptr, dw[0], db[0], v, h0, w0, h1, w1, st);
    HNR_CHECK_LAUNCH("pyramid_bwd");
    return HNR_OK;
}
