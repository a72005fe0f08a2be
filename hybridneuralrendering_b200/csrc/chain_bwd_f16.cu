// Fused DATA-GRADIENT chain of the per-sample MLPs on the 5th-gen tensor cores (training backward of SURVEY.md §8a rows A4, I3, I5:
// colour-feature branch 280->128->128->128, blend-weight net 176->64->64->64 batched over the V views, colour mix-up
// 90->45->45->45; what autograd computes over models/aggregators/point_aggregators.py:1028-1037, :1188-1217, :1285-1334):
//
//     dZ_top = dY * act'(Y_top)            (loader warps, from the fp32 gradient of the chain's output)
//     dZ_{l-1} = (dZ_l W_l) * act'(H_{l-1})      l = top .. 1          (tensor cores + gating epilogue)
//     dX       =  dZ_0 W_0[:, :NX]                                      (fp32 rows for the chain's producers)
//
// The generic-width sibling of nbr_bwd_f16.cu: same roles, same split-bf16 arithmetic (3 MMAs per product, fp32 accumulation in
// TMEM), the gradient tile stays in shared memory between layers, every dZ_l also goes to HBM as a split image (img_common.cuh)
// for the weight-gradient kernel (wgrad_img.cu), activation signs come from the hi planes of the images the training forward
// (chain_f16.cu) saved.  One persistent CTA per SM, tiles of 128 rows; layer widths <= 128, dX up to 256 columns.
#include "common.cuh"
#include "hnr.h"
#define TRACE_SRC ((long long*)nullptr)
#include "tc_common.cuh"
#include "img_common.cuh"

namespace {
using namespace tc;

constexpr int TM = 128, KC = 16, MAXL = 4;
constexpr int AMAX = 128;               // widest layer
constexpr int XMAX = 256;               // widest dX block
constexpr int NSW = 3, NSA = 3, NSG = 3;
constexpr int W_STAGE = XMAX * KC * 2 * 2;    // 16384: one chunk image of the widest B operand
constexpr int A_PART = TM * KC * 2, A_STAGE = 2 * A_PART;       // 4096, 8192
constexpr int G_STAGE = TM * 32 * 2;    // 8192
constexpr int ACT_PART = TM * AMAX * 2; // 32768
constexpr int NEPI = 256, NTHREADS = 512;     // warps 0-7 epilogue, 8-11 loaders, 12 MMA, 13 bulk copy
constexpr int OFF_W = 0;
constexpr int OFF_A = OFF_W + NSW * W_STAGE;          // 49152
constexpr int OFF_G = OFF_A + NSA * A_STAGE;          // 73728
constexpr int OFF_ACT = OFF_G + NSG * G_STAGE;        // 98304
constexpr int OFF_BAR = OFF_ACT + 2 * ACT_PART;       // 163840
constexpr int NBAR = 2 * NSW + 2 * NSA + 2 * NSG + 4 + 2 + 2;
constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
constexpr uint32_t A_LBO = (TM / 8) * 128, SBO = 128;

struct CBArgs {
    int nl;
    int Np[MAXL], N[MAXL];      // padded / real output width of layer l
    int NX;                     // dX columns computed (multiple of 16, <= 256)
    int act_top;                // HNR_ACT_LRELU or HNR_ACT_NONE
    const float* dY; int lddy;
    const float* Ytop; int ldyt;
    const uint8_t* gimg[MAXL];  // saved H_l images (l < nl-1): activation signs
    uint8_t* dzimg[MAXL];       // dZ_l images (written)
    const uint8_t* wpack;
    int64_t w_off[MAXL];
    float* dX; int ldx;
    int64_t M;
};

__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, unsigned ns) {
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}
__host__ __device__ constexpr uint32_t idesc_bf16(int N) {   // D = f32, A = B = bf16, K-major, M = 128
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ float slope_lo(uint32_t p) { return img::bf16_lo_f(p) > 0.f ? 1.f : 0.01f; }
__device__ __forceinline__ float slope_hi(uint32_t p) { return img::bf16_hi_f(p) > 0.f ? 1.f : 0.01f; }

// 16 consecutive columns col0.. of row `row` of a (rows, ld) fp32 matrix with n real columns (zero beyond)
__device__ __forceinline__ void load16(const float* base, int ld, int64_t row, int col0, int n, float4 (&o)[4]) {
    const float* p = base + row * ld + col0;
    if (col0 + 16 <= n && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __ldg(reinterpret_cast<const float4*>(p) + i);
    } else {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = col0 + i < n ? __ldg(p + i) : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) chain_bwd_f16_kernel(const __grid_constant__ CBArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    const uint32_t b0 = smem_u32(bars);
    const uint32_t bar_wfull = b0, bar_wempty = bar_wfull + 8 * NSW, bar_afull = bar_wempty + 8 * NSW, bar_aempty = bar_afull + 8 * NSA,
                   bar_gfull = bar_aempty + 8 * NSA, bar_gempty = bar_gfull + 8 * NSG, bar_actfull = bar_gempty + 8 * NSG,
                   bar_accfull = bar_actfull + 32, bar_accfree = bar_accfull + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    const int64_t ntiles = (A.M + TM - 1) / TM;
    const int nl = A.nl, top = nl - 1;
    // per-tile item bookkeeping (uniform): n_top operand chunks, then the gate blocks of layers top-1 .. 0
    const int n_top = A.Np[top] / KC;
    int gpre[MAXL + 1];                      // gate-item prefix per step j >= 0 (step j gates with layer top-1-j)
    gpre[0] = 0;
#pragma unroll
    for (int j = 0; j < MAXL; ++j) gpre[j + 1] = gpre[j] + ((j < top) ? (A.Np[top - 1 - j] + 31) / 32 : 0);
    const int ngate = gpre[MAXL];
    const int nitem = n_top + ngate;

    if (tid == 0) {
        for (int s = 0; s < NSW; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1); }
        for (int s = 0; s < NSA; ++s) { mbar_init(bar_afull + 8 * s, 4); mbar_init(bar_aempty + 8 * s, 1); }
        for (int s = 0; s < NSG; ++s) { mbar_init(bar_gfull + 8 * s, 4); mbar_init(bar_gempty + 8 * s, 4); }
        for (int s = 0; s < 4; ++s) mbar_init(bar_actfull + 8 * s, 4);
        for (int s = 0; s < 2; ++s) { mbar_init(bar_accfull + 8 * s, 1); mbar_init(bar_accfree + 8 * s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 13) {
        // ================= bulk-copy producer: one W^T chunk image (ow * 64 bytes) per stage =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int j = 0; j < nl; ++j) {
                    const int l = top - j;
                    const uint32_t ow = (uint32_t)(l > 0 ? A.Np[l - 1] : A.NX), cbytes = ow * 64u;
                    const uint8_t* src = A.wpack + A.w_off[l];
                    const int nc = A.Np[l] / KC;
                    for (int c = 0; c < nc; ++c, ++it) {
                        const uint32_t s = it % NSW, ph = (it / NSW) & 1;
                        mbar_wait_relaxed(bar_wempty + 8 * s, ph ^ 1, 32);
                        mbar_arrive_expect_tx(bar_wfull + 8 * s, cbytes);
                        bulk_g2s(smem_u32(smem + OFF_W + s * W_STAGE), src + (size_t)c * cbytes, cbytes, bar_wfull + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 12) {
        // ================= MMA issuer =================
        uint32_t wit = 0, ait = 0, use[2] = {0, 0}, actcnt[4] = {0, 0, 0, 0};
        const uint32_t act_hi = smem_u32(smem + OFF_ACT), act_lo = act_hi + ACT_PART;
        const uint32_t w_base = smem_u32(smem + OFF_W), a_base = smem_u32(smem + OFF_A);
        const uint64_t dA = umma_desc(0, A_LBO, SBO);
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int j = 0; j < nl; ++j) {
                const int l = top - j;
                const uint32_t b = j & 1, acc = tmem_base + b * XMAX;
                const uint32_t ow = (uint32_t)(l > 0 ? A.Np[l - 1] : A.NX), idesc = idesc_bf16((int)ow), w_lbo = (ow / 8) * 128, w_part = ow * 32;
                const uint64_t dW = umma_desc(0, w_lbo, SBO);
                const int nc = A.Np[l] / KC;
                if (use[b] > 0) mbar_wait_relaxed(bar_accfree + 8 * b, (use[b] - 1) & 1, 20);       // previous readers of this accumulator done
                ++use[b];
                for (int c = 0; c < nc; ++c) {
                    uint32_t a_hi_addr, a_lo_addr, as = 0;
                    if (j == 0) {
                        as = ait % NSA;
                        const uint32_t ph = (ait / NSA) & 1;
                        ++ait;
                        mbar_wait_relaxed(bar_afull + 8 * as, ph, 20);
                        a_hi_addr = a_base + as * A_STAGE;
                        a_lo_addr = a_hi_addr + A_PART;
                    } else {
                        if ((c & 1) == 0) {
                            mbar_wait_relaxed(bar_actfull + 8 * (c >> 1), actcnt[c >> 1] & 1, 20);
                            ++actcnt[c >> 1];
                        }
                        a_hi_addr = act_hi + c * 2 * A_LBO;
                        a_lo_addr = act_lo + c * 2 * A_LBO;
                    }
                    const uint32_t s = wit % NSW;
                    mbar_wait(bar_wfull + 8 * s, (wit / NSW) & 1);
                    ++wit;
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t w = w_base + s * W_STAGE;
                        const uint64_t w_hi = dW | (uint64_t)((w & 0x3FFFFu) >> 4), w_lo = dW | (uint64_t)(((w + w_part) & 0x3FFFFu) >> 4);
                        const uint64_t a_hi = dA | (uint64_t)((a_hi_addr & 0x3FFFFu) >> 4), a_lo = dA | (uint64_t)((a_lo_addr & 0x3FFFFu) >> 4);
                        tc_mma_bf16(acc, a_hi, w_hi, idesc, c > 0 ? 1u : 0u);
                        tc_mma_bf16(acc, a_lo, w_hi, idesc, 1u);
                        tc_mma_bf16(acc, a_hi, w_lo, idesc, 1u);
                        tc_commit(bar_wempty + 8 * s);
                        if (j == 0) tc_commit(bar_aempty + 8 * as);
                        if (c == nc - 1) tc_commit(bar_accfull + 8 * b);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 8 && warp < 12) {
        // ================= loaders: thread = row.  Items of a tile: n_top operand chunks (dY * act'(Y_top), split, also written to
        //                   the dZ_top image), then the gate blocks (hi plane pieces of the saved H images).  Two items in flight. =================
        const int r = tid - NEPI;
        const int64_t my_tiles = (int64_t)blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int64_t total = my_tiles * nitem;
        uint32_t ait = 0, git = 0;
        struct Item { float4 a[4]; float4 y[4]; };
        auto load = [&](int64_t g, Item& it) {
            if (g >= total) return;
            const int64_t tile = blockIdx.x + (g / nitem) * gridDim.x;
            const int i = (int)(g % nitem);
            const int64_t row = tile * TM + r;
            if (i < n_top) {
                if (row < A.M) {
                    load16(A.dY, A.lddy, row, i * KC, A.N[top], it.a);
                    if (A.act_top == HNR_ACT_LRELU) load16(A.Ytop, A.ldyt, row, i * KC, A.N[top], it.y);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) it.a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                int gi = i - n_top, j = 0;
                while (j + 1 < MAXL && gi >= gpre[j + 1]) ++j;
                const int jb = gi - gpre[j], l = top - 1 - j, C = A.Np[l];
                const uint8_t* p = A.gimg[l] + img::piece_off(row, 4 * jb, C);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if ((4 * jb + k) * 8 < C) v = __ldg(reinterpret_cast<const uint4*>(p + k * 512));
                    it.a[k] = *reinterpret_cast<float4*>(&v);
                }
            }
        };
        auto put = [&](int64_t g, const Item& it) {
            if (g >= total) return;
            const int64_t tile = blockIdx.x + (g / nitem) * gridDim.x;
            const int i = (int)(g % nitem);
            if (i < n_top) {
                const int64_t row = tile * TM + r;
                float v[16];
#pragma unroll
                for (int k = 0; k < 4; ++k) { v[4 * k] = it.a[k].x; v[4 * k + 1] = it.a[k].y; v[4 * k + 2] = it.a[k].z; v[4 * k + 3] = it.a[k].w; }
                if (A.act_top == HNR_ACT_LRELU && row < A.M) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        v[4 * k] *= it.y[k].x > 0.f ? 1.f : 0.01f; v[4 * k + 1] *= it.y[k].y > 0.f ? 1.f : 0.01f;
                        v[4 * k + 2] *= it.y[k].z > 0.f ? 1.f : 0.01f; v[4 * k + 3] *= it.y[k].w > 0.f ? 1.f : 0.01f;
                    }
                }
                uint4 h0, l0, h1, l1;
                img::split8_bf16(v, h0, l0);
                img::split8_bf16(v + 8, h1, l1);
                const int C = A.Np[top];
                uint8_t* gp = A.dzimg[top] + img::piece_off(row, 2 * i, C);
                *reinterpret_cast<uint4*>(gp) = h0; *reinterpret_cast<uint4*>(gp + 512) = h1;
                *reinterpret_cast<uint4*>(gp + img::plane_bytes(C)) = l0; *reinterpret_cast<uint4*>(gp + img::plane_bytes(C) + 512) = l1;
                const uint32_t st = ait % NSA, ph = (ait / NSA) & 1;
                ++ait;
                mbar_wait_relaxed(bar_aempty + 8 * st, ph ^ 1, 32);
                uint8_t* stage = smem + OFF_A + st * A_STAGE + r * 16;
                *reinterpret_cast<uint4*>(stage) = h0;
                *reinterpret_cast<uint4*>(stage + A_LBO) = h1;
                *reinterpret_cast<uint4*>(stage + A_PART) = l0;
                *reinterpret_cast<uint4*>(stage + A_PART + A_LBO) = l1;
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_afull + 8 * st);
            } else {
                const uint32_t st = git % NSG, ph = (git / NSG) & 1;
                ++git;
                mbar_wait_relaxed(bar_gempty + 8 * st, ph ^ 1, 32);
                uint8_t* stage = smem + OFF_G + st * G_STAGE + r * 16;
#pragma unroll
                for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(stage + k * 2048) = it.a[k];
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_gfull + 8 * st);
            }
        };
        Item i0, i1;
        load(0, i0);
        load(1, i1);
        for (int64_t g = 0; g < total; g += 2) {
            put(g, i0);
            load(g + 2, i0);
            put(g + 1, i1);
            load(g + 3, i1);
        }
    } else if (warp < 8) {
        // ================= epilogue warps: thread = (row = TMEM lane, group wg taking 32-column blocks wg, wg+2, ...) =================
        const int wg = warp >> 2;
        const int r = tid & (TM - 1);
        const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
        uint8_t* act_hi = smem + OFF_ACT;
        uint32_t use[2] = {0, 0}, ti = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            const int64_t row = tile * TM + r;
#pragma unroll 1
            for (int j = 0; j < nl; ++j) {
                const int l = top - j;
                const uint32_t b = j & 1;
                mbar_wait_relaxed(bar_accfull + 8 * b, use[b] & 1, 20);
                ++use[b];
                tc_fence_after();
                const int ow = l > 0 ? A.Np[l - 1] : A.NX;
                const int nblk = (ow + 31) / 32;
                const uint32_t taddr = tmem_base + lane_base + b * XMAX;
                if (l > 0) {
                    const int C = ow;
                    uint8_t* out = A.dzimg[l - 1];
                    for (int jb = wg; jb < nblk; jb += 2) {
                        const uint32_t gi = ti * (uint32_t)ngate + (uint32_t)gpre[j] + (uint32_t)jb, st = gi % NSG, ph = (gi / NSG) & 1;
                        mbar_wait_relaxed(bar_gfull + 8 * st, ph, 20);
                        uint4 g4[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) g4[k] = *reinterpret_cast<const uint4*>(smem + OFF_G + st * G_STAGE + k * 2048 + r * 16);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_gempty + 8 * st);
                        float acc[32];
                        tmem_ld32(taddr + jb * 32, acc);
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            if (jb * 32 + qd * 8 < ow) {
                                const uint32_t gw[4] = {g4[qd].x, g4[qd].y, g4[qd].z, g4[qd].w};
                                float y[8];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    y[2 * u] = acc[8 * qd + 2 * u] * slope_lo(gw[u]);
                                    y[2 * u + 1] = acc[8 * qd + 2 * u + 1] * slope_hi(gw[u]);
                                }
                                uint4 hi, lo;
                                img::split8_bf16(y, hi, lo);
                                uint8_t* dst = act_hi + (jb * 4 + qd) * A_LBO + r * 16;
                                *reinterpret_cast<uint4*>(dst) = hi;
                                *reinterpret_cast<uint4*>(dst + ACT_PART) = lo;
                                uint8_t* gdst = out + img::piece_off(row, jb * 4 + qd, C);
                                *reinterpret_cast<uint4*>(gdst) = hi;
                                *reinterpret_cast<uint4*>(gdst + img::plane_bytes(C)) = lo;
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_actfull + 8 * jb);
                    }
                } else {
                    for (int jb = wg; jb < nblk; jb += 2) {
                        float acc[32];
                        tmem_ld32(taddr + jb * 32, acc);
                        if (row < A.M) {
                            float4* o = reinterpret_cast<float4*>(A.dX + row * A.ldx + jb * 32);
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (jb * 32 + 4 * i < ow) o[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_accfree + 8 * b);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// Fused data-gradient chain of a per-sample MLP (see the header of this file).  Arrays have nlayer entries.
//   dY (M, lddy): gradient w.r.t. the chain's last output (N[nlayer-1] columns); Ytop: that output (read only when act_top is
//   LeakyReLU); gimg[l]: split image of H_l saved by hnr_chain_f16_forward (l < nlayer-1), Np[l] columns; dzimg[l]: split
//   image of dZ_l, Np[l] columns (written, rows padded to 128); wpackT + w_off[l]: Np[l]/16 chunk images of W_l^T,
//   rows = (l > 0 ? Np[l-1] : NX), built by chain.pack_chain_bwd; dX (M, ldx): first NX input-gradient columns (kernel source
//   order), NX % 16 == 0, NX <= 256, ldx % 4 == 0.
extern "C" int hnr_chain_bwd_f16(int nlayer, const int64_t* Np, const int64_t* N, int64_t NX, int act_top, const float* dY, int64_t lddy,
                                 const float* Ytop, int64_t ldyt, const void* const* gimg, void* const* dzimg, const void* wpackT,
                                 const int64_t* w_off, float* dX, int64_t ldx, int64_t M, void* stream) {
    HNR_CHECK_ARG(nlayer >= 1 && nlayer <= MAXL, "chain_bwd_f16: 1..4 layers");
    HNR_CHECK_ARG(act_top == HNR_ACT_LRELU || act_top == HNR_ACT_NONE, "chain_bwd_f16: top activation must be LeakyReLU or none");
    HNR_CHECK_ARG(NX > 0 && NX % 16 == 0 && NX <= XMAX && ldx >= NX && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(dX) & 15) == 0,
                  "chain_bwd_f16: NX must be a multiple of 16 <= 256, dX rows 16-byte aligned");
    if (M == 0) return HNR_OK;
    CBArgs A{};
    A.nl = nlayer;
    for (int l = 0; l < nlayer; ++l) {
        HNR_CHECK_ARG(Np[l] % 16 == 0 && Np[l] >= 16 && Np[l] <= AMAX && N[l] <= Np[l] && N[l] > 0, "chain_bwd_f16: layer width must be <= 128");
        A.Np[l] = (int)Np[l]; A.N[l] = (int)N[l]; A.w_off[l] = w_off[l];
        A.gimg[l] = (l < nlayer - 1) ? (const uint8_t*)gimg[l] : nullptr;
        A.dzimg[l] = (uint8_t*)dzimg[l];
        HNR_CHECK_ARG(A.dzimg[l] && (l == nlayer - 1 || A.gimg[l]), "chain_bwd_f16: missing image");
    }
    HNR_CHECK_ARG(act_top == HNR_ACT_NONE || Ytop, "chain_bwd_f16: Ytop required for a gated top layer");
    A.NX = (int)NX; A.act_top = act_top; A.dY = dY; A.lddy = (int)lddy; A.Ytop = Ytop; A.ldyt = (int)ldyt;
    A.wpack = (const uint8_t*)wpackT; A.dX = dX; A.ldx = (int)ldx; A.M = M;
    static bool configured = false;
    if (!configured) {
        HNR_CUDA(cudaFuncSetAttribute(chain_bwd_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const int64_t ntiles = hnr_cdiv(M, TM);
    const int grid = (int)(ntiles < HNR_NUM_SMS ? ntiles : HNR_NUM_SMS);
    chain_bwd_f16_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(A);
    HNR_CHECK_LAUNCH("chain_bwd_f16");
    return HNR_OK;
}

// out[n, :] = sum over v < V of the fp32 value (hi + lo) of row v*Nv + n of a split image with C columns: the gradient of a layer-0
// addend shared by the V views of a sample (hnr_chain_f16_forward_add0), read from the dZ_0 image the data-gradient chain wrote.
// thread = (column group of 8, sample n) with n fastest: 32 consecutive samples read 512 contiguous bytes of a slab plane.
namespace {
__global__ void __launch_bounds__(256) img_sum_views_kernel(const uint8_t* __restrict__ im, int C, int64_t Nv, int V, float* __restrict__ out, int ldo) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = idx % Nv;
    const int g = (int)(idx / Nv);
    if (g >= C / 8) return;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int64_t pl = img::plane_bytes(C);
    for (int v = 0; v < V; ++v) {
        const uint8_t* p = im + img::piece_off((int64_t)v * Nv + n, g, C);
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(p)), lo = __ldg(reinterpret_cast<const uint4*>(p + pl));
        float t[8];
        img::join8_bf16(hi, lo, t);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += t[i];
    }
    float4* o = reinterpret_cast<float4*>(out + n * ldo + 8 * g);
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}
}  // namespace

extern "C" int hnr_img_sum_views(const void* im, int64_t C, int64_t Nv, int64_t V, float* out, int64_t ldo, void* stream) {
    HNR_CHECK_ARG(C % 8 == 0 && ldo % 4 == 0 && ldo >= C && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "img_sum_views: C % 8, aligned output rows");
    if (Nv == 0 || V == 0) return HNR_OK;
    const int64_t total = Nv * (C / 8);
    img_sum_views_kernel<<<(unsigned)hnr_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)im, (int)C, Nv, (int)V, out, (int)ldo);
    HNR_CHECK_LAUNCH("img_sum_views");
    return HNR_OK;
}
