// Fused dense-layer CHAIN on the 5th-gen tensor cores (3xFP16 split, fp32 accuracy) for the per-sample MLPs of the
// aggregator: colour-feature branch 280->128->128->128, blend-weight net
// 176->64->64->64->1 (sigmoid head) batched over the V reference views, colour mix-up 90->45->45->45 (+ residual).
// Reference layers: models/aggregators/point_aggregators.py:556-683, used at :1028-1037, :1188-1217, :1285-1334.
//
//   x_0 = concat(src_0, src_1, src_2)[m, :]            (fp32 rows in global memory, row re-use via `mod`)
//   x_{l+1} = act_l(x_l W_l^T + b_l)                   l = 0 .. nlayer-1, widths <= 128
//   out = x_nlayer (+ residual)  and / or  head_act(x_nlayer . head_w + head_b)
//
// One launch runs the whole chain; the activations between layers never leave the SM:
//   * a tile is 128 rows; two CTAs are resident per SM (256 threads, 2 x 128 TMEM columns each) so that one tile's
//     epilogue overlaps the other's MMAs across the serial layer dependency;
//   * operands are split fp16 hi/lo (pre-multiplied by a power-of-two scale); a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
//     with three tcgen05.mma.kind::f16 into one fp32 TMEM accumulator (22 mantissa bits);
//   * layer 0: two generator warps read the fp32 source rows, split them and feed an operand ring in the canonical
//     no-swizzle K-major UMMA layout; layers >= 1: the epilogue of layer l writes layer l+1's A operand directly in
//     that layout (hi+lo fp16 = the 4 bytes of the fp32 value) and releases it to the MMA warp per 32 columns;
//   * weights come pre-split / pre-tiled from the host (chain.py), one cp.async.bulk per 16-wide K chunk;
//   * warp roles: 4 epilogue warps (thread = row = TMEM lane), 2 generator warps, 1 MMA warp, 1 bulk-copy warp.
// Optional fp32 copies of every layer's output (Y_l) make the same kernel usable as the forward of a training step.
#include <cuda_fp16.h>

#include "common.cuh"
#include "hnr.h"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int TM = 128;
constexpr int NMAX = 128;               // widest layer == TMEM columns per accumulator
constexpr int KC = 16;
constexpr int MAXL = 4;
constexpr int NSW = 2, NSA = 3;
constexpr int W_STAGE = 2 * NMAX * KC * 2;   // 8192
constexpr int A_PART = TM * KC * 2;          // 4096
constexpr int A_STAGE = 2 * A_PART;          // 8192
constexpr int ACT_PART = TM * NMAX * 2;      // 32768
constexpr int NTHREADS = 256;
constexpr int NGEN = 64;

constexpr int OFF_W = 0;
constexpr int OFF_A = OFF_W + NSW * W_STAGE;           // 16384
constexpr int OFF_ACT = OFF_A + NSA * A_STAGE;         // 40960
constexpr int OFF_BIAS = OFF_ACT + 2 * ACT_PART;       // 106496
constexpr int OFF_HEAD = OFF_BIAS + MAXL * NMAX * 4;   // 108544
constexpr int OFF_BAR = OFF_HEAD + NMAX * 4;           // 109056
constexpr int NBAR = 2 * NSW + 2 * NSA + 4 + 2 + 2;
constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
static_assert(2 * (SMEM_BYTES + 1024) <= 227 * 1024, "two CTAs per SM");
constexpr uint32_t A_LBO = (TM / 8) * 128, SBO = 128;

struct ChainArgs {
    const float* src[3];
    int ld[3], k[3];
    int64_t mod[3];
    float in_scale;
    int nlayer;
    int Kp[MAXL], N[MAXL], Np[MAXL], act[MAXL];
    int64_t w_off[MAXL];
    const uint8_t* wpack;
    const float* bias;           // (MAXL, NMAX), pre-scaled, zero padded
    float mul[MAXL];             // accumulator -> (scaled) pre-activation
    float inv_next[MAXL];        // 1 / input scale of layer l+1 (to recover the unscaled output for Y_l), 1 for the last
    float* Y[MAXL];
    int ldy[MAXL];
    const float* res;
    int ldres;
    const float* head_w;         // (N_last) or NULL
    const float* head_b;
    int head_act;
    float* head_out;
    int64_t M;
};

__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                   "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
__device__ __forceinline__ uint32_t pack_sat(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        h[i] = pack_sat(a, b);
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
        l[i] = pack_sat(a - hf.x, b - hf.y);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
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
__host__ __device__ constexpr uint32_t idesc_f16(int N) {   // D=f32, A=B=f16, both K-major, M=128
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// 16 consecutive columns [c0, c0+16) of row m of the concatenated layer-0 input (zero beyond the last column)
__device__ __forceinline__ void load_chunk(const ChainArgs& A, int64_t m, int c0, float (&v)[16]) {
    int base = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int ks = A.k[s];
        if (ks > 0 && c0 >= base && c0 + 16 <= base + ks) {             // whole chunk inside source s
            const int64_t row = A.mod[s] > 0 ? m % A.mod[s] : m;
            const float* p = A.src[s] + row * A.ld[s] + (c0 - base);
            if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(p) + i);
                    v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __ldg(p + i);
            }
            return;
        }
        base += ks;
    }
    // chunk straddles sources (or the zero padding): per-column lookup
    const int b1 = A.k[0], b2 = A.k[0] + A.k[1], b3 = b2 + A.k[2];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int c = c0 + i;
        float x = 0.f;
        if (c < b3) {
            const int s = c < b1 ? 0 : (c < b2 ? 1 : 2);
            const int off = c - (s == 0 ? 0 : (s == 1 ? b1 : b2));
            const int64_t row = A.mod[s] > 0 ? m % A.mod[s] : m;
            x = __ldg(A.src[s] + row * A.ld[s] + off);
        }
        v[i] = x;
    }
}

__global__ void __launch_bounds__(NTHREADS, 2) chain_f16_kernel(const __grid_constant__ ChainArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    const uint32_t b0 = smem_u32(bars);
    const uint32_t bar_wfull = b0, bar_wempty = b0 + 8 * NSW, bar_afull = bar_wempty + 8 * NSW, bar_aempty = bar_afull + 8 * NSA,
                   bar_actfull = bar_aempty + 8 * NSA, bar_accfull = bar_actfull + 32, bar_accfree = bar_accfull + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
    float* head_s = reinterpret_cast<float*>(smem + OFF_HEAD);
    const int64_t ntiles = (A.M + TM - 1) / TM;
    const int nl = A.nlayer;

    if (tid == 0) {
        for (int s = 0; s < NSW; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1); }
        for (int s = 0; s < NSA; ++s) { mbar_init(bar_afull + 8 * s, 2); mbar_init(bar_aempty + 8 * s, 1); }
        for (int s = 0; s < 4; ++s) mbar_init(bar_actfull + 8 * s, 4);
        for (int s = 0; s < 2; ++s) { mbar_init(bar_accfull + 8 * s, 1); mbar_init(bar_accfree + 8 * s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 6) tmem_alloc(smem_u32(tmem_slot), 2 * NMAX);
    for (int i = tid; i < MAXL * NMAX; i += NTHREADS) bias_s[i] = A.bias[i];
    if (tid < NMAX) head_s[tid] = (A.head_w && tid < A.N[nl - 1]) ? A.head_w[tid] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 7) {
        // ================= bulk-copy producer: one weight chunk image (Np*64 bytes) per stage =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int l = 0; l < nl; ++l) {
                    const uint32_t bytes = (uint32_t)A.Np[l] * 64u;
                    const uint8_t* src = A.wpack + A.w_off[l];
                    const int nc = A.Kp[l] / KC;
                    for (int c = 0; c < nc; ++c, ++it) {
                        const uint32_t s = it % NSW, ph = (it / NSW) & 1;
                        mbar_wait_relaxed(bar_wempty + 8 * s, ph ^ 1, 64);
                        mbar_arrive_expect_tx(bar_wfull + 8 * s, bytes);
                        bulk_g2s(smem_u32(smem + OFF_W + s * W_STAGE), src + (size_t)c * bytes, bytes, bar_wfull + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 6) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t wit = 0, ait = 0, use[2] = {0, 0}, actcnt[4] = {0, 0, 0, 0};
            const uint32_t act_hi = smem_u32(smem + OFF_ACT), act_lo = act_hi + ACT_PART;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int l = 0; l < nl; ++l) {
                    const uint32_t b = l & 1, acc = tmem_base + b * NMAX;
                    const uint32_t np = (uint32_t)A.Np[l], idesc = idesc_f16((int)np), w_lbo = (np / 8) * 128, w_part = np * 32;
                    const int nc = A.Kp[l] / KC;
                    if (use[b] > 0) mbar_wait_relaxed(bar_accfree + 8 * b, (use[b] - 1) & 1, 20);   // previous reader of this accumulator done
                    ++use[b];
                    for (int c = 0; c < nc; ++c) {
                        uint32_t a_hi_addr, a_lo_addr, as = 0;
                        if (l == 0) {
                            as = ait % NSA;
                            const uint32_t ph = (ait / NSA) & 1;
                            ++ait;
                            mbar_wait_relaxed(bar_afull + 8 * as, ph, 20);
                            a_hi_addr = smem_u32(smem + OFF_A + as * A_STAGE);
                            a_lo_addr = a_hi_addr + A_PART;
                        } else {
                            if ((c & 1) == 0) {
                                mbar_wait_relaxed(bar_actfull + 8 * (c >> 1), actcnt[c >> 1] & 1, 20);
                                ++actcnt[c >> 1];
                            }
                            a_hi_addr = act_hi + c * 2 * A_LBO;
                            a_lo_addr = act_lo + c * 2 * A_LBO;
                        }
                        const uint32_t s = wit % NSW, ph = (wit / NSW) & 1;
                        ++wit;
                        mbar_wait(bar_wfull + 8 * s, ph);
                        tc_fence_after();
                        const uint32_t w = smem_u32(smem + OFF_W + s * W_STAGE);
                        const uint64_t w_hi = umma_desc(w, w_lbo, SBO), w_lo = umma_desc(w + w_part, w_lbo, SBO);
                        const uint64_t a_hi = umma_desc(a_hi_addr, A_LBO, SBO), a_lo = umma_desc(a_lo_addr, A_LBO, SBO);
                        tc_mma_f16(acc, a_hi, w_hi, idesc, c > 0 ? 1u : 0u);
                        tc_mma_f16(acc, a_lo, w_hi, idesc, 1u);
                        tc_mma_f16(acc, a_hi, w_lo, idesc, 1u);
                        tc_commit(bar_wempty + 8 * s);
                        if (l == 0) tc_commit(bar_aempty + 8 * as);
                    }
                    tc_commit(bar_accfull + 8 * b);
                }
            }
        }
    } else if (warp >= 4) {
        // ================= generators: fp32 source rows -> split fp16 operand chunks (thread = 2 rows) =================
        const int g = tid - 128;                         // 0..63
        const int nc0 = A.Kp[0] / KC;
        const float sc = A.in_scale;
        uint32_t ait = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t m0 = tile * TM + g, m1 = m0 + 64;
            const bool ok0 = m0 < A.M, ok1 = m1 < A.M;
            float va[16], vb[16], na[16], nb[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { va[i] = vb[i] = 0.f; }
            if (ok0) load_chunk(A, m0, 0, va);
            if (ok1) load_chunk(A, m1, 0, vb);
            for (int c = 0; c < nc0; ++c) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { na[i] = nb[i] = 0.f; }
                if (c + 1 < nc0) {                       // next chunk's loads in flight while this one is converted
                    if (ok0) load_chunk(A, m0, (c + 1) * KC, na);
                    if (ok1) load_chunk(A, m1, (c + 1) * KC, nb);
                }
                const uint32_t st = ait % NSA, ph = (ait / NSA) & 1;
                ++ait;
                mbar_wait_relaxed(bar_aempty + 8 * st, ph ^ 1, 32);
                uint8_t* stage = smem + OFF_A + st * A_STAGE;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float* v = h ? vb : va;
                    const int r = g + 64 * h;
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] *= sc;
                    uint4 hi, lo;
                    split8(v, hi, lo);
                    *reinterpret_cast<uint4*>(stage + r * 16) = hi;
                    *reinterpret_cast<uint4*>(stage + A_PART + r * 16) = lo;
                    split8(v + 8, hi, lo);
                    *reinterpret_cast<uint4*>(stage + A_LBO + r * 16) = hi;
                    *reinterpret_cast<uint4*>(stage + A_PART + A_LBO + r * 16) = lo;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_afull + 8 * st);
#pragma unroll
                for (int i = 0; i < 16; ++i) { va[i] = na[i]; vb[i] = nb[i]; }
            }
        }
    } else {
        // ================= epilogue warps: thread = row = TMEM lane =================
        const int r = tid;
        const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
        uint8_t* act_hi = smem + OFF_ACT;
        uint32_t use[2] = {0, 0};
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t m = tile * TM + r;
            const bool live = m < A.M;
#pragma unroll 1
            for (int l = 0; l < nl; ++l) {
                const uint32_t b = l & 1;
                mbar_wait_relaxed(bar_accfull + 8 * b, use[b] & 1, 20);
                ++use[b];
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_base + b * NMAX;
                const int np = A.Np[l], n = A.N[l], act = A.act[l];
                const int nblk = (np + 31) >> 5;
                const bool last = (l == nl - 1);
                const float mul = A.mul[l], inv_next = A.inv_next[l];
                const float* bl = bias_s + l * NMAX;
                float* yout = A.Y[l];
                const int ldy = A.ldy[l];
                float dot = 0.f;
                uint32_t va[32], vb[32];
                tmem_ld32_issue(taddr, va);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j < nblk) {
                        uint32_t(&cur)[32] = (j & 1) ? vb : va;
                        uint32_t(&nxt)[32] = (j & 1) ? va : vb;
                        tmem_ld_wait(cur);
                        if (j + 1 < nblk) tmem_ld32_issue(taddr + (j + 1) * 32, nxt);
                        float y[32];
                        const float4* b4 = reinterpret_cast<const float4*>(bl + j * 32);
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            const float4 bb = b4[i4];
                            y[4 * i4 + 0] = fmaf(__uint_as_float(cur[4 * i4 + 0]), mul, bb.x);
                            y[4 * i4 + 1] = fmaf(__uint_as_float(cur[4 * i4 + 1]), mul, bb.y);
                            y[4 * i4 + 2] = fmaf(__uint_as_float(cur[4 * i4 + 2]), mul, bb.z);
                            y[4 * i4 + 3] = fmaf(__uint_as_float(cur[4 * i4 + 3]), mul, bb.w);
                        }
                        if (act == HNR_ACT_LRELU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) y[i] = fmaxf(y[i], 0.01f * y[i]);
                        } else if (act != HNR_ACT_NONE) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) y[i] = apply_act(y[i], act);
                        }
                        if (np - j * 32 < 32) {                       // columns beyond Np were never written by the MMA
#pragma unroll
                            for (int i = 16; i < 32; ++i) y[i] = 0.f;
                        }
                        if (last) {
                            if (A.res && live) {
#pragma unroll
                                for (int i = 0; i < 32; ++i)
                                    if (j * 32 + i < n) y[i] += __ldg(A.res + m * A.ldres + j * 32 + i);
                            }
                            if (A.head_w) {
                                const float4* h4 = reinterpret_cast<const float4*>(head_s + j * 32);
#pragma unroll
                                for (int i4 = 0; i4 < 8; ++i4) {
                                    const float4 hh = h4[i4];
                                    dot = fmaf(y[4 * i4 + 0], hh.x, dot); dot = fmaf(y[4 * i4 + 1], hh.y, dot);
                                    dot = fmaf(y[4 * i4 + 2], hh.z, dot); dot = fmaf(y[4 * i4 + 3], hh.w, dot);
                                }
                            }
                        }
                        if (yout && live) {
                            float* o = yout + m * ldy + j * 32;
                            if ((ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(yout) & 15) == 0 && j * 32 + 32 <= n) {
#pragma unroll
                                for (int i4 = 0; i4 < 8; ++i4)
                                    reinterpret_cast<float4*>(o)[i4] = make_float4(y[4 * i4] * inv_next, y[4 * i4 + 1] * inv_next,
                                                                                   y[4 * i4 + 2] * inv_next, y[4 * i4 + 3] * inv_next);
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i)
                                    if (j * 32 + i < n) o[i] = y[i] * inv_next;
                            }
                        }
                        if (!last) {
                            const int nkb = min(4, (np - j * 32) >> 3);       // 8-column k-blocks of this block that exist
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (q < nkb) {
                                    uint4 hi, lo;
                                    split8(y + q * 8, hi, lo);
                                    uint8_t* dst = act_hi + (j * 4 + q) * A_LBO + r * 16;
                                    *reinterpret_cast<uint4*>(dst) = hi;
                                    *reinterpret_cast<uint4*>(dst + ACT_PART) = lo;
                                }
                            }
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bar_actfull + 8 * j);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_accfree + 8 * b);
                if (last && A.head_w && live) A.head_out[m] = apply_act(dot + A.head_b[0], A.head_act);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 6) tmem_dealloc(tmem_base, 2 * NMAX);
}

}  // namespace

extern "C" int64_t hnr_chain_f16_chunk_bytes(int64_t Np) { return Np * 64; }

// Fused chain of up to 4 dense layers (widths <= 128) over M rows, 3xFP16 on tcgen05 (see the header of this file).
// Arrays have nlayer entries.  Kp[0] = concat width padded to 16, Kp[l] = Np[l-1]; Np = N padded to 16.
// wpack / bias / mul / inv_next: built by hybridneuralrendering_b200/chain.py.  Y[l] may be NULL (inner layers) --
// at least one of Y[nlayer-1] and head_out must be given.
extern "C" int hnr_chain_f16_forward(const float* const* src, const int64_t* src_ld, const int64_t* src_k, const int64_t* src_mod,
                                     float in_scale, int nlayer, const int64_t* Kp, const int64_t* N, const int64_t* Np, const int* act,
                                     const void* wpack, const int64_t* w_off, const float* bias, const float* mul, const float* inv_next,
                                     float* const* Y, const int64_t* ldy, const float* res, int64_t ldres, const float* head_w,
                                     const float* head_b, int head_act, float* head_out, int64_t M, void* stream) {
    HNR_CHECK_ARG(nlayer >= 1 && nlayer <= MAXL, "chain_f16_forward: 1..4 layers");
    if (M == 0) return HNR_OK;
    ChainArgs A{};
    int64_t ksum = 0;
    for (int i = 0; i < 3; ++i) {
        A.src[i] = src[i]; A.ld[i] = (int)src_ld[i]; A.k[i] = (int)src_k[i]; A.mod[i] = src_mod ? src_mod[i] : 0;
        HNR_CHECK_ARG(src_k[i] == 0 || src[i] != nullptr, "chain_f16_forward: null source");
        ksum += src_k[i];
    }
    HNR_CHECK_ARG(ksum > 0 && Kp[0] >= ksum && Kp[0] % KC == 0, "chain_f16_forward: Kp[0] must be the concat width padded to 16");
    for (int l = 0; l < nlayer; ++l) {
        HNR_CHECK_ARG(Np[l] % 16 == 0 && Np[l] >= 16 && Np[l] <= NMAX && N[l] <= Np[l] && N[l] > 0, "chain_f16_forward: layer width must be <= 128");
        HNR_CHECK_ARG(l == 0 || Kp[l] == Np[l - 1], "chain_f16_forward: Kp[l] must equal Np[l-1]");
        A.Kp[l] = (int)Kp[l]; A.N[l] = (int)N[l]; A.Np[l] = (int)Np[l]; A.act[l] = act[l]; A.w_off[l] = w_off[l];
        A.mul[l] = mul[l]; A.inv_next[l] = inv_next[l]; A.Y[l] = Y ? Y[l] : nullptr; A.ldy[l] = ldy ? (int)ldy[l] : 0;
        HNR_CHECK_ARG(l == nlayer - 1 || act[l] == HNR_ACT_LRELU || act[l] == HNR_ACT_NONE, "chain_f16_forward: inner activations must be LeakyReLU or none");
    }
    HNR_CHECK_ARG(A.Y[nlayer - 1] || (head_w && head_out), "chain_f16_forward: no output requested");
    HNR_CHECK_ARG(!head_w || (head_b && head_out), "chain_f16_forward: head needs head_b and head_out");
    A.in_scale = in_scale; A.nlayer = nlayer; A.wpack = (const uint8_t*)wpack; A.bias = bias; A.res = res; A.ldres = (int)ldres;
    A.head_w = head_w; A.head_b = head_b; A.head_act = head_act; A.head_out = head_out; A.M = M;
    static bool configured = false;
    if (!configured) {
        HNR_CUDA(cudaFuncSetAttribute(chain_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const int64_t ntiles = hnr_cdiv(M, TM);
    const int grid = (int)(ntiles < 2 * HNR_NUM_SMS ? ntiles : 2 * HNR_NUM_SMS);
    chain_f16_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(A);
    HNR_CHECK_LAUNCH("chain_f16_forward");
    return HNR_OK;
}
