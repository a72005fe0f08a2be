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
//   * a tile is 128 rows; two CTAs are resident per SM (320 threads, 2 x 128 TMEM columns each) so that one tile's
//     epilogue overlaps the other's MMAs across the serial layer dependency;
//   * operands are split fp16 hi/lo (pre-multiplied by a power-of-two scale); a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
//     with three tcgen05.mma.kind::f16 into one fp32 TMEM accumulator (22 mantissa bits);
//   * layer 0: two generator warps read the fp32 source rows, split them and feed an operand ring in the canonical
//     no-swizzle K-major UMMA layout; layers >= 1: the epilogue of layer l writes layer l+1's A operand directly in
//     that layout (hi+lo fp16 = the 4 bytes of the fp32 value) and releases it to the MMA warp per 32 columns;
//   * weights come pre-split / pre-tiled from the host (chain.py), one cp.async.bulk per 16-wide K chunk;
//   * warp roles: 8 worker warps that first generate the tile's layer-0 operand (two chunk pairs of loads in flight per
//     warp: the chains are bound by the latency of reading their inputs) and then run the epilogues of its layers (thread =
//     row = TMEM lane, the 32-column blocks split between the two warps of a lane quarter), 1 MMA warp, 1 bulk-copy warp.
//     A tile's phases depend on each other serially, so sharing the warps costs no overlap (the second CTA on the SM
//     supplies it) and puts two warps per scheduler on every phase; clock64 traces (scripts/trace_chain.py) showed the
//     round-1 split (4 generator + 4 epilogue warps, one pair in flight) at ~0.1 IPC per warp on pure latency chains.
// Optional fp32 copies of every layer's output (Y_l) make the same kernel usable as the forward of a training step.
#include <cuda_fp16.h>
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "hnr.h"
// the clock64 event trace (scripts/trace_chain.py) costs predicated instructions in every hot loop of an issue-bound kernel: it is
// compiled in only with -DHNR_CHAIN_TRACE
#ifdef HNR_CHAIN_TRACE
#define TRACE_SRC A.trace
#else
#define TRACE_SRC ((long long*)nullptr)
#endif
#include "tc_common.cuh"
#include "img_common.cuh"

namespace {
using namespace tc;

constexpr int TM = 128;
constexpr int NMAX = 128;               // widest layer == TMEM columns per accumulator
constexpr int KC = 16;
constexpr int MAXL = 4;
constexpr int NSW = 2, NSA = 4;           // weight stages (8 KB, several chunks each when narrow) / layer-0 operand stages (pairs)
constexpr int W_STAGE = 2 * NMAX * KC * 2;   // 8192
constexpr int A_PART = TM * KC * 2;          // 4096
constexpr int A_STAGE = 2 * A_PART;          // 8192
constexpr int ACT_PART = TM * NMAX * 2;      // 32768
constexpr int NTHREADS = 320;            // warps 0-7 workers (layer-0 generation + epilogues), 8 MMA, 9 bulk copy

constexpr int OFF_W = 0;
constexpr int OFF_A = OFF_W + NSW * W_STAGE;           // 16384
constexpr int OFF_ACT = OFF_A + NSA * A_STAGE;         // 49152
constexpr int OFF_BAR = OFF_ACT + 2 * ACT_PART;        // 114688 (bias / head weights are read through the read-only cache)
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
    int32_t* status;             // optional: bit 1 is set when a scaled value left fp16's range and was saturated
    // training forward: split images (img_common.cuh) of the concatenated input (Kp[0] columns, kernel source order) and of the
    // inner layers' outputs (Np[l] columns), rows padded to the tile -- operands of chain_bwd_f16.cu / wgrad_img.cu.  NULL = off
    uint8_t* x0img;
    uint8_t* himg[MAXL];
    // optional addend of layer 0's pre-activation: row (m % add0_mod) of an fp32 (rows, N[0]) matrix, times add0_scale.  Used by the
    // blend-weight net: its first layer's input [g | aux_v | dview_v] shares g between the V views of a sample, so W0g.g is computed
    // once per sample by another launch and added here; this chain then only reads the 48 view-dependent columns
    const float* add0;
    int ld_add0;
    int64_t mod_add0;
    float add0_scale;
    int variant;                 // bring-up / profiling switches (HNR_CHAIN_VARIANT): 1 = no L2 prefetch, 2 = L1-bypassing loads, 4 = no source loads at all
    long long* trace;            // optional event trace of CTA 0 (bring-up / profiling): [count, (clock, id, a, b) ...]
};

__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
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
// mbarrier wait for warps that are ahead of the pipeline anyway.  The kernel is instruction-issue bound (ncu: IPC 2.0 of 4 with a
// quarter of all executed instructions in try_wait / nanosleep polling loops), so the wait is left to the hardware: try_wait with a
// suspend-time hint parks the warp until the phase completes or the hint expires.  (`ns` is kept for the call sites' documentation.)
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, unsigned ns) {
    (void)ns;
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
        if (done) break;
    }
}
// epilogue of an INNER layer for a block of W (32 or 16) accumulator columns of one row: bias, LeakyReLU / identity, optional split
// bf16 image of the output (training), fp16 hi/lo split into the next layer's A operand, release of the block to the MMA warp
template <int W>
__device__ __forceinline__ void epi_inner_block(uint32_t taddr, int c0, const float* __restrict__ bl, float mul, float inv_next, bool do_lrelu,
                                                uint8_t* himg, int np, int64_t m, uint8_t* dst, int32_t* status, uint32_t bar_block, int lane,
                                                const float* __restrict__ addrow, float add_scale) {
    float y[W];
    float4 bb[W / 4];
#pragma unroll
    for (int i4 = 0; i4 < W / 4; ++i4) bb[i4] = __ldg(reinterpret_cast<const float4*>(bl + c0) + i4);     // in flight together with the TMEM load
    if (addrow) {                                  // bias + scaled addend (rows beyond M pass NULL: their values are never used)
#pragma unroll
        for (int i4 = 0; i4 < W / 4; ++i4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(addrow + c0) + i4);
            bb[i4].x = fmaf(a.x, add_scale, bb[i4].x); bb[i4].y = fmaf(a.y, add_scale, bb[i4].y);
            bb[i4].z = fmaf(a.z, add_scale, bb[i4].z); bb[i4].w = fmaf(a.w, add_scale, bb[i4].w);
        }
    }
    if constexpr (W == 32) tmem_ld32(taddr + c0, y); else tmem_ld16(taddr + c0, y);
#pragma unroll
    for (int i4 = 0; i4 < W / 4; ++i4) {
        y[4 * i4 + 0] = fmaf(y[4 * i4 + 0], mul, bb[i4].x);
        y[4 * i4 + 1] = fmaf(y[4 * i4 + 1], mul, bb[i4].y);
        y[4 * i4 + 2] = fmaf(y[4 * i4 + 2], mul, bb[i4].z);
        y[4 * i4 + 3] = fmaf(y[4 * i4 + 3], mul, bb[i4].w);
    }
    if (do_lrelu) {
#pragma unroll
        for (int i = 0; i < W; ++i) y[i] = fmaxf(y[i], 0.01f * y[i]);
    }
    if (himg) {
        uint8_t* gp = himg + img::piece_off(m, c0 >> 3, np);
        const int64_t pl = img::plane_bytes(np);
#pragma unroll
        for (int g = 0; g < W / 8; ++g) {
            float u[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = y[8 * g + i] * inv_next;
            uint4 hi, lo;
            img::split8_bf16(u, hi, lo);
            *reinterpret_cast<uint4*>(gp + 512 * g) = hi;
            *reinterpret_cast<uint4*>(gp + 512 * g + pl) = lo;
        }
    }
    float am = 0.f;
#pragma unroll
    for (int i = 0; i < W; ++i) am = fmaxf(am, fabsf(y[i]));
    if (am > 65000.f && status) atomicOr(status, 2);
#pragma unroll
    for (int g = 0; g < W / 8; ++g) {
        uint4 hi, lo;
        split8(y + 8 * g, hi, lo);
        *reinterpret_cast<uint4*>(dst + g * A_LBO) = hi;
        *reinterpret_cast<uint4*>(dst + g * A_LBO + ACT_PART) = lo;
    }
    fence_proxy_async();                    // the block (32 columns, or the 16-column tail) is complete: release it to the MMA warp
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_block);
}
__host__ __device__ constexpr uint32_t idesc_f16(int N) {   // D=f32, A=B=f16, both K-major, M=128
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

__global__ void __launch_bounds__(NTHREADS, 2) chain_f16_kernel(const __grid_constant__ ChainArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    const uint32_t b0 = smem_u32(bars);
    const uint32_t bar_wfull = b0, bar_wempty = b0 + 8 * NSW, bar_afull = bar_wempty + 8 * NSW, bar_aempty = bar_afull + 8 * NSA,
                   bar_actfull = bar_aempty + 8 * NSA, bar_accfull = bar_actfull + 32, bar_accfree = bar_accfull + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    const int64_t ntiles = (A.M + TM - 1) / TM;
    const int nl = A.nlayer;

    if (tid == 0) {
        for (int s = 0; s < NSW; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1); }
        for (int s = 0; s < NSA; ++s) { mbar_init(bar_afull + 8 * s, 4); mbar_init(bar_aempty + 8 * s, 1); }
        for (int s = 0; s < 4; ++s) mbar_init(bar_actfull + 8 * s, 4);
        for (int s = 0; s < 2; ++s) { mbar_init(bar_accfull + 8 * s, 1); mbar_init(bar_accfree + 8 * s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 2 * NMAX);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 9) {
        // ================= bulk-copy producer: one stage = as many whole weight chunks (Np*64 bytes each) as fit in 8 KB =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int l = 0; l < nl; ++l) {
                    const uint32_t cbytes = (uint32_t)A.Np[l] * 64u;
                    const int cps = W_STAGE / (int)cbytes;             // chunks per stage
                    const uint8_t* src = A.wpack + A.w_off[l];
                    const int nc = A.Kp[l] / KC;
                    for (int c = 0; c < nc; c += cps, ++it) {
                        const uint32_t s = it % NSW, ph = (it / NSW) & 1;
                        const uint32_t bytes = (uint32_t)min(cps, nc - c) * cbytes;
                        mbar_wait_relaxed(bar_wempty + 8 * s, ph ^ 1, 32);
                        mbar_arrive_expect_tx(bar_wfull + 8 * s, bytes);
                        bulk_g2s(smem_u32(smem + OFF_W + s * W_STAGE), src + (size_t)c * cbytes, bytes, bar_wfull + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 8) {
        // ================= MMA issuer: the whole warp walks the schedule (warp-uniform control flow), one elected lane issues =================
        {
            uint32_t wit = 0, ait = 0, use[2] = {0, 0}, actcnt[4] = {0, 0, 0, 0};
            TRACE_DECL(0);
            if (lane != 0) tr__ = nullptr;
            const uint32_t act_hi = smem_u32(smem + OFF_ACT), act_lo = act_hi + ACT_PART;
            const uint32_t w_base = smem_u32(smem + OFF_W), a_base = smem_u32(smem + OFF_A);
            const uint64_t dA = umma_desc(0, A_LBO, SBO);
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int l = 0; l < nl; ++l) {
                    const uint32_t b = l & 1, acc = tmem_base + b * NMAX;
                    const uint32_t np = (uint32_t)A.Np[l], idesc = idesc_f16((int)np), w_lbo = (np / 8) * 128, w_part = np * 32;
                    const uint64_t dW = umma_desc(0, w_lbo, SBO);
                    const int nc = A.Kp[l] / KC, cps = W_STAGE / (int)(np * 64u);
                    if (use[b] > 0) mbar_wait_relaxed(bar_accfree + 8 * b, (use[b] - 1) & 1, 20);   // previous reader of this accumulator done
                    ++use[b];
                    int sub = 0;
                    for (int c = 0; c < nc; ++c) {
                        uint32_t a_hi_addr, a_lo_addr, as = 0;
                        if (l == 0) {
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
                        TRACE(1, l, c);                                 // operand ready
                        const uint32_t s = wit % NSW;
                        if (sub == 0) mbar_wait(bar_wfull + 8 * s, (wit / NSW) & 1);
                        tc_fence_after();
                        const bool release_w = (sub == cps - 1 || c == nc - 1);
                        if (elect_one()) {
                            const uint32_t w = w_base + s * W_STAGE + (uint32_t)sub * (np * 64u);
                            const uint64_t w_hi = dW | (uint64_t)((w & 0x3FFFFu) >> 4), w_lo = dW | (uint64_t)(((w + w_part) & 0x3FFFFu) >> 4);
                            const uint64_t a_hi = dA | (uint64_t)((a_hi_addr & 0x3FFFFu) >> 4), a_lo = dA | (uint64_t)((a_lo_addr & 0x3FFFFu) >> 4);
                            tc_mma_f16(acc, a_hi, w_hi, idesc, c > 0 ? 1u : 0u);
                            tc_mma_f16(acc, a_lo, w_hi, idesc, 1u);
                            tc_mma_f16(acc, a_hi, w_lo, idesc, 1u);
                            if (release_w) tc_commit(bar_wempty + 8 * s);
                            if (l == 0) tc_commit(bar_aempty + 8 * as);
                            if (c == nc - 1) tc_commit(bar_accfull + 8 * b);
                        }
                        __syncwarp();
                        if (release_w) { ++wit; sub = 0; } else { ++sub; }
                        TRACE(2, l, c);                                 // MMAs issued
                    }
                }
            }
        }
    } else {
        // ================= worker warps 0-7: layer-0 operand generation, then the epilogues of the tile's layers =================
        // A tile's phases are serially dependent anyway (its second CTA on the SM supplies the overlap), so the same eight warps do
        // both jobs: two warps per scheduler work on every phase instead of one.  wq = TMEM lane quarter = 32-row group,
        // half = which half of the K-chunk pairs (generation) / of the 32-column blocks (epilogue) this warp owns.
        const int wq = warp & 3, half = warp >> 2;
        TRACE_DECL(1);
        if (tid != 0) tr__ = nullptr;
        const int nc0 = A.Kp[0] / KC, npair = (nc0 + 1) / 2;
        const float sc = A.in_scale;
        const int rsub = lane >> 3, piece = lane & 7;
        const int ktot = A.k[0] + A.k[1] + A.k[2], kb1 = A.k[0], kb2 = A.k[0] + A.k[1];
        uint32_t abase = 0;                              // layer-0 chunks issued by this CTA in earlier tiles (operand-ring position)
        const int r = 32 * wq + lane;                    // epilogue: thread = row = TMEM lane
        const uint32_t lane_base = (uint32_t)(32 * wq) << 16;
        uint8_t* act_hi = smem + OFF_ACT;
        float* dotbuf = reinterpret_cast<float*>(smem + OFF_A);      // head partial sums (the operand ring is idle during the epilogues)
        uint32_t use[2] = {0, 0};

        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t m0 = tile * TM;
            // ---------------- generation: fp32 source rows -> split fp16 operand chunks (canonical K-major UMMA layout) ----------------
            // Line-coalesced loads: a step covers a PAIR of K chunks = 32 columns = 128 bytes of every row.  lane = (row-in-4,
            // 16-byte piece), so one load instruction reads 4 rows x one full 128-byte line; a warp owns 32 rows (8 instructions).
            // The loads of a warp's NEXT pair are issued before the current pair is converted (two pairs in flight per warp).
            // A warp's work is cut into UNITS of half a pair (rows 4j .. 4j+3 of its 8 row slots, j = unit & 1): the loads of
            // the next unit are in flight while the current one is converted (16 + 16 data registers).
            const int live_rows = (int)min((int64_t)TM, A.M - m0);
            // per-tile source descriptors: pointer to the source row of the tile's first row, and the first tile row that wraps around
            // a shared-row-block source (mod > 0); INT_MAX = no wrap inside this tile
            const float* sb[3];
            int wrap[3];
#pragma unroll
            for (int s_ = 0; s_ < 3; ++s_) {
                const int64_t smod = A.mod[s_];
                const int64_t first = smod > 0 ? m0 % smod : m0;
                sb[s_] = A.src[s_] ? A.src[s_] + first * A.ld[s_] : nullptr;
                wrap[s_] = (smod > 0 && first + TM > smod) ? (int)(smod - first) : INT_MAX;
            }
            auto load_unit = [&](int u, float4 (&v)[4]) {
                const int p = half + 2 * (u >> 1), j = u & 1;
                const int col0 = 32 * p + 4 * piece;
                const bool have1 = 2 * p + 1 < nc0;      // second chunk of the pair exists
                const bool mine = piece < 4 || have1;
                const int s_ = col0 < kb1 ? 0 : (col0 < kb2 ? 1 : 2);
                const int k_ = col0 - (s_ == 0 ? 0 : (s_ == 1 ? kb1 : kb2));
                const int ks_ = s_ == 0 ? A.k[0] : (s_ == 1 ? A.k[1] : A.k[2]);
                const bool whole = col0 + 4 <= ktot && k_ + 4 <= ks_;
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (mine && whole) {
                    const int sld = s_ == 0 ? A.ld[0] : (s_ == 1 ? A.ld[1] : A.ld[2]);
                    const float* cb = (s_ == 0 ? sb[0] : (s_ == 1 ? sb[1] : sb[2])) + k_;
                    const int wr = s_ == 0 ? wrap[0] : (s_ == 1 ? wrap[1] : wrap[2]);
                    const bool vec = ((reinterpret_cast<uintptr_t>(cb) & 15) == 0) && ((sld & 3) == 0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr_ = 32 * wq + 4 * (4 * j + i) + rsub;
                        if (rr_ < live_rows) {
                            int ro = rr_;                                       // row relative to the tile's first source row
                            if (rr_ >= wr) {                                    // rare: the tile straddles the end of a shared-row-block source
                                const int64_t smod = A.mod[s_], first = m0 % smod;
                                ro = (int)((first + rr_) % smod - first);
                            }
                            const float* qd = cb + (int64_t)ro * sld;
                            if (A.variant & 4) continue;
                            if (vec) v[i] = __ldg(reinterpret_cast<const float4*>(qd));
                            else v[i] = make_float4(__ldg(qd), __ldg(qd + 1), __ldg(qd + 2), __ldg(qd + 3));
                        }
                    }
                } else if (mine) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr_ = 32 * wq + 4 * (4 * j + i) + rsub;
                        if (rr_ < live_rows) {
                            float e[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {            // straddles two sources or the zero padding: per element
                                const int c = col0 + q;
                                e[q] = 0.f;
                                if (c < ktot) {
                                    const int su = c < kb1 ? 0 : (c < kb2 ? 1 : 2);
                                    const int ku = c - (su == 0 ? 0 : (su == 1 ? kb1 : kb2));
                                    const int64_t rr = A.mod[su] > 0 ? (m0 + rr_) % A.mod[su] : m0 + rr_;
                                    e[q] = __ldg(A.src[su] + rr * A.ld[su] + ku);
                                }
                            }
                            v[i] = make_float4(e[0], e[1], e[2], e[3]);
                        }
                    }
                }
            };
            auto put_unit = [&](int u, const float4 (&v)[4]) {
                const int p = half + 2 * (u >> 1), j = u & 1;
                const int col0 = 32 * p + 4 * piece;
                const bool have1 = 2 * p + 1 < nc0;
                const bool mine = piece < 4 || have1;
                if (A.x0img && mine) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int64_t mr = m0 + 32 * wq + 4 * (4 * j + i) + rsub;
                        const uint32_t h0 = img::pack_bf16(v[i].x, v[i].y), h1 = img::pack_bf16(v[i].z, v[i].w);
                        const uint32_t l0 = img::pack_bf16(v[i].x - img::bf16_lo_f(h0), v[i].y - img::bf16_hi_f(h0));
                        const uint32_t l1 = img::pack_bf16(v[i].z - img::bf16_lo_f(h1), v[i].w - img::bf16_hi_f(h1));
                        uint8_t* gp = A.x0img + img::piece_off(mr, col0 >> 3, A.Kp[0]) + (col0 & 7) * 2;
                        *reinterpret_cast<uint2*>(gp) = make_uint2(h0, h1);
                        *reinterpret_cast<uint2*>(gp + img::plane_bytes(A.Kp[0])) = make_uint2(l0, l1);
                    }
                }
                const uint32_t ait = abase + 2 * (uint32_t)p;
                const uint32_t st0 = ait % NSA, ph0 = (ait / NSA) & 1, st1 = (ait + 1) % NSA, ph1 = ((ait + 1) / NSA) & 1;
                if (j == 0) {
                    TRACE(10, p, 0);
                    mbar_wait_relaxed(bar_aempty + 8 * st0, ph0 ^ 1, 32);
                    if (have1) mbar_wait_relaxed(bar_aempty + 8 * st1, ph1 ^ 1, 32);
                    TRACE(11, p, 0);
                }
                if (mine) {
                    uint8_t* stage = smem + OFF_A + (piece < 4 ? st0 : st1) * A_STAGE + ((piece >> 1) & 1) * A_LBO + (piece & 1) * 8;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr_ = 32 * wq + 4 * (4 * j + i) + rsub;
                        const float a = v[i].x * sc, b = v[i].y * sc, c = v[i].z * sc, d = v[i].w * sc;
                        if (fmaxf(fmaxf(fabsf(a), fabsf(b)), fmaxf(fabsf(c), fabsf(d))) > 65000.f && A.status) atomicOr(A.status, 2);
                        const uint32_t h0 = pack_sat(a, b), h1 = pack_sat(c, d);
                        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
                        const uint32_t l0 = pack_sat(a - f0.x, b - f0.y), l1 = pack_sat(c - f1.x, d - f1.y);
                        *reinterpret_cast<uint2*>(stage + rr_ * 16) = make_uint2(h0, h1);
                        *reinterpret_cast<uint2*>(stage + A_PART + rr_ * 16) = make_uint2(l0, l1);
                    }
                }
                if (j == 1) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(bar_afull + 8 * st0);
                        if (have1) mbar_arrive(bar_afull + 8 * st1);
                    }
                    TRACE(12, p, 0);
                }
            };
            // the NEXT tile's source rows are requested into L2 a whole tile time ahead (no registers involved): this tile's loads
            // then pay an L2 hit instead of an HBM round trip
            {
                const int64_t tn = tile + gridDim.x, mrow = tn * TM + (tid & 127);
                if (tn < ntiles && mrow < A.M && !(A.variant & 1)) {
#pragma unroll
                    for (int s_ = 0; s_ < 3; ++s_) {
                        const int ks_ = A.k[s_];
                        if (ks_ > 0) {
                            const int64_t rr = A.mod[s_] > 0 ? mrow % A.mod[s_] : mrow;
                            const char* rb = reinterpret_cast<const char*>(A.src[s_] + rr * A.ld[s_]);
                            const char* re = rb + 4 * ks_;
                            for (const char* q = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(rb) & ~(uintptr_t)127) + 128 * (tid >> 7); q < re; q += 256)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
                        }
                    }
                }
            }
            {
                float4 va[4], vb[4];
                const int nu = 2 * ((npair - half + 1) / 2);          // this warp's pairs: half, half + 2, ...
                if (nu > 0) load_unit(0, va);
                for (int u = 0; u < nu; u += 2) {
                    load_unit(u + 1, vb);
                    put_unit(u, va);
                    if (u + 2 < nu) load_unit(u + 2, va);
                    put_unit(u + 1, vb);
                }
            }
            abase += (uint32_t)nc0;

            // ---------------- epilogues: thread = row = TMEM lane; this warp's half of the 32-column blocks ----------------
            const int64_t m = m0 + r;
            const bool live = m < A.M;
#pragma unroll 1
            for (int l = 0; l < nl; ++l) {
                const uint32_t b = l & 1;
                mbar_wait_relaxed(bar_accfull + 8 * b, use[b] & 1, 20);
                ++use[b];
                tc_fence_after();
                TRACE(20, l, 0);                                                    // accumulator complete
                const uint32_t taddr = tmem_base + lane_base + b * NMAX;
                const int np = A.Np[l], n = A.N[l], act = A.act[l];
                const bool last = (l == nl - 1);
                const float mul = A.mul[l], inv_next = A.inv_next[l];
                const float* bl = A.bias + l * NMAX;
                float* yout = A.Y[l];
                const int ldy = A.ldy[l];
                const bool yvec = yout && (ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(yout) & 15) == 0;
                // last layer with dense, 32-column-multiple rows: stage the fp32 tile in the (now free) activation buffer and
                // store it with whole-row coalesced writes instead of one 64-byte piece per thread
                const bool staged = last && yvec && ldy == n && (n & 31) == 0;
                float4* stg = reinterpret_cast<float4*>(smem + OFF_ACT);
                float dot = 0.f;
                const int nblk = (np + 31) >> 5, bsplit = (nblk + 1) >> 1;
                const int cbeg = half == 0 ? 0 : 32 * bsplit, cend = half == 0 ? min(np, 32 * bsplit) : np;
                if (!last) {
                    // inner layer: whole 32-column blocks (16-column tail), nothing leaves the SM except the optional training image
                    const float* addrow = (l == 0 && A.add0 && live) ? A.add0 + (A.mod_add0 > 0 ? m % A.mod_add0 : m) * A.ld_add0 : nullptr;
#pragma unroll 1
                    for (int c0 = cbeg; c0 < cend; c0 += 32) {
                        uint8_t* dst = act_hi + (c0 >> 3) * A_LBO + r * 16;
                        if (c0 + 32 <= cend)
                            epi_inner_block<32>(taddr, c0, bl, mul, inv_next, act == HNR_ACT_LRELU, A.himg[l], np, m, dst, A.status, bar_actfull + 8 * (c0 >> 5), lane,
                                                addrow, A.add0_scale);
                        else
                            epi_inner_block<16>(taddr, c0, bl, mul, inv_next, act == HNR_ACT_LRELU, A.himg[l], np, m, dst, A.status, bar_actfull + 8 * (c0 >> 5), lane,
                                                addrow, A.add0_scale);
                    }
                    if (yout) {
                        // (inference never asks for inner outputs; the layer-by-layer training cross-check does) re-read the block from
                        // TMEM -- tcgen05.ld is warp-collective: every lane executes it, only the stores depend on the row being live
#pragma unroll 1
                        for (int c0 = cbeg; c0 < cend; c0 += 16) {
                            float y[16];
                            tmem_ld16(taddr + c0, y);
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                float t = fmaf(y[i], mul, __ldg(bl + c0 + i));
                                if (addrow) t = fmaf(__ldg(addrow + c0 + i), A.add0_scale, t);
                                if (act == HNR_ACT_LRELU) t = fmaxf(t, 0.01f * t);
                                if (live && c0 + i < n) yout[m * ldy + c0 + i] = t * inv_next;
                            }
                        }
                    }
                } else {
#pragma unroll 1
                for (int c0 = cbeg; c0 < cend; c0 += 16) {
                    float y[16];
                    tmem_ld16_nowait(taddr + c0, y);
                    const float4* b4 = reinterpret_cast<const float4*>(bl + c0);
                    float4 bb[4];
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) bb[i4] = __ldg(b4 + i4);          // in flight together with the TMEM load
                    tmem_ld_wait(y);
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        y[4 * i4 + 0] = fmaf(y[4 * i4 + 0], mul, bb[i4].x);
                        y[4 * i4 + 1] = fmaf(y[4 * i4 + 1], mul, bb[i4].y);
                        y[4 * i4 + 2] = fmaf(y[4 * i4 + 2], mul, bb[i4].z);
                        y[4 * i4 + 3] = fmaf(y[4 * i4 + 3], mul, bb[i4].w);
                    }
                    if (act == HNR_ACT_LRELU) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) y[i] = fmaxf(y[i], 0.01f * y[i]);
                    } else if (act != HNR_ACT_NONE) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) y[i] = apply_act(y[i], act);
                    }
                    if (A.res && live) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < n) y[i] += __ldg(A.res + m * A.ldres + c0 + i);
                    }
                    if (A.head_w) {
                        const float4* h4 = reinterpret_cast<const float4*>(A.head_w + c0);      // zero padded to 128 by the host
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4) {
                            const float4 hh = __ldg(h4 + i4);
                            dot = fmaf(y[4 * i4 + 0], hh.x, dot); dot = fmaf(y[4 * i4 + 1], hh.y, dot);
                            dot = fmaf(y[4 * i4 + 2], hh.z, dot); dot = fmaf(y[4 * i4 + 3], hh.w, dot);
                        }
                    }
                    if (staged) {
#pragma unroll
                        for (int i4 = 0; i4 < 4; ++i4)      // 16-byte granule g of row r lives at r*(n/4) + (g ^ (r & 7)): conflict-free both ways
                            stg[r * (n >> 2) + (((c0 >> 2) + i4) ^ (r & 7))] = make_float4(y[4 * i4], y[4 * i4 + 1], y[4 * i4 + 2], y[4 * i4 + 3]);
                    } else if (yout && live) {
                        float* o = yout + m * ldy + c0;
                        if (yvec && c0 + 16 <= n) {
#pragma unroll
                            for (int i4 = 0; i4 < 4; ++i4)
                                reinterpret_cast<float4*>(o)[i4] = make_float4(y[4 * i4] * inv_next, y[4 * i4 + 1] * inv_next,
                                                                               y[4 * i4 + 2] * inv_next, y[4 * i4 + 3] * inv_next);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < n) o[i] = y[i] * inv_next;
                        }
                    }
                }
                }
                // this warp has read its part of the accumulator
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_accfree + 8 * b);
                if (staged) {
                    asm volatile("bar.sync 1, 256;" ::: "memory");                // all eight worker warps have staged their columns
                    const int gpr = n >> 2;                                        // granules per row (multiple of 8)
                    for (int row = warp; row < TM; row += 8) {
                        const int64_t mm = m0 + row;
                        if (mm < A.M)
                            for (int g = lane; g < gpr; g += 32) reinterpret_cast<float4*>(yout + mm * ldy)[g] = stg[row * gpr + (g ^ (row & 7))];
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");                // staging consumed before anybody writes the activation buffer again
                }
                if (last && A.head_w) {
                    if (half == 1) dotbuf[r] = dot;
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                    if (half == 0 && live) A.head_out[m] = apply_act(dot + dotbuf[r] + A.head_b[0], A.head_act);
                    asm volatile("bar.sync 2, 256;" ::: "memory");                // partial sums consumed before the next tile's generation overwrites the ring
                }
                TRACE(22, l, 0);                                                    // epilogue of the layer done
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 2 * NMAX);
}

}  // namespace

extern "C" int64_t hnr_chain_f16_chunk_bytes(int64_t Np) { return Np * 64; }

// profiling aid (see tc_common.cuh TRACE): device buffer that CTA 0 of the following tensor-core launches fills; NULL = off
static long long* g_trace = nullptr;
extern "C" void hnr_chain_f16_set_trace(void* buf) { g_trace = (long long*)buf; }
long long* hnr_trace_ptr() { return g_trace; }

// Fused chain of up to 4 dense layers (widths <= 128) over M rows, 3xFP16 on tcgen05 (see the header of this file).
// (hnr_chain_f16_forward / hnr_chain_f16_forward_train below share this launcher.)
// Arrays have nlayer entries.  Kp[0] = concat width padded to 16, Kp[l] = Np[l-1]; Np = N padded to 16.
// wpack / bias / mul / inv_next: built by hybridneuralrendering_b200/chain.py.  Y[l] may be NULL (inner layers) --
// at least one of Y[nlayer-1] and head_out must be given.
static int chain_f16_launch(const float* const* src, const int64_t* src_ld, const int64_t* src_k, const int64_t* src_mod,
                            float in_scale, int nlayer, const int64_t* Kp, const int64_t* N, const int64_t* Np, const int* act,
                            const void* wpack, const int64_t* w_off, const float* bias, const float* mul, const float* inv_next,
                            float* const* Y, const int64_t* ldy, const float* res, int64_t ldres, const float* head_w,
                            const float* head_b, int head_act, float* head_out, int64_t M, int32_t* status, void* x0img,
                            void* const* himg, void* stream, const float* add0 = nullptr, int64_t add0_ld = 0, int64_t add0_mod = 0,
                            float add0_scale = 0.f) {
    HNR_CHECK_ARG(nlayer >= 1 && nlayer <= MAXL, "chain_f16_forward: 1..4 layers");
    HNR_CHECK_ARG(!add0 || (nlayer >= 2 && add0_ld % 4 == 0 && add0_ld >= Np[0] && (reinterpret_cast<uintptr_t>(add0) & 15) == 0),
                  "chain_f16_forward: the layer-0 addend needs >= 2 layers and 16-byte aligned rows of at least Np[0] floats");
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
    A.add0 = add0; A.ld_add0 = (int)add0_ld; A.mod_add0 = add0_mod; A.add0_scale = add0_scale;
    A.trace = g_trace;
    { const char* e = getenv("HNR_CHAIN_VARIANT"); A.variant = e ? atoi(e) : 0; }
    A.status = status;
    A.x0img = (uint8_t*)x0img;
    for (int l = 0; l < nlayer; ++l) A.himg[l] = (himg && l < nlayer - 1) ? (uint8_t*)himg[l] : nullptr;
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

extern "C" int hnr_chain_f16_forward(const float* const* src, const int64_t* src_ld, const int64_t* src_k, const int64_t* src_mod,
                                     float in_scale, int nlayer, const int64_t* Kp, const int64_t* N, const int64_t* Np, const int* act,
                                     const void* wpack, const int64_t* w_off, const float* bias, const float* mul, const float* inv_next,
                                     float* const* Y, const int64_t* ldy, const float* res, int64_t ldres, const float* head_w,
                                     const float* head_b, int head_act, float* head_out, int64_t M, int32_t* status, void* stream) {
    return chain_f16_launch(src, src_ld, src_k, src_mod, in_scale, nlayer, Kp, N, Np, act, wpack, w_off, bias, mul, inv_next, Y, ldy, res, ldres,
                            head_w, head_b, head_act, head_out, M, status, nullptr, nullptr, stream);
}

// Training forward of a chain: same arithmetic; additionally saves the concatenated input (x0img: Kp[0] columns, kernel source
// order) and the outputs of the inner layers (himg[l], l < nlayer-1: Np[l] columns) as split images with ceil(M/128)*128 rows.
extern "C" int hnr_chain_f16_forward_train(const float* const* src, const int64_t* src_ld, const int64_t* src_k, const int64_t* src_mod,
                                           float in_scale, int nlayer, const int64_t* Kp, const int64_t* N, const int64_t* Np, const int* act,
                                           const void* wpack, const int64_t* w_off, const float* bias, const float* mul, const float* inv_next,
                                           float* const* Y, const int64_t* ldy, const float* res, int64_t ldres, const float* head_w,
                                           const float* head_b, int head_act, float* head_out, int64_t M, int32_t* status, void* x0img,
                                           void* const* himg, void* stream) {
    HNR_CHECK_ARG(x0img && himg, "chain_f16_forward_train: images required");
    return chain_f16_launch(src, src_ld, src_k, src_mod, in_scale, nlayer, Kp, N, Np, act, wpack, w_off, bias, mul, inv_next, Y, ldy, res, ldres,
                            head_w, head_b, head_act, head_out, M, status, x0img, himg, stream);
}

// The same with an addend of layer 0's pre-activation (see ChainArgs::add0): y_0 = act(x_0 W_0^T + b_0 + add0[m % add0_mod, :]).
// add0 rows must hold Np[0] floats (zero padded), 16-byte aligned; add0_scale = the chain's activation scale (chain.py ACT_SCALE).
// x0img / himg may be NULL (inference) or the training images of hnr_chain_f16_forward_train.
extern "C" int hnr_chain_f16_forward_add0(const float* const* src, const int64_t* src_ld, const int64_t* src_k, const int64_t* src_mod,
                                          float in_scale, int nlayer, const int64_t* Kp, const int64_t* N, const int64_t* Np, const int* act,
                                          const void* wpack, const int64_t* w_off, const float* bias, const float* mul, const float* inv_next,
                                          float* const* Y, const int64_t* ldy, const float* res, int64_t ldres, const float* head_w,
                                          const float* head_b, int head_act, float* head_out, int64_t M, int32_t* status, void* x0img,
                                          void* const* himg, const float* add0, int64_t add0_ld, int64_t add0_mod, float add0_scale,
                                          void* stream) {
    HNR_CHECK_ARG(add0 != nullptr, "chain_f16_forward_add0: addend required");
    return chain_f16_launch(src, src_ld, src_k, src_mod, in_scale, nlayer, Kp, N, Np, act, wpack, w_off, bias, mul, inv_next, Y, ldy, res, ldres,
                            head_w, head_b, head_act, head_out, M, status, x0img, x0img ? himg : nullptr, stream, add0, add0_ld, add0_mod, add0_scale);
}
