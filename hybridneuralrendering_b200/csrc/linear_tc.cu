// Dense layer on the 5th-gen tensor cores with fp32 accuracy (3xTF32), for the per-sample MLPs of the
// aggregator: colour-feature branch 280->128->128->128, blend-weight net 176->64->64->64(->1) batched over
// the V reference views, mix-up 90->45->45->45, colour head (reference layers:
// models/aggregators/point_aggregators.py:556-683, used at :1028-1037, :1188-1217, :1285-1334).
//
//   Y[m, n] = act( sum_k concat(A0,A1,A2)[m,k] * W[n,k] + b[n] ) (+ res[m,n])
//
// Same machinery as mlp_tc.cu (tcgen05.mma kind::tf32 x3, TMEM accumulators, cp.async.bulk weight ring,
// mbarrier pipelines) but one layer per launch with global-memory operands:
//   * persistent CTA per SM over 128-row tiles; N padded to a multiple of 16 (<= 256);
//   * 8 converter warps (thread = row x K-half) read fp32 rows one super-chunk (4 K-chunks) ahead, split them
//     into TF32 hi/lo and store the canonical UMMA operand layout;
//   * the TMEM accumulator is double buffered (2 x N columns): 4 epilogue warps drain tile t (bias,
//     activation, residual, store) while the MMA warp already accumulates tile t+1;
//   * weights come pre-split / pre-tiled from the host (linear_tc.py), one bulk copy per K-chunk.
#include <limits.h>

#include "common.cuh"
#include "hnr.h"
#define TRACE_SRC L.trace
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int TM = 128;
constexpr int KC = 8;
constexpr int NSW = 8;                       // weight ring depth
constexpr int NSA = 8;                       // activation ring depth
constexpr int A_PART = TM * KC * 4;          // 4096
constexpr int A_STAGE = 2 * A_PART;          // hi + lo
constexpr int W_STAGE_MAX = 2 * 256 * KC * 4;   // 16384
constexpr int NCONV_WARPS = 8, NEPI_WARPS = 8;   // two epilogue warps per TMEM lane quarter, alternating 16-column steps
constexpr int WARP_MMA = NCONV_WARPS + NEPI_WARPS, WARP_TMA = WARP_MMA + 1;
constexpr int NTHREADS = (WARP_TMA + 1) * 32;   // 576
constexpr int SUPER = 4;                     // K-chunks per super-chunk

constexpr int OFF_W = 0;
constexpr int OFF_A = OFF_W + NSW * W_STAGE_MAX;        // 131072
constexpr int OFF_BIAS = OFF_A + NSA * A_STAGE;         // +65536 = 196608
constexpr int OFF_BAR = OFF_BIAS + 256 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 512;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr uint32_t A_LBO = (TM / 8) * 128, SBO = 128;

struct Src {
    const float* p[3];
    int ld[3];
    int k[3];
    int64_t mod[3];
};

struct LArgs {
    Src A;
    const uint8_t* wpack;
    const float* bias;
    const float* res;
    const float* head_w;     // optional fused 1-output head: out_head[m] = head_act(sum_n y[m,n] head_w[n] + head_b[0])
    const float* head_b;
    float* out_head;
    int head_act;
    const float* gateY;      // backward-data mode: the A operand is dY * act'(gateY) (gateY = the layer's saved output)
    int ldgate, gate_act;
    float* Y;
    long long* trace;        // profiling aid (tc_common.cuh TRACE)
    int64_t M;
    int ldres, ldy;
    int K, Kp, N, Npad, act;
};

__device__ __forceinline__ float src_load(const Src& a, int64_t m, int k) {
    int s = 0;
    if (k >= a.k[0]) { k -= a.k[0]; s = 1; if (k >= a.k[1]) { k -= a.k[1]; s = 2; } }
    if (a.mod[s] > 0) m %= a.mod[s];
    return __ldg(a.p[s] + m * a.ld[s] + k);
}

// four consecutive K elements of row m starting at k0 (zero beyond K / M)
__device__ __forceinline__ float4 load4_raw(const LArgs& L, int64_t m, int k0) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m >= L.M || k0 >= L.K) return v;
    // fast path: the group lies inside one source and is 16-byte aligned
    int s = 0, kk = k0;
    if (kk >= L.A.k[0]) { kk -= L.A.k[0]; s = 1; if (kk >= L.A.k[1]) { kk -= L.A.k[1]; s = 2; } }
    if (kk + 4 <= L.A.k[s]) {
        int64_t mm = L.A.mod[s] > 0 ? m % L.A.mod[s] : m;
        const float* p = L.A.p[s] + mm * L.A.ld[s] + kk;
        if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) return __ldg(reinterpret_cast<const float4*>(p));
        v.x = __ldg(p); v.y = __ldg(p + 1); v.z = __ldg(p + 2); v.w = __ldg(p + 3);
        return v;
    }
    v.x = src_load(L.A, m, k0);
    if (k0 + 1 < L.K) v.y = src_load(L.A, m, k0 + 1);
    if (k0 + 2 < L.K) v.z = src_load(L.A, m, k0 + 2);
    if (k0 + 3 < L.K) v.w = src_load(L.A, m, k0 + 3);
    return v;
}
__device__ __forceinline__ float4 load4(const LArgs& L, int64_t m, int k0) {
    float4 v = load4_raw(L, m, k0);
    if (L.gateY && m < L.M && k0 < L.K) {
        const float* g = L.gateY + m * L.ldgate + k0;
        if (k0 + 4 <= L.K && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
            const float4 y = __ldg(reinterpret_cast<const float4*>(g));
            v.x *= act_grad_from_out(y.x, L.gate_act); v.y *= act_grad_from_out(y.y, L.gate_act);
            v.z *= act_grad_from_out(y.z, L.gate_act); v.w *= act_grad_from_out(y.w, L.gate_act);
            return v;
        }
        v.x *= act_grad_from_out(__ldg(g), L.gate_act);
        if (k0 + 1 < L.K) v.y *= act_grad_from_out(__ldg(g + 1), L.gate_act);
        if (k0 + 2 < L.K) v.z *= act_grad_from_out(__ldg(g + 2), L.gate_act);
        if (k0 + 3 < L.K) v.w *= act_grad_from_out(__ldg(g + 3), L.gate_act);
    }
    return v;
}

__global__ void __launch_bounds__(NTHREADS, 1) linear_tc_kernel(LArgs L) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    const uint32_t bar_fullW = smem_u32(bars), bar_emptyW = smem_u32(bars + NSW);
    const uint32_t bar_fullA = smem_u32(bars + 2 * NSW), bar_emptyA = smem_u32(bars + 2 * NSW + NSA);
    const uint32_t bar_accF = smem_u32(bars + 2 * NSW + 2 * NSA), bar_accE = bar_accF + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSW + 2 * NSA + 4);
    float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);

    const int64_t ntiles = (L.M + TM - 1) / TM;
    const int nchunk = L.Kp / KC;
    const uint32_t w_part = (uint32_t)L.Npad * KC * 4, w_stage = 2 * w_part;
    const uint32_t W_LBO = (uint32_t)(L.Npad / 8) * 128;
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2u * (uint32_t)L.Npad) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < NSW; ++s) { mbar_init(bar_fullW + 8 * s, 1); mbar_init(bar_emptyW + 8 * s, 1); }
        for (int s = 0; s < NSA; ++s) { mbar_init(bar_fullA + 8 * s, NCONV_WARPS); mbar_init(bar_emptyA + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_accF + 8 * b, 1); mbar_init(bar_accE + 8 * b, NEPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    for (int i = tid; i < 256; i += NTHREADS) bias_s[i] = (L.bias && i < L.N) ? L.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == WARP_TMA) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int c = 0; c < nchunk; ++c, ++it) {
                    const uint32_t s = it % NSW, ph = (it / NSW) & 1;
                    mbar_wait(bar_emptyW + 8 * s, ph ^ 1);
                    mbar_arrive_expect_tx(bar_fullW + 8 * s, w_stage);
                    bulk_g2s(smem_u32(smem + OFF_W + s * W_STAGE_MAX), L.wpack + (size_t)c * w_stage, w_stage, bar_fullW + 8 * s);
                }
        }
    } else if (warp == WARP_MMA) {
        // the whole warp walks the schedule (warp-uniform control flow keeps the tcgen05 operands in uniform registers);
        // one elected lane issues
        {
            const uint32_t idesc = idesc_tf32(TM, L.Npad);
            const uint64_t dW = umma_desc(0, W_LBO, SBO), dA = umma_desc(0, A_LBO, SBO);
            const uint32_t w_base = smem_u32(smem + OFF_W), a_base = smem_u32(smem + OFF_A);
            uint32_t it = 0, tcount = 0;
            TRACE_DECL(0);
            if (lane != 0) tr__ = nullptr;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
                const uint32_t b = tcount & 1, bph = (tcount >> 1) & 1;
                mbar_wait(bar_accE + 8 * b, bph ^ 1);            // epilogue drained this accumulator
                TRACE(3, tcount, 0);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + b * (uint32_t)L.Npad;
                for (int c = 0; c < nchunk; ++c, ++it) {
                    const uint32_t sw = it % NSW, pw = (it / NSW) & 1, sa = it % NSA, pa = (it / NSA) & 1;
                    mbar_wait(bar_fullW + 8 * sw, pw);
                    TRACE(4, c, 0);
                    mbar_wait(bar_fullA + 8 * sa, pa);
                    TRACE(1, c, 0);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t ws = w_base + sw * W_STAGE_MAX, as = a_base + sa * A_STAGE;
                        const uint64_t w_hi = dW | (uint64_t)((ws & 0x3FFFFu) >> 4), w_lo = dW | (uint64_t)(((ws + w_part) & 0x3FFFFu) >> 4);
                        const uint64_t a_hi = dA | (uint64_t)((as & 0x3FFFFu) >> 4), a_lo = dA | (uint64_t)(((as + A_PART) & 0x3FFFFu) >> 4);
                        tc_mma_tf32(d_tmem, a_hi, w_hi, idesc, c > 0 ? 1u : 0u);
                        tc_mma_tf32(d_tmem, a_lo, w_hi, idesc, 1u);
                        tc_mma_tf32(d_tmem, a_hi, w_lo, idesc, 1u);
                        tc_commit(bar_emptyW + 8 * sw);
                        tc_commit(bar_emptyA + 8 * sa);
                        if (c == nchunk - 1) tc_commit(bar_accF + 8 * b);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < NCONV_WARPS) {
        // ---------------- converters: fp32 rows -> TF32 hi/lo operand chunks ----------------
        // A warp owns 16 rows.  One load instruction covers 4 rows x one 128 B super-chunk (4 K-chunks): lane =
        // (row-in-4, 16 B piece), i.e. fully coalesced lines; DEPTH super-chunks are kept in flight per thread
        // (statically indexed register buffers), also across tile boundaries.  EVERY converter thread visits EVERY
        // ring slot use in order (a waiter that skipped a phase of an mbarrier would alias its parity).
        const int piece = lane & 7, u_of = piece >> 1, h_of = piece & 1, rsub = lane >> 3;
        const int nsc = (nchunk + SUPER - 1) / SUPER;
        const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int64_t total = my_tiles * nsc;
        constexpr int DEPTH = 3;
        float4 buf[DEPTH][4];
        // The address arithmetic of a super-chunk is resolved once per (tile, source): the 4 row offsets of this lane are
        // cached, so the common case costs one add + one 16-byte load per row (the generic load4 path -- source lookup,
        // re-use modulus, alignment tests per element group -- measured 3.5 k cycles per super-chunk and paced the kernel).
        int cs_src = -1;
        int64_t cs_tile = -1;
        const float* cs_base = nullptr;
        const float* cs_gate = nullptr;
        int cs_off[4], cs_goff[4];                      // element offsets of the 4 rows (INT_MIN = beyond M)
        bool cs_vec = false;
        auto load_super = [&](int64_t q, float4 (&dst)[4]) {
            const int64_t tile = blockIdx.x + (q / nsc) * gridDim.x;
            const int k0 = (int)(q % nsc) * SUPER * KC + piece * 4;
            int s_ = 0, kk = k0;
            if (kk >= L.A.k[0]) { kk -= L.A.k[0]; s_ = 1; if (kk >= L.A.k[1]) { kk -= L.A.k[1]; s_ = 2; } }
            const int ks_ = s_ == 0 ? L.A.k[0] : (s_ == 1 ? L.A.k[1] : L.A.k[2]);
            if (k0 + 4 <= L.K && kk + 4 <= ks_) {
                if (s_ != cs_src || tile != cs_tile) {
                    const float* sp = s_ == 0 ? L.A.p[0] : (s_ == 1 ? L.A.p[1] : L.A.p[2]);
                    const int sld = s_ == 0 ? L.A.ld[0] : (s_ == 1 ? L.A.ld[1] : L.A.ld[2]);
                    const int64_t smod = s_ == 0 ? L.A.mod[0] : (s_ == 1 ? L.A.mod[1] : L.A.mod[2]);
                    const int64_t m0 = tile * TM, first = smod > 0 ? m0 % smod : m0;
                    cs_base = sp + first * sld;
                    cs_gate = L.gateY ? L.gateY + m0 * L.ldgate : nullptr;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int r = warp * 16 + jj * 4 + rsub;
                        int64_t rr = first + r;
                        if (smod > 0 && rr >= smod) rr %= smod;
                        cs_off[jj] = (m0 + r < L.M) ? (int)((rr - first) * sld) : INT_MIN;
                        cs_goff[jj] = r * L.ldgate;
                    }
                    cs_vec = ((sld & 3) == 0) && (!L.gateY || (L.ldgate & 3) == 0);
                    cs_src = s_; cs_tile = tile;
                }
                const float* cb = cs_base + kk;
                const float* gb = cs_gate ? cs_gate + k0 : nullptr;
                const bool vec = cs_vec && ((reinterpret_cast<uintptr_t>(cb) & 15) == 0) && (!gb || (reinterpret_cast<uintptr_t>(gb) & 15) == 0);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (cs_off[jj] != INT_MIN) {
                        const float* qd = cb + cs_off[jj];
                        if (vec) v = __ldg(reinterpret_cast<const float4*>(qd));
                        else v = make_float4(__ldg(qd), __ldg(qd + 1), __ldg(qd + 2), __ldg(qd + 3));
                        if (gb) {
                            const float* gq = gb + cs_goff[jj];
                            float4 y;
                            if (vec) y = __ldg(reinterpret_cast<const float4*>(gq));
                            else y = make_float4(__ldg(gq), __ldg(gq + 1), __ldg(gq + 2), __ldg(gq + 3));
                            if (L.gate_act == HNR_ACT_LRELU) {           // the common case without the generic switch
                                v.x *= y.x > 0.f ? 1.f : 0.01f; v.y *= y.y > 0.f ? 1.f : 0.01f;
                                v.z *= y.z > 0.f ? 1.f : 0.01f; v.w *= y.w > 0.f ? 1.f : 0.01f;
                            } else {
                                v.x *= act_grad_from_out(y.x, L.gate_act); v.y *= act_grad_from_out(y.y, L.gate_act);
                                v.z *= act_grad_from_out(y.z, L.gate_act); v.w *= act_grad_from_out(y.w, L.gate_act);
                            }
                        }
                    }
                    dst[jj] = v;
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) dst[jj] = load4(L, tile * TM + warp * 16 + jj * 4 + rsub, k0);     // straddling / padded group
            }
        };
        uint32_t it = 0;
        TRACE_DECL(1);
        if (tid != 0) tr__ = nullptr;
        // a super-chunk fills up to SUPER ring slots at once: wait for all of them, let EVERY lane store its own 16-byte piece
        // (lane's chunk = u_of), then one proxy fence and one arrive per slot
        auto process = [&](int64_t q, const float4 (&src)[4]) {
            const int c0 = (int)(q % nsc) * SUPER;
            const int nvalid = min(SUPER, nchunk - c0);
            TRACE(10, c0, 0);
#pragma unroll
            for (int u = 0; u < SUPER; ++u) {
                if (u < nvalid) {
                    const uint32_t s = (it + u) % NSA, ph = ((it + u) / NSA) & 1;
                    mbar_wait(bar_emptyA + 8 * s, ph ^ 1);
                }
            }
            TRACE(11, c0, 0);
            if (u_of < nvalid) {
                uint8_t* st = smem + OFF_A + ((it + u_of) % NSA) * A_STAGE + h_of * A_LBO;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    float4 hi, lo;
                    split4(src[jj], hi, lo);
                    const int row = warp * 16 + jj * 4 + rsub;
                    *reinterpret_cast<float4*>(st + row * 16) = hi;
                    *reinterpret_cast<float4*>(st + A_PART + row * 16) = lo;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane < nvalid) mbar_arrive(bar_fullA + 8 * ((it + lane) % NSA));
            it += nvalid;
            TRACE(12, c0, 0);
        };
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
            if (d < total) load_super(d, buf[d]);
        for (int64_t q = 0; q < total; q += DEPTH) {
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) {
                if (q + d < total) {
                    process(q + d, buf[d]);
                    if (q + d + DEPTH < total) load_super(q + d + DEPTH, buf[d]);
                }
            }
        }
    } else {
        // ---------------- epilogue: TMEM -> bias / activation / residual -> global ----------------
        // Two warps serve each TMEM lane quarter and take alternate 16-column steps (with a fused head one warp takes all
        // steps of its rows: the dot product needs the whole row).  The common case -- full 16-column step, vector store,
        // LeakyReLU or no activation, no residual / head -- runs without per-element branches.
        const int ew = warp - NCONV_WARPS;
        const int q = ew & 3, half = ew >> 2;                      // TMEM lane quarter, step parity
        const int erow = q * 32 + lane;
        const bool solo = L.head_w != nullptr;                     // head: half 0 does every step, half 1 only signals
        const bool vec_ok = L.Y && (L.ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(L.Y) & 15) == 0;
        const bool simple = !L.res && !L.head_w && (L.act == HNR_ACT_LRELU || L.act == HNR_ACT_NONE);
        const float slope = L.act == HNR_ACT_LRELU ? 0.01f : 1.f;
        uint32_t tcount = 0;
        TRACE_DECL(2);
        if (ew != 0 || lane != 0) tr__ = nullptr;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
            const uint32_t b = tcount & 1, bph = (tcount >> 1) & 1;
            mbar_wait(bar_accF + 8 * b, bph);
            tc_fence_after();
            TRACE(20, tcount, 0);
            const int64_t m = tile * TM + erow;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * (uint32_t)L.Npad;
            float dot = 0.f;
            if (!(solo && half == 1)) {
                const int step0 = solo ? 0 : half, dstep = solo ? 1 : 2;
                for (int st = step0; st * 16 < L.Npad; st += dstep) {
                    const int n0 = st * 16;
                    float v[16];
                    tmem_ld16(taddr + n0, v);
                    if (m < L.M) {
                        if (simple && vec_ok && n0 + 16 <= L.N) {
                            const float4* b4 = reinterpret_cast<const float4*>(bias_s + n0);
                            float4* o = reinterpret_cast<float4*>(L.Y + m * L.ldy + n0);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float4 bb = b4[i];
                                const float t0 = v[4 * i] + bb.x, t1 = v[4 * i + 1] + bb.y, t2 = v[4 * i + 2] + bb.z, t3 = v[4 * i + 3] + bb.w;
                                o[i] = make_float4(fmaxf(t0, slope * t0), fmaxf(t1, slope * t1), fmaxf(t2, slope * t2), fmaxf(t3, slope * t3));
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int n = n0 + i;
                                if (n < L.N) {
                                    float y = apply_act(v[i] + bias_s[n], L.act);
                                    if (L.res) y += L.res[m * L.ldres + n];
                                    if (L.head_w) dot = fmaf(y, __ldg(L.head_w + n), dot);
                                    v[i] = y;
                                }
                            }
                            if (L.Y) {
                                float* o = L.Y + m * L.ldy + n0;
                                if (n0 + 16 <= L.N && vec_ok) {
#pragma unroll
                                    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                                } else {
#pragma unroll
                                    for (int i = 0; i < 16; ++i)
                                        if (n0 + i < L.N) o[i] = v[i];
                                }
                            }
                        }
                    }
                }
                if (L.head_w && m < L.M) L.out_head[m] = apply_act(dot + L.head_b[0], L.head_act);
            }
            TRACE(22, tcount, 0);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_accE + 8 * b);
        }
    }
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace

extern "C" int64_t hnr_linear_tc_packed_bytes(int64_t Npad, int64_t Kp) { return (Kp / KC) * 2 * Npad * KC * 4; }

// Y = act(concat(A) W^T + b) (+res) on tensor cores; optional fused single-output head applied to Y's rows in the
// epilogue (out_head = head_act(Y . head_w + head_b)); Y may be NULL when only the head is wanted.  wpack: image of W zero-padded to (Npad, Kp), Npad % 16 == 0,
// Npad <= 256, Kp % 8 == 0 (layout: linear_tc.py).  Same argument meaning as hnr_linear_fwd.
extern "C" int hnr_linear_tc_fwd(const float* const* a_ptr, const int64_t* a_ld, const int64_t* a_k, const int64_t* a_mod, const void* wpack,
                                 int64_t Npad, int64_t Kp, const float* bias, const float* res, int64_t ldres, float* Y, int64_t ldy,
                                 int64_t M, int64_t N, int64_t K, int act, const float* head_w, const float* head_b, int head_act,
                                 float* out_head, void* stream) {
    if (M == 0) return HNR_OK;
    HNR_CHECK_ARG(Npad % 16 == 0 && Npad >= 16 && Npad <= 256 && N <= Npad, "linear_tc_fwd: Npad must be a multiple of 16 in [16,256]");
    HNR_CHECK_ARG(Kp % KC == 0 && K <= Kp && K > 0, "linear_tc_fwd: Kp must be a multiple of 8 and >= K");
    HNR_CHECK_ARG(!(res && act != HNR_ACT_NONE), "linear_tc_fwd: residual only with act=none");
    HNR_CHECK_ARG(a_k[0] + a_k[1] + a_k[2] == K, "linear_tc_fwd: concat widths must sum to K");
    LArgs L{};
    for (int i = 0; i < 3; ++i) { L.A.p[i] = a_ptr[i]; L.A.ld[i] = (int)a_ld[i]; L.A.k[i] = (int)a_k[i]; L.A.mod[i] = a_mod ? a_mod[i] : 0; }
    HNR_CHECK_ARG(Y || (head_w && out_head), "linear_tc_fwd: no output requested");
    HNR_CHECK_ARG(!head_w || (head_b && out_head), "linear_tc_fwd: head needs head_b and out_head");
    L.head_w = head_w; L.head_b = head_b; L.out_head = out_head; L.head_act = head_act;
    L.wpack = (const uint8_t*)wpack; L.bias = bias; L.res = res; L.Y = Y; L.M = M; L.ldres = (int)ldres; L.ldy = (int)ldy;
    L.K = (int)K; L.Kp = (int)Kp; L.N = (int)N; L.Npad = (int)Npad; L.act = act;
    L.trace = hnr_trace_ptr();
    static bool configured = false;
    if (!configured) {
        HNR_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const int64_t ntiles = hnr_cdiv(M, TM);
    const int grid = (int)(ntiles < HNR_NUM_SMS ? ntiles : HNR_NUM_SMS);
    linear_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(L);
    HNR_CHECK_LAUNCH("linear_tc_fwd");
    return HNR_OK;
}

// Backward w.r.t. the layer input on tensor cores: dX (M, K) = (dY * act'(Y)) (M, N) . W (N, K).
// wpackT: image (hnr_linear_tc_packed_bytes(Kpad, Np)) of W^T zero-padded to (Kpad % 16 == 0 <= 256 output columns,
// Np % 8 == 0 >= N reduction columns); a layer with more than 256 inputs is handled by the caller as two column
// slices.  dX has row stride lddx; only its first Kout columns are written.
extern "C" int hnr_linear_tc_bwd_data(const float* dY, int64_t lddy, const float* Y, int64_t ldy, int act, const void* wpackT,
                                      int64_t Kpad, int64_t Np, float* dX, int64_t lddx, int64_t M, int64_t N, int64_t Kout,
                                      void* stream) {
    if (M == 0) return HNR_OK;
    HNR_CHECK_ARG(Kpad % 16 == 0 && Kpad >= 16 && Kpad <= 256 && Kout <= Kpad, "linear_tc_bwd_data: at most 256 output columns per call");
    HNR_CHECK_ARG(Np % KC == 0 && N <= Np && N > 0, "linear_tc_bwd_data: Np must be a multiple of 8 and >= N");
    LArgs L{};
    L.A.p[0] = dY; L.A.ld[0] = (int)lddy; L.A.k[0] = (int)N;
    L.gateY = (act == HNR_ACT_NONE) ? nullptr : Y; L.ldgate = (int)ldy; L.gate_act = act;
    L.wpack = (const uint8_t*)wpackT; L.Y = dX; L.M = M; L.ldy = (int)lddx;
    L.K = (int)N; L.Kp = (int)Np; L.N = (int)Kout; L.Npad = (int)Kpad; L.act = HNR_ACT_NONE;
    L.trace = hnr_trace_ptr();
    static bool configured = false;
    if (!configured) {
        HNR_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const int64_t ntiles = hnr_cdiv(M, TM);
    const int grid = (int)(ntiles < HNR_NUM_SMS ? ntiles : HNR_NUM_SMS);
    linear_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(L);
    HNR_CHECK_LAUNCH("linear_tc_bwd_data");
    return HNR_OK;
}
