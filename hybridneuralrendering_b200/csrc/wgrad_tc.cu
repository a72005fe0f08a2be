// Weight / bias gradient of a dense layer on the 5th-gen tensor cores (training path of SURVEY.md §8a rows A3-A4, I3, I5):
//
//     dZ = dY * act'(Y)                        (M x N)
//     dW[n, k] += sum_m dZ[m, n] * X[m, k]     (N x K),     db[n] += sum_m dZ[m, n]
//
// i.e. a GEMM whose reduction runs over the M rows (hundreds of thousands) while the output is tiny.  Mapping:
//   * UMMA "M" = 128 output rows n (one n-tile per blockIdx.y), UMMA "N" = the K input columns plus one column of
//     ones that yields db for free (padded; two halves when wider than 256), UMMA "K" = 8 rows of m per step;
//   * both operands are needed transposed (reduction index contiguous): 8 generator warps read 4(m) x 8(column)
//     patches of dY / Y / X -- one element per lane, 32-byte row segments -- apply the activation derivative, split the
//     value into TF32 hi/lo and write one 128-byte core matrix of the canonical K-major UMMA layout per patch
//     (consecutive lanes -> consecutive words: conflict free);
//   * fp32 accuracy from 3 TF32 MMAs per product (hi*hi + lo*hi + hi*lo) into one TMEM accumulator that stays resident
//     for the CTA's whole share of the rows; the CTAs of a column split the rows and add their partial dW with
//     red.global.add.f32 at the end;
//   * mbarrier ring between the generator warps and the single MMA-issuing thread.
// Replaces hnr_linear_bwd_weight (linear_simt.cu) in LinearFn.backward; same argument meaning.
#include "common.cuh"
#include "hnr.h"
#define TRACE_SRC A.trace
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int TN = 128;                 // n rows per CTA == UMMA M
constexpr int KC = 8;                   // rows of m per MMA step (tf32 K)
constexpr int NSTAGE = 3;
constexpr int NGEN_WARPS = 8;
constexpr int NTHREADS = (NGEN_WARPS + 1) * 32;
constexpr int KW_MAX = 320;             // widest supported input (+1) after padding
constexpr int A_PART = TN * KC * 4;     // 4096
constexpr int B_PART_MAX = KW_MAX * KC * 4;   // 10240
constexpr int STAGE_BYTES = 2 * A_PART + 2 * B_PART_MAX;   // 28672
constexpr int OFF_BAR = NSTAGE * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 128;
constexpr int SUPER = 4;                // K steps whose loads a generator warp keeps in flight together
constexpr int MAX_TASKS = 14;           // core-matrix patches per warp per step: ceil((32 + 2*KW_MAX/8) / 8)
static_assert((2 * TN / 8 + 2 * KW_MAX / 8 + NGEN_WARPS - 1) / NGEN_WARPS <= MAX_TASKS, "task bound");

struct WArgs {
    const float *dY, *Y;
    int lddy, ldy, act;
    const float* x[3];
    int ld[3], k[3];
    int64_t mod[3];
    float *dW, *db;
    long long* trace;           // profiling aid (tc_common.cuh TRACE)
    int64_t M;
    int N, K, KWp, half;        // KWp: padded width of [X | 1]; half = KWp when one MMA group, else KWp / 2
};

template <bool HAS_MOD>
__global__ void __launch_bounds__(NTHREADS, 1) wgrad_tc_kernel(const __grid_constant__ WArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t nsteps = (A.M + KC - 1) / KC, nsuper = (nsteps + SUPER - 1) / SUPER;
    const int64_t my_super = (int64_t)blockIdx.x < nsuper ? (nsuper - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (my_super == 0) return;                                 // uniform for the CTA: nothing allocated yet
    // K steps of this CTA: every super-step holds SUPER steps except the globally last one
    const int64_t last_super = blockIdx.x + (my_super - 1) * gridDim.x;
    const int64_t my_steps = (my_super - 1) * SUPER + (last_super == nsuper - 1 ? nsteps - last_super * SUPER : SUPER);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * NSTAGE, bar_acc = bar_empty + 8 * NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 1);
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)A.KWp) tmem_cols <<= 1;
    const int n0 = blockIdx.y * TN;
    const uint32_t b_part = (uint32_t)A.KWp * KC * 4, B_LBO = (uint32_t)(A.KWp / 8) * 128;
    constexpr uint32_t A_LBO = (TN / 8) * 128, SBO = 128;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full + 8 * s, NGEN_WARPS); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NGEN_WARPS) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == NGEN_WARPS) {
        // ================= MMA issuer: warp-uniform schedule, one elected lane issues =================
        {
            const uint32_t idesc = idesc_tf32(TN, A.half);
            const int ngroup = A.KWp / A.half;                 // 1 or 2
            const uint64_t dA = umma_desc(0, A_LBO, SBO), dB = umma_desc(0, B_LBO, SBO);
            const uint32_t s_base = smem_u32(smem);
            TRACE_DECL(0);
            if (lane != 0 || blockIdx.y != 0) tr__ = nullptr;
            for (int64_t it = 0; it < my_steps; ++it) {
                const uint32_t s = (uint32_t)(it % NSTAGE), ph = (uint32_t)((it / NSTAGE) & 1);
                mbar_wait(bar_full + 8 * s, ph);
                TRACE(1, it & 0xffff, 0);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t st = s_base + s * STAGE_BYTES;
                    const uint64_t a_hi = dA | (uint64_t)((st & 0x3FFFFu) >> 4), a_lo = dA | (uint64_t)(((st + A_PART) & 0x3FFFFu) >> 4);
                    for (int g = 0; g < ngroup; ++g) {
                        const uint32_t boff = st + 2 * A_PART + (uint32_t)g * (uint32_t)(A.half / 8) * 128;
                        const uint64_t b_hi = dB | (uint64_t)((boff & 0x3FFFFu) >> 4), b_lo = dB | (uint64_t)(((boff + b_part) & 0x3FFFFu) >> 4);
                        const uint32_t d = tmem_base + (uint32_t)g * (uint32_t)A.half;
                        tc_mma_tf32(d, a_hi, b_hi, idesc, it > 0 ? 1u : 0u);
                        tc_mma_tf32(d, a_lo, b_hi, idesc, 1u);
                        tc_mma_tf32(d, a_hi, b_lo, idesc, 1u);
                    }
                    tc_commit(bar_empty + 8 * s);
                    if (it == my_steps - 1) tc_commit(bar_acc);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= generators: transposing loads -> TF32 hi/lo core matrices =================
        // task = warp + 8t.  Tasks t < NA are A-side (dY) patches, the rest B-side (X | 1).  Both task counts are even, so the
        // K-half of a patch is h = warp & 1 and its 8-column group g = (warp >> 1) + 4t (A) / (warp >> 1) + 4(t - NA) (B).
        // Everything that depends only on the patch -- column pointer, row stride, destination -- is resolved once; a step
        // costs one address + one load per patch, then one split + two stores.
        const int j = lane >> 2, mi = lane & 3;                // patch element: column 8g + j, row 4h + mi
        constexpr int NA = 2 * (TN / 8) / NGEN_WARPS;
        const int ntask_a = 2 * (TN / 8), ntask = ntask_a + 2 * (A.KWp / 8);
        const int h = warp & 1, crow = 4 * h + mi;
        const float* cptr[MAX_TASKS];                          // column base (row 0); nullptr = constant column
        const float* gptr[NA];                                 // saved-output column for the activation derivative
        int cld[MAX_TASKS];                                    // row stride
        int csrc[MAX_TASKS];                                   // source index (row re-use modulus lookup), HAS_MOD only
        uint32_t ones_mask = 0;                                // bit t: the patch column is the column of ones (-> db)
        const uint32_t dbase_a = (uint32_t)h * A_LBO + (uint32_t)(warp >> 1) * 128u + (uint32_t)lane * 4u;
        const uint32_t dbase_b = 2u * A_PART + (uint32_t)h * B_LBO + (uint32_t)(warp >> 1) * 128u + (uint32_t)lane * 4u;
#pragma unroll
        for (int t = 0; t < MAX_TASKS; ++t) {
            cptr[t] = nullptr; cld[t] = 0; csrc[t] = 0;
            if (t < NA) gptr[t] = nullptr;
            if (warp + t * NGEN_WARPS < ntask) {
                const int col = 8 * ((warp >> 1) + 4 * (t < NA ? t : t - NA)) + j;
                if (t < NA) {
                    const int n = n0 + col;
                    if (n < A.N) {
                        cptr[t] = A.dY + n; cld[t] = A.lddy;
                        if (A.act != HNR_ACT_NONE) gptr[t] = A.Y + n;
                    }
                } else if (col < A.K) {
                    int s_ = 0, k = col;
                    if (k >= A.k[0]) { k -= A.k[0]; s_ = 1; if (k >= A.k[1]) { k -= A.k[1]; s_ = 2; } }
                    cptr[t] = A.x[s_] + k; cld[t] = A.ld[s_]; csrc[t] = s_;
                } else if (col == A.K) {
                    ones_mask |= 1u << t;
                }
            }
        }
        int64_t step = 0;
        TRACE_DECL(1);
        if (tid != 0 || blockIdx.y != 0) tr__ = nullptr;
        for (int64_t sit = 0; sit < my_super; ++sit) {
            TRACE(10, sit & 0xffff, 0);
            const int64_t ms = (blockIdx.x + sit * gridDim.x) * (int64_t)(SUPER * KC);
            float val[SUPER][MAX_TASKS];
            float gate[SUPER][NA];
            const bool full = ms + SUPER * KC <= A.M;         // every row of the super-step exists: no per-row tests
#pragma unroll
            for (int q = 0; q < SUPER; ++q) {
                const int64_t m = ms + q * KC + crow;
                if (full && !HAS_MOD && m < (1ll << 31)) {
                    // fast path (trace-guided: the generic path spent ~75 cycles per load on 64-bit index arithmetic and row tests):
                    // one 32-bit multiply + one wide add per load
                    const uint32_t m32 = (uint32_t)m;
#pragma unroll
                    for (int t = 0; t < MAX_TASKS; ++t) {
                        val[q][t] = cptr[t] ? __ldg(cptr[t] + (size_t)m32 * (uint32_t)cld[t]) : (((ones_mask >> t) & 1u) ? 1.f : 0.f);
                        if (t < NA) gate[q][t] = gptr[t] ? __ldg(gptr[t] + (size_t)m32 * (uint32_t)A.ldy) : 1.f;
                    }
                } else {
                    const bool inr = m < A.M;
                    int64_t mm[3] = {m, m, m};
                    if (HAS_MOD) {
#pragma unroll
                        for (int s_ = 0; s_ < 3; ++s_)
                            if (A.mod[s_] > 0) mm[s_] = m % A.mod[s_];
                    }
#pragma unroll
                    for (int t = 0; t < MAX_TASKS; ++t) {
                        float v = 0.f;
                        if (inr) {
                            if (cptr[t]) {
                                const int64_t row = (HAS_MOD && t >= NA) ? (csrc[t] == 0 ? mm[0] : (csrc[t] == 1 ? mm[1] : mm[2])) : m;
                                v = __ldg(cptr[t] + row * cld[t]);
                            } else if ((ones_mask >> t) & 1u) {
                                v = 1.f;
                            }
                        }
                        val[q][t] = v;
                        if (t < NA) gate[q][t] = (inr && gptr[t]) ? __ldg(gptr[t] + m * A.ldy) : 1.f;
                    }
                }
            }
            TRACE(11, sit & 0xffff, 0);
            const bool lrelu = A.act == HNR_ACT_LRELU;
#pragma unroll
            for (int q = 0; q < SUPER; ++q) {
                if (ms + q * KC < A.M) {
                    const uint32_t s = (uint32_t)(step % NSTAGE), ph = (uint32_t)((step / NSTAGE) & 1);
                    ++step;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    uint8_t* st = smem + s * STAGE_BYTES;
#pragma unroll
                    for (int t = 0; t < MAX_TASKS; ++t) {
                        if (warp + t * NGEN_WARPS < ntask) {
                            float v = val[q][t];
                            if (t < NA && gptr[t]) v *= lrelu ? (gate[q][t] > 0.f ? 1.f : 0.01f) : act_grad_from_out(gate[q][t], A.act);
                            const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u), lo = v - hi;
                            const uint32_t off = t < NA ? dbase_a + (uint32_t)t * 512u : dbase_b + (uint32_t)(t - NA) * 512u;
                            *reinterpret_cast<float*>(st + off) = hi;
                            *reinterpret_cast<float*>(st + off + (t < NA ? (uint32_t)A_PART : b_part)) = lo;
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + 8 * s);
                    TRACE(12, sit & 0xffff, q);
                }
            }
        }
        // ================= epilogue (warps 0-3): TMEM partial -> red.global.add =================
        if (warp < 4) {
            mbar_wait(bar_acc, 0);
            tc_fence_after();
            const int n = n0 + 32 * warp + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
            for (int c0 = 0; c0 < A.KWp; c0 += 16) {
                float v[16];
                tmem_ld16(taddr + c0, v);
                if (n < A.N) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = c0 + i;
                        if (c < A.K) atomicAdd(A.dW + (int64_t)n * A.K + c, v[i]);
                        else if (c == A.K && A.db) atomicAdd(A.db + n, v[i]);
                    }
                }
            }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == NGEN_WARPS) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace

// dW (N,K) += (dY * act'(Y))^T . concat(X), db (N) += column sums of dY * act'(Y); dW / db must be initialised by the
// caller (normally zeros).  3xTF32 on tcgen05; K + 1 <= 320 after padding.  Same argument meaning as hnr_linear_bwd_weight.
extern "C" int hnr_linear_tc_bwd_weight(const float* dY, int64_t lddy, const float* Y, int64_t ldy, const float* const* a_ptr,
                                        const int64_t* a_ld, const int64_t* a_k, const int64_t* a_mod, float* dW, float* db, int64_t M,
                                        int64_t N, int64_t K, int act, void* stream) {
    if (M == 0) return HNR_OK;
    HNR_CHECK_ARG(N > 0 && K > 0 && a_k[0] + a_k[1] + a_k[2] == K, "linear_tc_bwd_weight: concat widths must sum to K");
    int64_t kw = (K + 1 + 15) / 16 * 16, half = kw;
    if (kw > 256) { kw = (K + 1 + 31) / 32 * 32; half = kw / 2; }
    HNR_CHECK_ARG(kw <= KW_MAX, "linear_tc_bwd_weight: layer input too wide (K + 1 must be <= 320)");
    WArgs A{};
    A.dY = dY; A.Y = Y; A.lddy = (int)lddy; A.ldy = (int)ldy; A.act = act;
    for (int i = 0; i < 3; ++i) { A.x[i] = a_ptr[i]; A.ld[i] = (int)a_ld[i]; A.k[i] = (int)a_k[i]; A.mod[i] = a_mod ? a_mod[i] : 0; }
    A.trace = hnr_trace_ptr();
    A.dW = dW; A.db = db; A.M = M; A.N = (int)N; A.K = (int)K; A.KWp = (int)kw; A.half = (int)half;
    static bool configured = false;
    if (!configured) {
        HNR_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        HNR_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const bool has_mod = A.mod[0] > 0 || A.mod[1] > 0 || A.mod[2] > 0;
    const int ny = (int)hnr_cdiv(N, TN);
    const int64_t nsteps = hnr_cdiv(M, KC);
    const int64_t nsuper = hnr_cdiv(nsteps, SUPER);
    int gx = HNR_NUM_SMS / ny;
    if (gx < 1) gx = 1;
    if (nsuper < gx) gx = (int)nsuper;
    if (has_mod) wgrad_tc_kernel<true><<<dim3((unsigned)gx, (unsigned)ny), NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(A);
    else wgrad_tc_kernel<false><<<dim3((unsigned)gx, (unsigned)ny), NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(A);
    HNR_CHECK_LAUNCH("linear_tc_bwd_weight");
    return HNR_OK;
}
