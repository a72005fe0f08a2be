// Weight / bias gradients of the fused MLPs straight from split images (img_common.cuh) on the 5th-gen tensor cores:
//
//     dW[n, k] = sum_m dZ[m, n] * X[m, k]          db[n] = sum_m dZ[m, n]            (reduction over the M rows)
//
// Both operands lie in HBM exactly as the tensor core wants them for this product: the reduction index m is the ROW of the
// images, so a 32-row slab plane is a canonical MN-major UMMA operand (SBO = 512 B between 8-column groups, LBO = 128 B
// between 8-row groups).  No thread ever touches an operand element:
//   * one bulk-copy warp streams slabs (A = a 128-column half of dZ, B = all columns of X [+ an extra image, e.g. the 16-wide
//     block3 extras]) through a 3-stage ring;
//   * one MMA warp issues, per 16 rows, kind::f16 (bf16) MMAs with M = 128 (n), N = up to 256 (+ a second block for inputs
//     wider than 256), 3 per product (hi*hi + lo*hi + hi*lo), plus two MMAs against a constant block of ones that yield db;
//     the accumulator (128 lanes x <= 304 fp32 columns) stays in TMEM for the CTA's whole share of the rows;
//   * 4 epilogue warps add the CTA's partial sums to the output with red.global.add at the very end.
// Work split: every CTA owns (job = layer, half of the 256 output rows, a residue class of the slabs).  The two halves of a
// (job, class) run on neighbouring CTAs and read the same X slabs at about the same time (second read served by L2).
// Replaces the per-layer wgrad_tc.cu launches (transposing generator warps over fp32 rows) for the per-neighbour MLP.
#include <stdlib.h>

#include "common.cuh"
#include "hnr.h"
#define TRACE_SRC ((long long*)nullptr)
#include "tc_common.cuh"
#include "img_common.cuh"

namespace {
using namespace tc;

constexpr int MAXJOB = 4;
// 3 stages x 55 KB keep ~160 KB of bulk copies in flight per SM (HBM needs ~45 KB to cover its latency at 44 GB/s per SM) and leave
// ~60 KB of shared memory to the small-grid kernels of the image-branch tail, which run on a side stream CONCURRENTLY with this
// kernel (ops.defer_weight_gradients): with 4 stages (222 KB) nothing else could be resident on the SM and the two lanes serialised
constexpr int NSTAGE = 3;
constexpr int A_BYTES = 2 * 16 * 512;           // hi + lo plane of a 128-column half: 16384
constexpr int B_MAX_COLS = 288 + 16;            // widest [X | extra]
constexpr int B_PLANE_MAX = B_MAX_COLS / 8 * 512;   // 19456
constexpr int STAGE_BYTES = A_BYTES + 2 * B_PLANE_MAX;     // 55296
constexpr int OFF_ONES = NSTAGE * STAGE_BYTES;  // 221184: 16 columns x 32 rows of bf16 1.0 (2 groups x 512 B)
constexpr int OFF_BAR = OFF_ONES + 1024;
constexpr int SMEM_BYTES = OFF_BAR + (2 * NSTAGE + 1) * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int NTHREADS = 192;                   // warp 0 bulk copy, warp 1 MMA, warps 2-5 epilogue (TMEM lane quarters 2,3,0,1)
constexpr uint32_t MN_SBO = 512, MN_LBO = 128;

struct WJob {
    const uint8_t* a;      // dZ image, ca columns (the layer's padded output width)
    const uint8_t* b;      // X image, cb columns
    const uint8_t* e;      // optional extra image, ce columns (NULL: none)
    float* dw;             // (n_rows, k_cols) fp32 row-major, += : operand column c lands in dW column colmap[c] (identity when NULL; < 0 or
    float* db;             //   >= k_cols: dropped -- padding columns); db (n_rows) += column sums of dZ
    const int32_t* colmap;
    int ca, cb, ce, n_rows, k_cols;
    int nhalf;             // ceil(ca / 128): 128-row blocks of the output, one per CTA of a group
    int cta_begin;         // first CTA of this job; the job owns CTAs [cta_begin, next job's cta_begin), a multiple of nhalf
    int64_t nslab;         // 32-row slabs of this job's images
};
struct WArgs {
    WJob job[MAXJOB];
    int njob, ncta;
    uint32_t lbo, sbo;     // descriptor strides (bring-up switch HNR_WG_SWAP exchanges them)
};

__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D = f32, A = B = bf16, both MN-major, M = 128
__host__ __device__ constexpr uint32_t idesc_mn(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(NTHREADS, 1) wgrad_img_kernel(const __grid_constant__ WArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- which job / half / slab class is mine (uniform for the CTA)
    int ji = 0;
#pragma unroll
    for (int q = 1; q < MAXJOB; ++q)
        if (q < A.njob && (int)blockIdx.x >= A.job[q].cta_begin) ji = q;
    const WJob& J = A.job[ji];
    const int cta_end = ji + 1 < A.njob ? A.job[ji + 1].cta_begin : A.ncta;
    const int npair = (cta_end - J.cta_begin) / J.nhalf;
    const int local = (int)blockIdx.x - J.cta_begin;
    const int half = local % J.nhalf, cls = local / J.nhalf;
    const int64_t my_slabs = cls < J.nslab ? (J.nslab - cls + npair - 1) / npair : 0;
    if (my_slabs == 0) return;                                  // uniform: nothing allocated yet
    const int ct = J.cb + J.ce;                                 // operand columns; the ones block follows
    const uint32_t b_plane = (uint32_t)ct / 8 * 512;
    const int hw = min(128, J.ca - 128 * half);                 // output rows (= dZ columns) of this CTA
    const uint32_t a_plane = (uint32_t)hw / 8 * 512;            // bytes of one A plane actually loaded (<= 8192)

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * NSTAGE, bar_acc = bar_empty + 8 * NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 1);
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 256; i += NTHREADS) reinterpret_cast<uint32_t*>(smem + OFF_ONES)[i] = 0x3F803F80u;     // bf16 1.0 pairs
    if (hw < 128) {
        // a layer narrower than the 128-row MMA: the column groups the copies never touch read as zeros
        for (int s = 0; s < NSTAGE; ++s)
            for (int part = 0; part < 2; ++part)
                for (uint32_t o = a_plane + tid * 16; o < 8192; o += NTHREADS * 16)
                    *reinterpret_cast<uint4*>(smem + s * STAGE_BYTES + part * 8192 + o) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ================= bulk-copy producer =================
            const uint32_t stage_tx = 2 * a_plane + 2 * b_plane;
            for (int64_t it = 0; it < my_slabs; ++it) {
                const int64_t slab = cls + it * npair;
                const uint32_t s = (uint32_t)(it % NSTAGE), ph = (uint32_t)((it / NSTAGE) & 1);
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                const uint32_t full = bar_full + 8 * s;
                mbar_arrive_expect_tx(full, stage_tx);
                const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
                const uint8_t* ap = J.a + slab * img::slab_bytes(J.ca) + (int64_t)half * 8192;
                bulk_g2s(st, ap, a_plane, full);                                        // A hi: the column groups of my half
                bulk_g2s(st + 8192, ap + img::plane_bytes(J.ca), a_plane, full);        // A lo
                const uint8_t* bp = J.b + slab * img::slab_bytes(J.cb);
                const uint32_t bmain = (uint32_t)J.cb / 8 * 512;
                bulk_g2s(st + A_BYTES, bp, bmain, full);                                // B hi
                bulk_g2s(st + A_BYTES + b_plane, bp + img::plane_bytes(J.cb), bmain, full);   // B lo
                if (J.ce) {
                    const uint8_t* ep = J.e + slab * img::slab_bytes(J.ce);
                    const uint32_t bex = (uint32_t)J.ce / 8 * 512;
                    bulk_g2s(st + A_BYTES + bmain, ep, bex, full);
                    bulk_g2s(st + A_BYTES + b_plane + bmain, ep + img::plane_bytes(J.ce), bex, full);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: warp-uniform schedule, one elected lane issues =================
        const int n1 = ct > 256 ? 256 : ct, n2 = ct - n1;           // first / second column block
        const uint32_t id1 = idesc_mn(n1), id2 = idesc_mn(n2 > 0 ? n2 : 16), id_ones = idesc_mn(16);
        const uint64_t dsc = umma_desc(0, A.lbo, A.sbo);
        const uint32_t s_base = smem_u32(smem), ones = smem_u32(smem + OFF_ONES);
        const uint64_t d_ones = dsc | (uint64_t)((ones & 0x3FFFFu) >> 4);
        for (int64_t it = 0; it < my_slabs; ++it) {
            const uint32_t s = (uint32_t)(it % NSTAGE), ph = (uint32_t)((it / NSTAGE) & 1);
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t st = s_base + s * STAGE_BYTES;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {                    // 2 x 16 rows of the slab; 8 rows = 128 B along K
                    const uint32_t koff = (uint32_t)ks * 256u;
                    const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                    const uint64_t a_hi = dsc | (uint64_t)(((st + koff) & 0x3FFFFu) >> 4), a_lo = dsc | (uint64_t)(((st + 8192 + koff) & 0x3FFFFu) >> 4);
                    const uint32_t bh = st + A_BYTES + koff, bl = bh + b_plane;
                    const uint64_t b_hi = dsc | (uint64_t)((bh & 0x3FFFFu) >> 4), b_lo = dsc | (uint64_t)((bl & 0x3FFFFu) >> 4);
                    tc_mma_bf16(tmem_base, a_hi, b_hi, id1, acc);
                    tc_mma_bf16(tmem_base, a_lo, b_hi, id1, 1u);
                    tc_mma_bf16(tmem_base, a_hi, b_lo, id1, 1u);
                    if (n2 > 0) {
                        const uint32_t o2 = (uint32_t)(n1 / 8) * 512u;
                        const uint64_t b2_hi = dsc | (uint64_t)(((bh + o2) & 0x3FFFFu) >> 4), b2_lo = dsc | (uint64_t)(((bl + o2) & 0x3FFFFu) >> 4);
                        tc_mma_bf16(tmem_base + 256, a_hi, b2_hi, id2, acc);
                        tc_mma_bf16(tmem_base + 256, a_lo, b2_hi, id2, 1u);
                        tc_mma_bf16(tmem_base + 256, a_hi, b2_lo, id2, 1u);
                    }
                    const uint64_t o_k = d_ones + (uint64_t)(koff >> 4);
                    tc_mma_bf16(tmem_base + (uint32_t)ct, a_hi, o_k, id_ones, acc);
                    tc_mma_bf16(tmem_base + (uint32_t)ct, a_lo, o_k, id_ones, 1u);
                }
                tc_commit(bar_empty + 8 * s);
                if (it == my_slabs - 1) tc_commit(bar_acc);
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue: TMEM partial -> red.global.add =================
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const int q = warp & 3;                                     // TMEM lane quarter this warp may read
        const int n = half * 128 + q * 32 + lane;
        const bool live = q * 32 + lane < hw && n < J.n_rows;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        float* o = J.dw + (int64_t)n * J.k_cols;
        for (int c0 = 0; c0 < ct + 16; c0 += 16) {
            float v[16];
            tmem_ld16(taddr + c0, v);
            if (!live) continue;
            if (c0 < ct) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int col = J.colmap ? J.colmap[c0 + i] : c0 + i;
                    if (col >= 0 && col < J.k_cols) atomicAdd(o + col, v[i]);
                }
            } else if (J.db) {
                atomicAdd(J.db + n, v[0]);                          // every column of the ones block holds db
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// Weight + bias gradients of up to 4 layers in one launch.  Per job i: a[i] = split image of dZ (ca[i] columns = the layer's
// padded output width, multiple of 16, <= 256), b[i] = split image of the layer input (cb[i] columns, multiple of 16, <= 288),
// e[i] = optional extra input image (ce[i] columns, 0 or 16).  Results are ACCUMULATED (zero them first) straight into the
// parameter-shaped gradients: dw[i] (n_rows[i], k_cols[i]) row-major -- operand column c goes to column colmap[i][c] (device int32
// array of cb+ce entries; NULL = identity; entries < 0 or >= k_cols are padding and dropped) -- and db[i] (n_rows[i]) (may be NULL).
// rows_pad[i] % 128 == 0 rows per image; padding rows of the dZ images must be zero and those of the input images finite.
extern "C" int hnr_wgrad_img_jobs(int njob, const void* const* a, const int64_t* ca, const void* const* b, const void* const* e,
                                  const int64_t* cb, const int64_t* ce, float* const* dw, float* const* db, const int32_t* const* colmap,
                                  const int64_t* n_rows, const int64_t* k_cols, const int64_t* rows_pad, void* stream) {
    HNR_CHECK_ARG(njob >= 1 && njob <= MAXJOB, "wgrad_img: 1..4 jobs");
    WArgs A{};
    A.njob = njob;
    // CTA groups per job proportional to the bytes the job streams
    double cost[MAXJOB], total = 0;
    int nhalf[MAXJOB];
    for (int i = 0; i < njob; ++i) {
        HNR_CHECK_ARG(rows_pad[i] % 128 == 0, "wgrad_img: rows_pad must be a multiple of 128");
        HNR_CHECK_ARG(ca[i] >= 16 && ca[i] % 16 == 0 && ca[i] <= 256, "wgrad_img: dZ width");
        HNR_CHECK_ARG(cb[i] > 0 && cb[i] % 16 == 0 && (ce[i] == 0 || ce[i] == 16) && cb[i] + ce[i] <= B_MAX_COLS, "wgrad_img: operand widths");
        HNR_CHECK_ARG(n_rows[i] > 0 && n_rows[i] <= ca[i] && k_cols[i] > 0 && dw[i], "wgrad_img: gradient shape");
        nhalf[i] = (int)((ca[i] + 127) / 128);
        cost[i] = (double)rows_pad[i] * (double)(ca[i] + nhalf[i] * (cb[i] + ce[i]));
        total += cost[i];
    }
    if (total == 0) return HNR_OK;
    int groups[MAXJOB], used = 0;
    for (int i = 0; i < njob; ++i) {
        int g = (int)((double)HNR_NUM_SMS * cost[i] / total / nhalf[i]);
        const int64_t nslab = rows_pad[i] / img::SLAB;
        if (g < 1) g = 1;
        if (g > nslab) g = (int)(nslab > 0 ? nslab : 1);
        groups[i] = g;
        used += g * nhalf[i];
    }
    for (int guard = 0; used > HNR_NUM_SMS && guard < 1000; ++guard)            // rounding up the small jobs may overshoot
        for (int i = 0; i < njob && used > HNR_NUM_SMS; ++i)
            if (groups[i] > 1) { --groups[i]; used -= nhalf[i]; }
    int cta = 0;
    for (int i = 0; i < njob; ++i) {
        WJob& J = A.job[i];
        J.a = (const uint8_t*)a[i]; J.b = (const uint8_t*)b[i]; J.e = (e && ce[i]) ? (const uint8_t*)e[i] : nullptr;
        J.ca = (int)ca[i]; J.cb = (int)cb[i]; J.ce = J.e ? (int)ce[i] : 0; J.dw = dw[i]; J.db = db ? db[i] : nullptr;
        J.colmap = colmap ? colmap[i] : nullptr; J.n_rows = (int)n_rows[i]; J.k_cols = (int)k_cols[i]; J.cta_begin = cta;
        J.nhalf = nhalf[i]; J.nslab = rows_pad[i] / img::SLAB;
        cta += groups[i] * nhalf[i];
    }
    A.ncta = cta;
    const char* sw = getenv("HNR_WG_SWAP");
    const bool swap = sw && sw[0] == '1';
    A.lbo = swap ? MN_SBO : MN_LBO;
    A.sbo = swap ? MN_LBO : MN_SBO;
    static bool configured = false;
    if (!configured) {
        HNR_CUDA(cudaFuncSetAttribute(wgrad_img_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    wgrad_img_kernel<<<cta, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(A);
    HNR_CHECK_LAUNCH("wgrad_img");
    return HNR_OK;
}
