// Fused DATA-GRADIENT chain of the per-neighbour MLP on the 5th-gen tensor cores (training backward of SURVEY.md §8a row A3;
// what autograd computes over models/aggregators/point_aggregators.py:921-972, :1002-1026 of the reference):
//
//     dZ_3 (given)  --W4-->  dZ_2 = (dZ_3 W4) * act'(H_2)  --W3[:, :256]-->  dZ_1  --W2-->  dZ_0  --W1[:, :224]-->  dX0
//
// dZ_l is the gradient w.r.t. the PRE-activation of layer l (act = LeakyReLU 0.01, so act'(H) is read off the sign of the
// saved output).  One persistent CTA per SM walks tiles of 128 neighbour rows; the gradient tile never leaves the SM between
// the four layers -- the skeleton is the forward kernel's (nbr_mlp_f16.cu) run backwards:
//   * operands are split bf16 hi/lo (fp32's exponent range: no data-dependent scale for gradients), three
//     tcgen05.mma.kind::f16 per product (hi*hi + lo*hi + hi*lo, ~16 mantissa bits), fp32 accumulation in TMEM;
//   * two TMEM accumulators alternate between layers; the epilogue of layer l gates the accumulator with act'(H_{l-1}),
//     splits it and writes it (a) into shared memory as the next layer's A operand (canonical K-major UMMA layout,
//     released to the MMA warp per 32 columns) and (b) into HBM as a split image (img_common.cuh) -- the operand the
//     weight-gradient kernel (wgrad_img.cu) reads later with bulk copies;
//   * the first layer's operand (dZ_3 image, written by the density-head / K-sum backward) and the activation signs (hi
//     planes of the forward's saved images) are streamed by four loader warps, four 16-byte pieces per row and item,
//     FOUR items in flight per thread, into an operand ring / a gate ring; W^T chunk images arrive by cp.async.bulk;
//   * warp roles: 2 x 4 epilogue warps (alternate 32-column blocks of the same TMEM lanes), 4 loader warps, 1 MMA warp
//     (one elected lane issues), 1 bulk-copy warp; registers re-balanced with setmaxnreg.
// Not computed here: the 7 "extras" columns of layer 2's input gradient (hnr_dz_extras_bwd below, an HBM-bound pass over
// the dZ_2 image) and every weight gradient (wgrad_img.cu).
#include "common.cuh"
#include "hnr.h"
#define TRACE_SRC ((long long*)nullptr)
#include "tc_common.cuh"
#include "img_common.cuh"

namespace {
using namespace tc;

constexpr int TM = 128;                 // rows per tile == UMMA M
constexpr int HID = 256;                // layer width == UMMA N == image columns
constexpr int KC = 16;                  // K elements per chunk == one MMA K-step
constexpr int NSW = 3, NSA = 3, NSG = 3;      // weight / first-layer operand / gate ring stages
constexpr int W_PART = HID * KC * 2;    // 8192 B
constexpr int W_STAGE = 2 * W_PART;     // 16384
constexpr int A_PART = TM * KC * 2;     // 4096 B
constexpr int A_STAGE = 2 * A_PART;     // 8192
constexpr int G_STAGE = TM * 32 * 2;    // 8192 B: hi plane of a 128 x 32 block, [4 groups][128 rows][16 B]
constexpr int ACT_PART = TM * HID * 2;  // 65536 B
constexpr int NLAYER = 4, NC = HID / KC;      // 16 chunks per layer
constexpr int NITEM = NC + 3 * 8;       // loader items per tile: 16 operand chunks + 3 x 8 gate blocks
constexpr int DEPTH = 4;                // loader items in flight per thread
constexpr int NEPI = 256, NTHREADS = 512;     // warps 0-7 epilogue, 8-11 loaders, 12 MMA, 13 bulk copy, 14-15 idle
static_assert(NITEM % DEPTH == 0, "the loader's register queue must line up across tiles");

constexpr int OFF_W = 0;
constexpr int OFF_A = OFF_W + NSW * W_STAGE;              // 49152
constexpr int OFF_G = OFF_A + NSA * A_STAGE;              // 73728
constexpr int OFF_ACT = OFF_G + NSG * G_STAGE;            // 98304
constexpr int OFF_BAR = OFF_ACT + 2 * ACT_PART;           // 229376
constexpr int NBAR = 2 * NSW + 2 * NSA + 2 * NSG + 8 + 2 + 2;
constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// D = f32, A = B = bf16, both K-major, M = 128, N = 256
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr uint32_t A_LBO = (TM / 8) * 128, W_LBO = (HID / 8) * 128, SBO = 128;
constexpr int64_t IMG_SLAB = 256 * 128, IMG_PLANE = 256 * 64;     // bytes (img_common.cuh, C = 256)

struct BArgs {
    const uint8_t* dz3;        // split image (rows_pad, 256) of dZ_3
    const uint8_t* h[3];       // split images of H_2, H_1, H_0 (only the hi planes are read: activation signs)
    uint8_t* dz[3];            // split images of dZ_2, dZ_1, dZ_0 (written)
    float* dX0;                // (rows, ldx) fp32; columns [0, nx0) written
    const uint8_t* wpack;      // 64 chunk images [hi 8 KB | lo 8 KB] of W4^T, W3[:, :256]^T, W2^T, W1[:, :nx0]^T (zero padded to 256 rows)
    int64_t rows;
    int ldx, nx0;
};

__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
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
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ uint4 ldg16(const uint8_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// LeakyReLU derivative from a packed pair of saved bf16 outputs
__device__ __forceinline__ float slope_lo(uint32_t p) { return img::bf16_lo_f(p) > 0.f ? 1.f : 0.01f; }
__device__ __forceinline__ float slope_hi(uint32_t p) { return img::bf16_hi_f(p) > 0.f ? 1.f : 0.01f; }

__global__ void __launch_bounds__(NTHREADS, 1) nbr_bwd_f16_kernel(const __grid_constant__ BArgs A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    const uint32_t b0 = smem_u32(bars);
    const uint32_t bar_wfull = b0, bar_wempty = bar_wfull + 8 * NSW, bar_afull = bar_wempty + 8 * NSW, bar_aempty = bar_afull + 8 * NSA,
                   bar_gfull = bar_aempty + 8 * NSA, bar_gempty = bar_gfull + 8 * NSG, bar_actfull = bar_gempty + 8 * NSG,
                   bar_accfull = bar_actfull + 64, bar_accempty = bar_accfull + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    const int64_t ntiles = (A.rows + TM - 1) / TM;

    if (tid == 0) {
        for (int s = 0; s < NSW; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1); }
        for (int s = 0; s < NSA; ++s) { mbar_init(bar_afull + 8 * s, 4); mbar_init(bar_aempty + 8 * s, 1); }
        for (int s = 0; s < NSG; ++s) { mbar_init(bar_gfull + 8 * s, 4); mbar_init(bar_gempty + 8 * s, 4); }
        for (int s = 0; s < 8; ++s) mbar_init(bar_actfull + 8 * s, 4);
        for (int s = 0; s < 2; ++s) { mbar_init(bar_accfull + 8 * s, 1); mbar_init(bar_accempty + 8 * s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 12) {
        setmaxnreg_dec<40>();
        if (warp == 13 && lane == 0) {
            // ================= bulk-copy producer: one 16 KB W^T chunk image per stage =================
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int c = 0; c < NLAYER * NC; ++c, ++it) {
                    const uint32_t s = it % NSW, ph = (it / NSW) & 1;
                    mbar_wait_relaxed(bar_wempty + 8 * s, ph ^ 1, 64);
                    mbar_arrive_expect_tx(bar_wfull + 8 * s, W_STAGE);
                    bulk_g2s(smem_u32(smem + OFF_W + s * W_STAGE), A.wpack + (size_t)c * W_STAGE, W_STAGE, bar_wfull + 8 * s);
                }
            }
        } else if (warp == 12) {
            // ================= MMA issuer (warp-uniform schedule, one elected lane issues 3 MMAs per K chunk) =================
            uint32_t wit = 0, ait = 0, gen = 0, ti = 0;
            const uint32_t act_hi = smem_u32(smem + OFF_ACT), act_lo = act_hi + ACT_PART;
            const uint32_t w_base = smem_u32(smem + OFF_W), a_base = smem_u32(smem + OFF_A);
            const uint64_t dW = umma_desc(0, W_LBO, SBO), dA = umma_desc(0, A_LBO, SBO);
            auto issue = [&](uint32_t acc, uint32_t a_hi_addr, uint32_t a_lo_addr, bool first, uint32_t extra_commit) {
                const uint32_t s = wit % NSW, ph = (wit / NSW) & 1;
                ++wit;
                mbar_wait(bar_wfull + 8 * s, ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t w = w_base + s * W_STAGE;
                    const uint64_t w_hi = dW | (uint64_t)((w & 0x3FFFFu) >> 4), w_lo = dW | (uint64_t)(((w + W_PART) & 0x3FFFFu) >> 4);
                    const uint64_t a_hi = dA | (uint64_t)((a_hi_addr & 0x3FFFFu) >> 4), a_lo = dA | (uint64_t)((a_lo_addr & 0x3FFFFu) >> 4);
                    tc_mma_bf16(acc, a_hi, w_hi, IDESC, first ? 0u : 1u);
                    tc_mma_bf16(acc, a_lo, w_hi, IDESC, 1u);
                    tc_mma_bf16(acc, a_hi, w_lo, IDESC, 1u);
                    tc_commit(bar_wempty + 8 * s);
                    if (extra_commit) tc_commit(extra_commit);
                }
                __syncwarp();
            };
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
                const uint32_t acc0 = tmem_base, acc1 = tmem_base + HID;
                // ---- layer j = 0 (W4): operand = dZ_3 chunks from the loader ring; accumulator 0 was last read by epilogue 2 of the previous tile
                if (ti > 0) mbar_wait(bar_accempty, (ti - 1) & 1);
                for (int c = 0; c < NC; ++c, ++ait) {
                    const uint32_t s = ait % NSA, ph = (ait / NSA) & 1;
                    mbar_wait_relaxed(bar_afull + 8 * s, ph, 20);
                    const uint32_t a = a_base + s * A_STAGE;
                    issue(acc0, a, a + A_PART, c == 0, bar_aempty + 8 * s);
                }
                if (elect_one()) tc_commit(bar_accfull);
                __syncwarp();
                // ---- layers j = 1..3: operand = gated gradient written by the previous epilogue, released per 32 columns.
                //      accumulator 1 was last read by the final epilogue of the previous tile
                if (ti > 0) mbar_wait(bar_accempty + 8, (ti - 1) & 1);
#pragma unroll 1
                for (int j = 1; j < NLAYER; ++j) {
                    const uint32_t acc = (j & 1) ? acc1 : acc0;
                    for (int c = 0; c < NC; ++c) {
                        if ((c & 1) == 0) mbar_wait_relaxed(bar_actfull + 8 * (c >> 1), gen & 1, 20);
                        issue(acc, act_hi + c * 2 * A_LBO, act_lo + c * 2 * A_LBO, c == 0, 0);
                    }
                    if (elect_one()) tc_commit(bar_accfull + 8 * (j & 1));
                    __syncwarp();
                    ++gen;
                }
            }
        }
    } else if (warp >= 8) {
        // ================= loaders: thread = row; item = four 16-byte pieces of that row =================
        //   items 0..15  : K chunk c of the dZ_3 image (hi groups 2c, 2c+1 | lo groups 2c, 2c+1)  -> operand ring
        //   items 16..39 : hi plane of block jb of H_{2-j} (groups 4jb..4jb+3), j = 0..2             -> gate ring
        setmaxnreg_dec<120>();
        const int r = tid - NEPI;
        uint32_t ait = 0, git = 0;
        uint4 q[DEPTH][4];
        auto load_item = [&](int64_t tile, int i, uint4 (&d)[4]) {
            if (tile >= ntiles) return;
            const int64_t row = tile * TM + r;
            const int64_t ro = (row >> 5) * IMG_SLAB + (row & 31) * 16;
            if (i < NC) {
                if (row < A.rows) {
                    const uint8_t* p = A.dz3 + ro + (int64_t)(2 * i) * 512;
                    d[0] = ldg16(p); d[1] = ldg16(p + 512); d[2] = ldg16(p + IMG_PLANE); d[3] = ldg16(p + IMG_PLANE + 512);
                } else {
                    // padding rows of the last tile carry no gradient; the image is zeroed there for the weight-gradient kernel
                    d[0] = d[1] = d[2] = d[3] = make_uint4(0u, 0u, 0u, 0u);
                    uint8_t* p = const_cast<uint8_t*>(A.dz3) + ro + (int64_t)(2 * i) * 512;
                    *reinterpret_cast<uint4*>(p) = d[0]; *reinterpret_cast<uint4*>(p + 512) = d[0];
                    *reinterpret_cast<uint4*>(p + IMG_PLANE) = d[0]; *reinterpret_cast<uint4*>(p + IMG_PLANE + 512) = d[0];
                }
            } else {
                const int j = (i - NC) >> 3, jb = (i - NC) & 7;
                const uint8_t* p = (j == 0 ? A.h[0] : (j == 1 ? A.h[1] : A.h[2])) + ro + (int64_t)(4 * jb) * 512;
                d[0] = ldg16(p); d[1] = ldg16(p + 512); d[2] = ldg16(p + 1024); d[3] = ldg16(p + 1536);
            }
        };
        auto put_item = [&](int i, const uint4 (&d)[4]) {
            if (i < NC) {
                const uint32_t st = ait % NSA, ph = (ait / NSA) & 1;
                ++ait;
                mbar_wait_relaxed(bar_aempty + 8 * st, ph ^ 1, 32);
                uint8_t* stage = smem + OFF_A + st * A_STAGE + r * 16;
                *reinterpret_cast<uint4*>(stage) = d[0];
                *reinterpret_cast<uint4*>(stage + A_LBO) = d[1];
                *reinterpret_cast<uint4*>(stage + A_PART) = d[2];
                *reinterpret_cast<uint4*>(stage + A_PART + A_LBO) = d[3];
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_afull + 8 * st);
            } else {
                const uint32_t st = git % NSG, ph = (git / NSG) & 1;
                ++git;
                mbar_wait_relaxed(bar_gempty + 8 * st, ph ^ 1, 32);
                uint8_t* stage = smem + OFF_G + st * G_STAGE + r * 16;
#pragma unroll
                for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(stage + k * 2048) = d[k];
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_gfull + 8 * st);
            }
        };
#pragma unroll
        for (int i = 0; i < DEPTH; ++i) load_item(blockIdx.x, i, q[i]);
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll
            for (int i = 0; i < NITEM; ++i) {
                put_item(i, q[i % DEPTH]);
                if (i + DEPTH < NITEM) load_item(tile, i + DEPTH, q[i % DEPTH]);
                else load_item(tile + gridDim.x, i + DEPTH - NITEM, q[i % DEPTH]);
            }
        }
    } else {
        // ================= epilogue warps: thread = (row = TMEM lane, group wg taking 32-column blocks wg, wg+2, ...) =================
        setmaxnreg_inc<168>();
        const int wg = warp >> 2;
        const int r = tid & (TM - 1);
        const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
        uint8_t* act_hi = smem + OFF_ACT;
        uint32_t ti = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            const int64_t row = tile * TM + r;
            const int64_t ro = (row >> 5) * IMG_SLAB + (row & 31) * 16;
#pragma unroll 1
            for (int j = 0; j < NLAYER; ++j) {
                const uint32_t b = j & 1;
                mbar_wait_relaxed(bar_accfull + 8 * b, (j >> 1) & 1, 20);     // each accumulator completes twice per tile
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_base + b * HID + wg * 32;
                uint32_t va[32], vb[32];
                tmem_ld32_issue(taddr, va);
                if (j < NLAYER - 1) {
                    uint8_t* out = j == 0 ? A.dz[0] : (j == 1 ? A.dz[1] : A.dz[2]);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int jb = 2 * jj + wg;                    // column block
                        uint32_t(&cur)[32] = (jj & 1) ? vb : va;
                        uint32_t(&nxt)[32] = (jj & 1) ? va : vb;
                        // activation signs of this block (gate ring: items arrive in block order, even blocks -> group 0, odd -> group 1)
                        const uint32_t gi = ti * 24u + (uint32_t)j * 8u + (uint32_t)jb, st = gi % NSG, ph = (gi / NSG) & 1;
                        mbar_wait_relaxed(bar_gfull + 8 * st, ph, 20);
                        uint4 g4[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) g4[k] = *reinterpret_cast<const uint4*>(smem + OFF_G + st * G_STAGE + k * 2048 + r * 16);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_gempty + 8 * st);
                        tmem_ld_wait(cur);
                        if (jj + 1 < 4) tmem_ld32_issue(taddr + (jj + 1) * 64, nxt);
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            const uint32_t gw[4] = {g4[qd].x, g4[qd].y, g4[qd].z, g4[qd].w};
                            float y[8];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                y[2 * u] = __uint_as_float(cur[8 * qd + 2 * u]) * slope_lo(gw[u]);
                                y[2 * u + 1] = __uint_as_float(cur[8 * qd + 2 * u + 1]) * slope_hi(gw[u]);
                            }
                            uint4 hi, lo;
                            img::split8_bf16(y, hi, lo);
                            uint8_t* dst = act_hi + (jb * 4 + qd) * A_LBO + r * 16;
                            *reinterpret_cast<uint4*>(dst) = hi;
                            *reinterpret_cast<uint4*>(dst + ACT_PART) = lo;
                            uint8_t* gdst = out + ro + (int64_t)(jb * 4 + qd) * 512;
                            *reinterpret_cast<uint4*>(gdst) = hi;
                            *reinterpret_cast<uint4*>(gdst + IMG_PLANE) = lo;
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_actfull + 8 * jb);
                    }
                    if (j == 2) {           // accumulator 0 drained: layer 0 of the next tile may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_accempty);
                    }
                } else {
                    // ---- last layer: dX0 rows (fp32, reference column order), no gate
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int jb = 2 * jj + wg;
                        uint32_t(&cur)[32] = (jj & 1) ? vb : va;
                        uint32_t(&nxt)[32] = (jj & 1) ? va : vb;
                        tmem_ld_wait(cur);
                        if (jj + 1 < 4) tmem_ld32_issue(taddr + (jj + 1) * 64, nxt);
                        if (row < A.rows && jb * 32 < A.nx0) {
                            float4* o = reinterpret_cast<float4*>(A.dX0 + row * A.ldx + jb * 32);
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                o[i] = make_float4(__uint_as_float(cur[4 * i]), __uint_as_float(cur[4 * i + 1]), __uint_as_float(cur[4 * i + 2]),
                                                   __uint_as_float(cur[4 * i + 3]));
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_accempty + 8);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------------------
// dE = dZ_2 . W3[:, 256:263]: gradient of block3's 7 extra inputs (colour, dir - view, <dir, view>), read from the dZ_2 image.
// thread = row; W slice in shared memory; HBM-bound (one pass over the image).
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) dz_extras_kernel(const uint8_t* __restrict__ dz, const float* __restrict__ W, int ldw, int k0,
                                                        int64_t rows, float* __restrict__ dE) {
    __shared__ float ws[HID * 8];                   // [n][8]: 7 weights + pad
    for (int i = threadIdx.x; i < HID * 8; i += blockDim.x) {
        const int n = i >> 3, jx = i & 7;
        ws[i] = jx < 7 ? W[(int64_t)n * ldw + k0 + jx] : 0.f;
    }
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const uint8_t* p = dz + (row >> 5) * IMG_SLAB + (row & 31) * 16;
    float acc[7];
#pragma unroll
    for (int jx = 0; jx < 7; ++jx) acc[jx] = 0.f;
#pragma unroll 4
    for (int g = 0; g < HID / 8; ++g) {
        const uint4 hi = ldg16(p + g * 512), lo = ldg16(p + IMG_PLANE + g * 512);
        float v[8];
        img::join8_bf16(hi, lo, v);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 w0 = *reinterpret_cast<const float4*>(ws + (g * 8 + u) * 8), w1 = *reinterpret_cast<const float4*>(ws + (g * 8 + u) * 8 + 4);
            acc[0] = fmaf(v[u], w0.x, acc[0]); acc[1] = fmaf(v[u], w0.y, acc[1]); acc[2] = fmaf(v[u], w0.z, acc[2]); acc[3] = fmaf(v[u], w0.w, acc[3]);
            acc[4] = fmaf(v[u], w1.x, acc[4]); acc[5] = fmaf(v[u], w1.y, acc[5]); acc[6] = fmaf(v[u], w1.z, acc[6]);
        }
    }
#pragma unroll
    for (int jx = 0; jx < 7; ++jx) dE[row * 7 + jx] = acc[jx];
}

}  // namespace

// bytes of the packed W^T image of hnr_nbr_bwd_f16: 4 layers x 16 chunks x [hi 8 KB | lo 8 KB] (host packer: mlp_tc.pack_mlp_bwd)
extern "C" int64_t hnr_nbr_bwd_f16_packed_bytes(void) { return (int64_t)NLAYER * NC * W_STAGE; }

// Fused data-gradient chain of the per-neighbour MLP (see the header of this file).  All images are split images
// (img_common.cuh) of (rows padded to 128) x 256; dz3 is read (its padding rows are zeroed here), h2/h1/h0 are the forward's
// saved activations (hi planes read), dz2/dz1/dz0 are written, dX0 (rows, ldx) receives the first nx0 (multiple of 32,
// <= 256) input-gradient columns of layer 0 in the order of wpackT's rows.
extern "C" int hnr_nbr_bwd_f16(const void* dz3, const void* h2, const void* h1, const void* h0, void* dz2, void* dz1, void* dz0, float* dX0,
                               int64_t ldx, int64_t nx0, const void* wpackT, int64_t rows, void* stream) {
    if (rows == 0) return HNR_OK;
    HNR_CHECK_ARG(nx0 > 0 && nx0 % 32 == 0 && nx0 <= HID && ldx >= nx0 && ldx % 4 == 0, "nbr_bwd_f16: nx0 must be a multiple of 32 <= 256, ldx % 4 == 0");
    BArgs A{};
    A.dz3 = (const uint8_t*)dz3; A.h[0] = (const uint8_t*)h2; A.h[1] = (const uint8_t*)h1; A.h[2] = (const uint8_t*)h0;
    A.dz[0] = (uint8_t*)dz2; A.dz[1] = (uint8_t*)dz1; A.dz[2] = (uint8_t*)dz0; A.dX0 = dX0; A.ldx = (int)ldx; A.nx0 = (int)nx0;
    A.wpack = (const uint8_t*)wpackT; A.rows = rows;
    static bool configured = false;
    if (!configured) {
        HNR_CUDA(cudaFuncSetAttribute(nbr_bwd_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const int64_t ntiles = hnr_cdiv(rows, TM);
    const int grid = (int)(ntiles < HNR_NUM_SMS ? ntiles : HNR_NUM_SMS);
    nbr_bwd_f16_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(A);
    HNR_CHECK_LAUNCH("nbr_bwd_f16");
    return HNR_OK;
}

// dE (rows, 7) = dZ (image, rows x 256) . W[:, k0:k0+7]   (W: (256, ldw) fp32 row-major)
extern "C" int hnr_dz_extras_bwd(const void* dz, const float* W, int64_t ldw, int64_t k0, int64_t rows, float* dE, void* stream) {
    if (rows == 0) return HNR_OK;
    dz_extras_kernel<<<(unsigned)hnr_cdiv(rows, 128), 128, 0, (cudaStream_t)stream>>>((const uint8_t*)dz, W, (int)ldw, (int)k0, rows, dE);
    HNR_CHECK_LAUNCH("dz_extras_bwd");
    return HNR_OK;
}
