// Fused per-neighbour MLP on the 5th-gen tensor cores, second generation (SURVEY.md §8a rows G1 + A1-A3).
//
//     x0(284) -> 256 -> 256 -> [+colour, dir-view, <dir,view>](263) -> 256 -> 256 -> density head, weighted K-sum
//
// One persistent CTA per SM walks tiles of 128 neighbour rows (16 shading samples x K=8).  What differs from the
// first-generation kernel (mlp_tc.cu, 3xTF32):
//   * fp32 accuracy from THREE FP16 MMAs per product ("3xFP16"): every operand v (pre-multiplied by a power-of-two
//     scale so that hi and lo stay in fp16's normal range) is split into hi = fp16(v), lo = fp16(v - hi) and
//     a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo is accumulated in fp32 in TMEM.  22 mantissa bits, like 3xTF32, but
//     tcgen05.mma.kind::f16 runs at twice the TF32 rate and the operands take half the shared memory;
//   * because hi+lo fp16 take exactly the 4 bytes of the fp32 value, the layer activations live in shared memory
//     ALREADY in the canonical UMMA operand layout: the epilogue of layer l writes layer l+1's A operand in place
//     and no conversion ring is needed for layers 1..3;
//   * two TMEM accumulators (2 x 256 columns): the MMAs of layer l+1 start on the first K-chunks of the new
//     activation while the epilogue of layer l is still draining the rest (per-32-column mbarriers), and the
//     layer-0 MMAs of the NEXT tile run under the last epilogue + K-sum of the current tile;
//   * warp roles (4 warpgroups, registers re-balanced with setmaxnreg): 2 x 4 epilogue warps (each TMEM lane quarter is
//     served by two warps that take alternate 32-column blocks), 4 generator warps (gather from the point tables,
//     positional encodings, layer-0 operand chunks, one tile ahead), 1 MMA warp (one elected thread), 1 bulk-copy warp;
//     waiting warps back off with nanosleep so that their polling does not take issue slots from the working ones.
//
// The arithmetic restated is models/aggregators/point_aggregators.py:921-972, :1002-1036 and
// models/helpers/networks.py:175-189 of the reference.
#include <cuda_fp16.h>

#include "common.cuh"
#include "hnr.h"
#define TRACE_SRC ((long long*)nullptr)
#include "tc_common.cuh"
#include "img_common.cuh"

namespace {
using namespace tc;

constexpr int TM = 128;                 // rows (neighbours) per tile == UMMA M
constexpr int HID = 256;                // layer width == UMMA N
constexpr int KC = 16;                  // K elements per chunk == one f16 MMA K-step (32 bytes per row)
constexpr int NSW = 3;                  // weight ring stages
constexpr int NSA = 3;                  // layer-0 operand ring stages
constexpr int W_PART = HID * KC * 2;    // 8192 B: hi (or lo) weights of one chunk
constexpr int W_STAGE = 2 * W_PART;     // 16384
constexpr int A_PART = TM * KC * 2;     // 4096 B
constexpr int A_STAGE = 2 * A_PART;     // 8192
constexpr int ACT_PART = TM * HID * 2;  // 65536 B: hi (or lo) of a 128 x 256 activation
constexpr int NLAYER = 4;
constexpr int NC0 = 18, NC1 = 16, NC2 = 17, NC3 = 16;      // K chunks per layer (288, 256, 16 + 256, 256)
constexpr int NCHUNK_TILE = NC0 + NC1 + NC2 + NC3;
constexpr int FEAT = 32, NF_VIEW = 4, X5_W = 280;
constexpr int NEPI = 256, NTHREADS = 512;      // warps 0-7 epilogue, 8-11 generators, 12 MMA, 13 bulk copy, 14-15 idle

// shared memory map (bytes)
constexpr int OFF_W = 0;
constexpr int OFF_A = OFF_W + NSW * W_STAGE;               // 49152
constexpr int OFF_E = OFF_A + NSA * A_STAGE;               // 73728   block3 extras chunk, double buffered by tile parity
constexpr int OFF_ACT = OFF_E + 2 * A_STAGE;               // 90112   hi | lo ; re-used as the fp32 staging of the K-sum
constexpr int OFF_BIAS = OFF_ACT + 2 * ACT_PART;           // 221184  (4,256) fp32, pre-scaled
constexpr int OFF_WALPHA = OFF_BIAS + NLAYER * HID * 4;    // 225280
constexpr int OFF_WC = OFF_WALPHA + HID * 4;               // 226304  neighbour weight * conf, double buffered
constexpr int OFF_ARAW = OFF_WC + 2 * TM * 4;              // 227328  density-head partial dot products of the two epilogue groups
constexpr int OFF_GIDX = OFF_ARAW + 2 * TM * 4;            // 228352  point index of every row (per-point layer-0 partial), double buffered
constexpr int OFF_BAR = OFF_GIDX + 2 * TM * 4;             // 229376
constexpr int NBAR = 2 * NSW + 2 * NSA + 2 + 8 + 2 + 1 + 1;
constexpr int SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);   // D=f32, A=B=f16, K-major
constexpr uint32_t A_LBO = (TM / 8) * 128, W_LBO = (HID / 8) * 128, SBO = 128;

struct F16Args {
    const float *xyz, *xyz_pers, *emb, *color, *dir;
    const int32_t *pidx, *vlist;
    const float *loc_w, *loc_pers, *raydirs, *cam, *weight, *confc;
    const uint8_t* wpack;          // 67 chunk images in consumption order
    const float *bias, *walpha, *balpha;   // bias: (4,256), rows 0..2 pre-multiplied by the next layer's input scale
    float *sigma, *X5, *dbg, *araw;     // dbg: optional (4, Nv*8, 256) activations of every layer (tests); araw: (Nv*8) density pre-activation
    // training forward (MODE 2): split images (img_common.cuh) of the layer-0 input in kernel column order (288 wide), the
    // block3 extras chunk (16 wide) and the four layers' outputs (256 wide), rows padded to the tile -- the operands of the
    // fused backward kernels (nbr_bwd_f16.cu, wgrad_img.cu)
    uint8_t *x0img, *eimg, *himg[NLAYER];
    float inv_scale0, inv_scale2;
    float inv_act;                      // 1 / input scale of layers 1..3 (the saved activations are unscaled)
    int32_t* status;                    // optional: bit 0 is set when a scaled activation left fp16's range and was saturated
    // per-point layer-0 partial (inference): 224 of layer 0's 284 inputs -- the embedding and its positional encoding -- depend on the
    // POINT only, and a point is the neighbour of ~28 samples of a frame.  pp (N_points, 256) fp32 = W1[:, :224] . [emb | PE(emb)] is
    // computed once per (weights, point set); the kernel then generates only the 4 distance-encoding chunks of layer 0 (nc0 = 4, wpack
    // starting at chunk 14) and the layer-0 epilogue adds pp[point] * pp_scale to the pre-activation.  NULL = off (nc0 = 18).
    const float* pp;
    float pp_scale;
    int nc0;
    int64_t Nv;
    float mul[NLAYER];             // accumulator -> (scaled) pre-activation factor per layer
    float scale0, scale2;          // input scales of layer 0 (generated features) and layer 2 (extras chunk)
};

__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// tcgen05.ld without the wait; the matching wait names the registers as in/out operands so that no use of them
// can be scheduled before it
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

template <int ID, int N>
__device__ __forceinline__ void named_barrier() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }

__device__ __forceinline__ void rot3(const float* m, float x, float y, float z, float& ox, float& oy, float& oz) {
    ox = x * m[0] + y * m[3] + z * m[6];
    oy = x * m[1] + y * m[4] + z * m[7];
    oz = x * m[2] + y * m[5] + z * m[8];
}
// sin/cos of small arguments (|x| of a few radians: embedding channels, scaled distances, view directions) on the MUFU
// pipe: absolute error ~5e-7, two orders of magnitude below what rtol 1e-4 on the rendered output needs, at a fifth
// of the instructions of the full-range routine
__device__ __forceinline__ void sincos_fast(float x, float* s, float* c) { __sincosf(x, s, c); }
__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// two fp32 -> packed fp16x2 {lo half = a, hi half = b}, round to nearest, saturating at +-65504 (an activation beyond fp16's
// range would otherwise become inf and poison the accumulator with inf - inf)
__device__ __forceinline__ uint32_t pack_sat(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// 8 fp32 values -> 8 fp16 hi (one 16-byte core-matrix row) and 8 fp16 lo = fp16(v - hi)
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

// mbarrier wait for warps that are ahead of the pipeline anyway: poll, then sleep between polls
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

// store one 16-wide K chunk of row r (already scaled) into an operand stage: [hi: kb0 | kb1][lo: kb0 | kb1]
__device__ __forceinline__ void store_chunk16(uint8_t* stage, int r, const float* v) {
    uint4 hi, lo;
    split8(v, hi, lo);
    *reinterpret_cast<uint4*>(stage + r * 16) = hi;
    *reinterpret_cast<uint4*>(stage + A_PART + r * 16) = lo;
    split8(v + 8, hi, lo);
    *reinterpret_cast<uint4*>(stage + A_LBO + r * 16) = hi;
    *reinterpret_cast<uint4*>(stage + A_PART + A_LBO + r * 16) = lo;
}

// one 16-wide chunk of row `row` (values pre-multiplied by 1/inv) -> column groups 2c, 2c+1 of a split image with C columns
__device__ __forceinline__ void store_chunk16_img(uint8_t* image, int C, int64_t row, int chunk, const float* v, float inv) {
    float u[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) u[i] = v[i] * inv;
    uint8_t* p = image + img::piece_off(row, 2 * chunk, C);
    uint4 hi, lo;
    img::split8_bf16(u, hi, lo);
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + img::plane_bytes(C)) = lo;
    img::split8_bf16(u + 8, hi, lo);
    *reinterpret_cast<uint4*>(p + 512) = hi;
    *reinterpret_cast<uint4*>(p + 512 + img::plane_bytes(C)) = lo;
}

// MODE 0: inference; 1: fp32 copies of every layer's output (tests); 2: training forward, split images saved for the backward
template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) nbr_mlp_f16_kernel(F16Args A) {
    constexpr bool DBG = MODE == 1;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    const uint32_t b0 = smem_u32(bars);
    const uint32_t bar_wfull = b0, bar_wempty = b0 + 8 * NSW, bar_afull = bar_wempty + 8 * NSW, bar_aempty = bar_afull + 8 * NSA,
                   bar_efull = bar_aempty + 8 * NSA, bar_actfull = bar_efull + 16, bar_accfull = bar_actfull + 64,
                   bar_accempty0 = bar_accfull + 16, bar_tiledone = bar_accempty0 + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
    float* walpha_s = reinterpret_cast<float*>(smem + OFF_WALPHA);
    float* wc_s = reinterpret_cast<float*>(smem + OFF_WC);
    float* araw_s = reinterpret_cast<float*>(smem + OFF_ARAW);

    const int64_t total_rows = A.Nv * 8;
    const int64_t ntiles = (total_rows + TM - 1) / TM;

    if (tid == 0) {
        for (int s = 0; s < NSW; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1); }
        for (int s = 0; s < NSA; ++s) { mbar_init(bar_afull + 8 * s, 4); mbar_init(bar_aempty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_efull + 8 * s, 4); mbar_init(bar_accfull + 8 * s, 1); }
        for (int s = 0; s < 8; ++s) mbar_init(bar_actfull + 8 * s, 4);
        mbar_init(bar_accempty0, 8);
        mbar_init(bar_tiledone, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) tmem_alloc(smem_u32(tmem_slot), 512);
    for (int i = tid; i < NLAYER * HID; i += NTHREADS) bias_s[i] = A.bias[i];
    if (tid < HID) walpha_s[tid] = A.walpha[tid];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 12) {
        setmaxnreg_dec<40>();
        if (warp == 13 && lane == 0) {
            // ================= bulk-copy producer: one 16 KB weight chunk image per stage =================
            uint32_t it = 0;
            const int nchunk = A.nc0 + NC1 + NC2 + NC3;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int c = 0; c < nchunk; ++c, ++it) {
                    const uint32_t s = it % NSW, ph = (it / NSW) & 1;
                    mbar_wait_relaxed(bar_wempty + 8 * s, ph ^ 1, 64);
                    mbar_arrive_expect_tx(bar_wfull + 8 * s, W_STAGE);
                    bulk_g2s(smem_u32(smem + OFF_W + s * W_STAGE), A.wpack + (size_t)c * W_STAGE, W_STAGE, bar_wfull + 8 * s);
                }
            }
        } else if (warp == 12) {
            // ================= MMA issuer: the whole warp walks the schedule (warp-uniform), one elected lane issues
            //                   3 MMAs (3xFP16) per K chunk =================
            uint32_t wit = 0, ait = 0, gen = 0, ti = 0;
            const uint32_t act_hi = smem_u32(smem + OFF_ACT), act_lo = act_hi + ACT_PART;
            const uint32_t w_base = smem_u32(smem + OFF_W), a_base = smem_u32(smem + OFF_A), e_base = smem_u32(smem + OFF_E);
            // descriptor = constant high word | (address >> 4) in the low 14 bits | LBO field
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
                    tc_mma_f16(acc, a_hi, w_hi, IDESC, first ? 0u : 1u);
                    tc_mma_f16(acc, a_lo, w_hi, IDESC, 1u);
                    tc_mma_f16(acc, a_hi, w_lo, IDESC, 1u);
                    tc_commit(bar_wempty + 8 * s);
                    if (extra_commit) tc_commit(extra_commit);
                }
                __syncwarp();
            };
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
                const uint32_t acc0 = tmem_base, acc1 = tmem_base + HID;
                // ---- layer 0: operands from the generator ring (accumulator 0 was last read by epilogue 2 of the previous tile)
                if (ti > 0) mbar_wait(bar_accempty0, (ti - 1) & 1);
                for (int c = 0; c < A.nc0; ++c, ++ait) {
                    const uint32_t s = ait % NSA, ph = (ait / NSA) & 1;
                    mbar_wait_relaxed(bar_afull + 8 * s, ph, 20);
                    const uint32_t a = a_base + s * A_STAGE;
                    issue(acc0, a, a + A_PART, c == 0, bar_aempty + 8 * s);
                }
                if (elect_one()) tc_commit(bar_accfull);
                __syncwarp();
                // ---- layer 1: operands = activation written by epilogue 0, released per 32 columns
                for (int c = 0; c < NC1; ++c) {
                    if ((c & 1) == 0) mbar_wait_relaxed(bar_actfull + 8 * (c >> 1), gen & 1, 20);
                    issue(acc1, act_hi + c * 2 * A_LBO, act_lo + c * 2 * A_LBO, c == 0, 0);
                }
                if (elect_one()) tc_commit(bar_accfull + 8);
                __syncwarp();
                ++gen;
                // ---- layer 2: extras chunk first (ready since the gather), then the activation of epilogue 1
                {
                    const uint32_t p = ti & 1;
                    mbar_wait(bar_efull + 8 * p, (ti >> 1) & 1);
                    const uint32_t e = e_base + p * A_STAGE;
                    issue(acc0, e, e + A_PART, true, 0);
                }
                for (int c = 0; c < NC1; ++c) {
                    if ((c & 1) == 0) mbar_wait_relaxed(bar_actfull + 8 * (c >> 1), gen & 1, 20);
                    issue(acc0, act_hi + c * 2 * A_LBO, act_lo + c * 2 * A_LBO, false, 0);
                }
                if (elect_one()) tc_commit(bar_accfull);
                __syncwarp();
                ++gen;
                // ---- layer 3
                for (int c = 0; c < NC3; ++c) {
                    if ((c & 1) == 0) mbar_wait_relaxed(bar_actfull + 8 * (c >> 1), gen & 1, 20);
                    issue(acc1, act_hi + c * 2 * A_LBO, act_lo + c * 2 * A_LBO, c == 0, 0);
                }
                if (elect_one()) tc_commit(bar_accfull + 8);
                __syncwarp();
                ++gen;
            }
        }
    } else if (warp >= 8) {
        // ================= generators: gather + layer-0 operand chunks, thread = row =================
        setmaxnreg_inc<200>();
        const int r = tid - NEPI;
        uint32_t ait = 0, ti = 0;
        const float* rt = A.cam + 12;
        int64_t s_cur = 0, g_cur = 0;
        bool live_cur = false;
        auto load_idx = [&](int64_t tile, int64_t& s, int64_t& g, bool& live) {
            const int64_t row = tile * TM + r;
            live = tile < ntiles && row < total_rows;
            const int64_t v = live ? (row >> 3) : 0;
            s = A.vlist[v];
            g = live ? max(A.pidx[s * 8 + (row & 7)], 0) : 0;
        };
        if ((int64_t)blockIdx.x < ntiles) load_idx(blockIdx.x, s_cur, g_cur, live_cur);
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            const int64_t row = tile * TM + r;
            const int k = (int)(row & 7);
            const int64_t s = s_cur, g = g_cur;
            const bool live = live_cur;
            // ---- issue every load of this row, then the index loads of the next tile ----
            const bool use_pp = MODE == 0 && A.pp != nullptr;
            float4 e4[8];
            const float4* ep = reinterpret_cast<const float4*>(A.emb + g * FEAT);
            if (!use_pp) {
#pragma unroll
                for (int i = 0; i < 8; ++i) e4[i] = __ldg(ep + i);
            } else {
                // the epilogue of layer 0 reads this point's 1 KB row of the partial one tile later: request it into L2 now
                const char* prow = reinterpret_cast<const char*>(A.pp + g * HID);
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(prow + 128 * i));
            }
            const float px = A.xyz[g * 3], py = A.xyz[g * 3 + 1], pz = A.xyz[g * 3 + 2];
            const float dx = A.dir[g * 3], dy = A.dir[g * 3 + 1], dz = A.dir[g * 3 + 2];
            const float c0 = A.color[g * 3], c1 = A.color[g * 3 + 1], c2 = A.color[g * 3 + 2];
            const float lx = A.loc_w[s * 3], ly = A.loc_w[s * 3 + 1], lz = A.loc_w[s * 3 + 2];
            const float sx = A.loc_pers[s * 3], sy = A.loc_pers[s * 3 + 1], sz = A.loc_pers[s * 3 + 2];
            const float rx0 = A.raydirs[s * 3], ry0 = A.raydirs[s * 3 + 1], rz0 = A.raydirs[s * 3 + 2];
            float wgt = live ? A.weight[s * 8 + k] : 0.f;
            if (live && A.confc) wgt *= A.confc[s * 8 + k];
            float qx, qy, qz;
            if (A.xyz_pers) { qx = A.xyz_pers[g * 3]; qy = A.xyz_pers[g * 3 + 1]; qz = A.xyz_pers[g * 3 + 2]; }
            load_idx(tile + gridDim.x, s_cur, g_cur, live_cur);
            // ---- per-row geometry ----
            float d[6];
            rot3(rt, px - lx, py - ly, pz - lz, d[0], d[1], d[2]);
            if (!A.xyz_pers) {
                float cx, cy, cz;
                rot3(A.cam + 3, px - A.cam[0], py - A.cam[1], pz - A.cam[2], cx, cy, cz);
                qx = cx / cz; qy = cy / cz; qz = cz;
            }
            d[3] = qx * qz - sx * sz; d[4] = qy * qz - sy * sz; d[5] = qz - sz;
            float vx, vy, vz, rx, ry, rz;
            rot3(rt, rx0, ry0, rz0, vx, vy, vz);
            rot3(rt, dx, dy, dz, rx, ry, rz);
            // view-direction encoding of the sample (X5 columns 256..279): row k of a sample writes 3 of the 12 angles
            if (live) {
                float* xo = A.X5 + (row >> 3) * X5_W + HID;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int q = k * 3 + j;
                    if (q < 3 * NF_VIEW) {
                        const int c = q / NF_VIEW, f = q - c * NF_VIEW;
                        const float val = c == 0 ? vx : (c == 1 ? vy : vz);
                        float a, b;
                        sincosf(val * (float)(1 << f), &a, &b);
                        xo[q] = a;
                        xo[3 * NF_VIEW + q] = b;
                    }
                }
            }
            // ---- block3 extras chunk + row weight (buffers of tile parity p; last read by tile ti-2) ----
            {
                const uint32_t p = ti & 1;
                if (ti >= 2) mbar_wait_relaxed(bar_tiledone, ti & 1, 128);
                float ev[16];
                ev[0] = c0; ev[1] = c1; ev[2] = c2; ev[3] = rx - vx; ev[4] = ry - vy; ev[5] = rz - vz; ev[6] = rx * vx + ry * vy + rz * vz;
#pragma unroll
                for (int i = 7; i < 16; ++i) ev[i] = 0.f;
#pragma unroll
                for (int i = 0; i < 7; ++i) ev[i] *= A.scale2;
                store_chunk16(smem + OFF_E + p * A_STAGE, r, ev);
                if (MODE == 2) store_chunk16_img(A.eimg, 16, row, 0, ev, A.inv_scale2);
                wc_s[p * TM + r] = wgt;
                reinterpret_cast<int32_t*>(smem + OFF_GIDX)[p * TM + r] = (int32_t)g;
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_efull + 8 * p);
            }
            // ---- layer-0 operand chunks: every transcendental is evaluated BEFORE the first ring wait, so that once the
            //      MMA warp starts freeing stages only doublings, scaling and the fp16 split remain on the critical path ----
            const float* ef = reinterpret_cast<const float*>(e4);
            const float sc0 = A.scale0;
            float pe[64], sn[32], cs[32];
#pragma unroll
            for (int j = 0; j < 6; ++j) {                  // encoded distances: idx = f*12 + (sin|cos)*6 + component
                float a, b;                                // octaves 0 and 2 evaluated directly, 1, 3, 4 by angle doubling
                sincos_fast(d[j], &a, &b);
                pe[j] = a; pe[6 + j] = b;
                pe[12 + j] = 2.f * a * b; pe[18 + j] = 1.f - 2.f * a * a;
                sincos_fast(d[j] * 4.f, &a, &b);
                pe[24 + j] = a; pe[30 + j] = b;
                float a2 = 2.f * a * b, b2 = 1.f - 2.f * a * a;
                pe[36 + j] = a2; pe[42 + j] = b2;
                pe[48 + j] = 2.f * a2 * b2; pe[54 + j] = 1.f - 2.f * a2 * a2;
            }
#pragma unroll
            for (int i = 0; i < 60; ++i) pe[i] *= sc0;
            pe[60] = pe[61] = pe[62] = pe[63] = 0.f;
            if (!use_pp) {
#pragma unroll
                for (int i = 0; i < 32; ++i) sincos_fast(ef[i], &sn[i], &cs[i]);
            }
            int chunk_no = 0;
            auto put = [&](const float* v) {
                if (MODE == 2) store_chunk16_img(A.x0img, NC0 * KC, row, chunk_no, v, A.inv_scale0);
                ++chunk_no;
                const uint32_t st = ait % NSA, ph = (ait / NSA) & 1;
                ++ait;
                mbar_wait_relaxed(bar_aempty + 8 * st, ph ^ 1, 32);
                store_chunk16(smem + OFF_A + st * A_STAGE, r, v);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_afull + 8 * st);
            };
            if (!use_pp) {
            {   // raw embedding: 2 chunks
                float v[16];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = ef[c * 16 + i] * sc0;
                    put(v);
                }
            }
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {   // sin/cos(2^f e) per block of 8 channels: chunk = [sin x 8 | cos x 8]
                float v[16];
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    if (f > 0) {               // next octave by angle doubling
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float a = sn[cb * 8 + i], b = cs[cb * 8 + i];
                            sn[cb * 8 + i] = 2.f * a * b;
                            cs[cb * 8 + i] = 1.f - 2.f * a * a;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { v[i] = sn[cb * 8 + i] * sc0; v[8 + i] = cs[cb * 8 + i] * sc0; }
                    put(v);
                }
            }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) put(pe + c * 16);
        }
    } else {
        // ================= epilogue warps: thread = (row = TMEM lane, group wg taking 32-column blocks wg, wg+2, ...) =================
        setmaxnreg_inc<136>();
        const int wg = warp >> 2;                              // 0 | 1
        const int r = tid & (TM - 1);                          // warps w and w+4 own TMEM lanes 32(w&3)..+31
        const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
        uint8_t* act_hi = smem + OFF_ACT;
        float* stage = reinterpret_cast<float*>(smem + OFF_ACT);
        uint32_t ti = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            const int64_t row0 = tile * TM;
#pragma unroll 1
            for (int l = 0; l < NLAYER; ++l) {
                const uint32_t b = l & 1;
                mbar_wait_relaxed(bar_accfull + 8 * b, (l >> 1) & 1, 20);     // each accumulator completes twice per tile
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_base + b * HID + wg * 32;
                const float mul = A.mul[l];
                const float4* bl4 = reinterpret_cast<const float4*>(bias_s + l * HID + wg * 32);
                uint32_t va[32], vb[32];
                tmem_ld32_issue(taddr, va);
                float amax = 0.f;                                      // largest scaled pre-activation magnitude of this thread's blocks
                // per-point layer-0 partial: this row's point (written by the generators with the extras chunk of this tile)
                const float* pprow = nullptr;
                if (MODE == 0 && l == 0 && A.pp) {
                    const uint32_t p = ti & 1;
                    mbar_wait_relaxed(bar_efull + 8 * p, (ti >> 1) & 1, 20);
                    pprow = A.pp + (int64_t)reinterpret_cast<const int32_t*>(smem + OFF_GIDX)[p * TM + r] * HID + wg * 32;
                }
                const float pps = A.pp_scale;
                if (l < NLAYER - 1) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = 2 * jj + wg;                     // column block
                        uint32_t(&cur)[32] = (jj & 1) ? vb : va;
                        uint32_t(&nxt)[32] = (jj & 1) ? va : vb;
                        float4 pq[8];
                        if (MODE == 0 && pprow) {                      // in flight while the TMEM load completes
#pragma unroll
                            for (int i4 = 0; i4 < 8; ++i4) pq[i4] = __ldg(reinterpret_cast<const float4*>(pprow + jj * 64) + i4);
                        }
                        tmem_ld_wait(cur);
                        if (jj + 1 < 4) tmem_ld32_issue(taddr + (jj + 1) * 64, nxt);
                        float y[32];
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            float4 bb = bl4[jj * 16 + i4];
                            if (MODE == 0 && pprow) {
                                bb.x = fmaf(pq[i4].x, pps, bb.x); bb.y = fmaf(pq[i4].y, pps, bb.y);
                                bb.z = fmaf(pq[i4].z, pps, bb.z); bb.w = fmaf(pq[i4].w, pps, bb.w);
                            }
                            const float t0 = fmaf(__uint_as_float(cur[4 * i4 + 0]), mul, bb.x), t1 = fmaf(__uint_as_float(cur[4 * i4 + 1]), mul, bb.y);
                            const float t2 = fmaf(__uint_as_float(cur[4 * i4 + 2]), mul, bb.z), t3 = fmaf(__uint_as_float(cur[4 * i4 + 3]), mul, bb.w);
                            y[4 * i4 + 0] = fmaxf(t0, 0.01f * t0); y[4 * i4 + 1] = fmaxf(t1, 0.01f * t1);
                            y[4 * i4 + 2] = fmaxf(t2, 0.01f * t2); y[4 * i4 + 3] = fmaxf(t3, 0.01f * t3);
                            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(t0), fabsf(t1)), fmaxf(fabsf(t2), fabsf(t3))));
                        }
                        if (DBG) {
                            const int64_t row = row0 + r;
                            if (row < total_rows) {
                                float4* o = reinterpret_cast<float4*>(A.dbg + ((int64_t)l * total_rows + row) * HID + j * 32);
                                const float ia = A.inv_act;
#pragma unroll
                                for (int i = 0; i < 8; ++i) o[i] = make_float4(y[4 * i] * ia, y[4 * i + 1] * ia, y[4 * i + 2] * ia, y[4 * i + 3] * ia);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 hi, lo;
                            split8(y + q * 8, hi, lo);
                            uint8_t* dst = act_hi + (j * 4 + q) * A_LBO + r * 16;
                            *reinterpret_cast<uint4*>(dst) = hi;
                            *reinterpret_cast<uint4*>(dst + ACT_PART) = lo;
                            if (MODE == 2) {            // unscaled output as bf16 hi/lo: operand of the backward kernels
                                float u[8];
                                const float ia = A.inv_act;
#pragma unroll
                                for (int i = 0; i < 8; ++i) u[i] = y[q * 8 + i] * ia;
                                img::split8_bf16(u, hi, lo);
                                uint8_t* gp = A.himg[l] + img::piece_off(row0 + r, j * 4 + q, HID);
                                *reinterpret_cast<uint4*>(gp) = hi;
                                *reinterpret_cast<uint4*>(gp + img::plane_bytes(HID)) = lo;
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_actfull + 8 * j);
                    }
                    if (amax > 65000.f && A.status) atomicOr(A.status, 1);      // the saturating pack clipped a value: results are not fp32-exact
                    if (l == 2) {           // accumulator 0 drained: layer 0 of the next tile may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_accempty0);
                    }
                } else {
                    // ---- last layer: density head, row weight, fp32 staging for the K-sum ----
                    const uint32_t p = ti & 1;
                    mbar_wait_relaxed(bar_efull + 8 * p, (ti >> 1) & 1, 20);   // makes the generator's wc_s of this tile visible
                    const float wrow = wc_s[p * TM + r];
                    const float4* wa4 = reinterpret_cast<const float4*>(walpha_s + wg * 32);
                    float dot = 0.f;
                    // staging: element (col,row) at col*128 + ((row/4) ^ (col%32))*4 + row%4  (conflict-free for the row-wise
                    // writes here and for the column-wise float4 reads of the K-sum)
                    float* srow = stage + (r & 3);
                    const int rg = r >> 2;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = 2 * jj + wg;
                        uint32_t(&cur)[32] = (jj & 1) ? vb : va;
                        uint32_t(&nxt)[32] = (jj & 1) ? va : vb;
                        tmem_ld_wait(cur);
                        if (jj + 1 < 4) tmem_ld32_issue(taddr + (jj + 1) * 64, nxt);
                        float* sblk = srow + j * 32 * TM;
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            const float4 bb = bl4[jj * 16 + i4], ww = wa4[jj * 16 + i4];
                            const float bv[4] = {bb.x, bb.y, bb.z, bb.w}, wv[4] = {ww.x, ww.y, ww.z, ww.w};
                            float tsave[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int i = 4 * i4 + u;
                                float t = fmaf(__uint_as_float(cur[i]), mul, bv[u]);
                                t = fmaxf(t, 0.01f * t);
                                if (DBG) tsave[u] = t;
                                if (MODE == 2) cur[i] = __float_as_uint(t);      // kept for the image store below
                                dot = fmaf(t, wv[u], dot);
                                sblk[i * TM + ((rg ^ i) << 2)] = t * wrow;
                            }
                            if (DBG && row0 + r < total_rows)
                                reinterpret_cast<float4*>(A.dbg + ((int64_t)l * total_rows + row0 + r) * HID + j * 32)[i4] =
                                    make_float4(tsave[0], tsave[1], tsave[2], tsave[3]);
                        }
                        if (MODE == 2) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float u8[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) u8[i] = __uint_as_float(cur[q * 8 + i]);
                                uint4 hi, lo;
                                img::split8_bf16(u8, hi, lo);
                                uint8_t* gp = A.himg[l] + img::piece_off(row0 + r, j * 4 + q, HID);
                                *reinterpret_cast<uint4*>(gp) = hi;
                                *reinterpret_cast<uint4*>(gp + img::plane_bytes(HID)) = lo;
                            }
                        }
                    }
                    araw_s[wg * TM + r] = dot;
                    tc_fence_before();
                    named_barrier<1, NEPI>();
                    if (wg == 0) {
                        const float raw = araw_s[r] + araw_s[TM + r] + A.balpha[0];
                        if (MODE != 0 && A.araw && row0 + r < total_rows) A.araw[row0 + r] = raw;
                        float sg = wrow * softplus_t(raw - 1.f);
                        sg += __shfl_xor_sync(0xffffffffu, sg, 1);
                        sg += __shfl_xor_sync(0xffffffffu, sg, 2);
                        sg += __shfl_xor_sync(0xffffffffu, sg, 4);
                        const int64_t v = (row0 + r) >> 3;
                        if ((r & 7) == 0 && v < A.Nv) A.sigma[v] = sg;
                    }
                    // ---- weighted K-sum: thread = column, 16 samples ----
                    {
                        const int64_t v0 = row0 >> 3;
                        const int col = wg * TM + r, cl = col & 31;
                        const float* sc = stage + col * TM;
#pragma unroll 4
                        for (int si = 0; si < 16; ++si) {
                            const float4 a = *reinterpret_cast<const float4*>(sc + (((2 * si) ^ cl) << 2));
                            const float4 bq = *reinterpret_cast<const float4*>(sc + (((2 * si + 1) ^ cl) << 2));
                            const float sum = ((a.x + a.y) + (a.z + a.w)) + ((bq.x + bq.y) + (bq.z + bq.w));
                            if (v0 + si < A.Nv) A.X5[(v0 + si) * X5_W + col] = sum;
                        }
                    }
                    named_barrier<1, NEPI>();           // staging consumed: epilogue 0 of the next tile may write the activation
                    if (lane == 0) mbar_arrive(bar_tiledone);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// bytes of the packed weight image: 67 chunks x [hi 8 KB | lo 8 KB] in consumption order (host packer: mlp_tc.py)
extern "C" int64_t hnr_nbr_mlp_f16_packed_bytes(void) { return (int64_t)NCHUNK_TILE * W_STAGE; }

// Fused per-neighbour MLP + density head + weighted K-sum for Nv valid samples (K == 8), 3xFP16 on tcgen05.
// mul[l]: accumulator -> pre-activation factor of layer l (includes the next layer's input scale for l < 3);
// bias (4,256): rows 0..2 pre-multiplied by the next layer's input scale; inv_act = 1 / that scale.  dbg (optional):
// (4, Nv*8, 256) unscaled activations of the four layers, araw (optional, with dbg): (Nv*8) density pre-activations --
// the training forward saves them for the backward pass.
static int nbr_mlp_f16_launch(F16Args& A, int mode, void* stream) {
    static bool configured = false;
    if (!configured) {
        HNR_CUDA(cudaFuncSetAttribute(nbr_mlp_f16_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        HNR_CUDA(cudaFuncSetAttribute(nbr_mlp_f16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        HNR_CUDA(cudaFuncSetAttribute(nbr_mlp_f16_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const int64_t ntiles = hnr_cdiv(A.Nv * 8, TM);
    const int grid = (int)(ntiles < HNR_NUM_SMS ? ntiles : HNR_NUM_SMS);
    if (mode == 2) nbr_mlp_f16_kernel<2><<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(A);
    else if (mode == 1) nbr_mlp_f16_kernel<1><<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(A);
    else nbr_mlp_f16_kernel<0><<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(A);
    HNR_CHECK_LAUNCH("nbr_mlp_f16_forward");
    return HNR_OK;
}

extern "C" int hnr_nbr_mlp_f16_forward(const float* xyz, const float* xyz_pers, const float* emb, const float* color, const float* dir,
                                       const int32_t* pidx, const int32_t* vlist, const float* loc_w, const float* loc_pers,
                                       const float* raydirs, const float* cam, const float* weight, const float* confc, const void* wpack,
                                       const float* bias, const float* walpha, const float* balpha, const float* mul, float scale0,
                                       float scale2, float inv_act, int64_t Nv, int64_t K, float* sigma, float* X5, float* dbg, float* araw,
                                       int32_t* status, void* stream) {
    HNR_CHECK_ARG(K == 8, "nbr_mlp_f16_forward: K must be 8 (128-row tiles hold 16 whole samples)");
    if (Nv == 0) return HNR_OK;
    F16Args A{};
    A.xyz = xyz; A.xyz_pers = xyz_pers; A.emb = emb; A.color = color; A.dir = dir; A.pidx = pidx; A.vlist = vlist;
    A.loc_w = loc_w; A.loc_pers = loc_pers; A.raydirs = raydirs; A.cam = cam; A.weight = weight; A.confc = confc;
    A.wpack = (const uint8_t*)wpack; A.bias = bias; A.walpha = walpha; A.balpha = balpha; A.sigma = sigma; A.X5 = X5; A.dbg = dbg;
    A.Nv = Nv;
    for (int l = 0; l < NLAYER; ++l) A.mul[l] = mul[l];
    A.scale0 = scale0; A.scale2 = scale2; A.inv_act = inv_act; A.araw = araw; A.status = status;
    A.nc0 = NC0;
    return nbr_mlp_f16_launch(A, dbg ? 1 : 0, stream);
}

// Inference with the per-point layer-0 partial (see F16Args::pp): pp (N_points, 256) fp32 = block1[0].weight[:, :224] . [emb | PE(emb)]
// of every point (unscaled), e.g. from hnr_linear_tc_fwd over the point table; wpack = the SAME image as for hnr_nbr_mlp_f16_forward
// (the kernel starts at the distance-encoding chunks of layer 0).  The embedding table is not read.
extern "C" int hnr_nbr_mlp_f16_forward_pp(const float* xyz, const float* xyz_pers, const float* pp, const float* color, const float* dir,
                                          const int32_t* pidx, const int32_t* vlist, const float* loc_w, const float* loc_pers,
                                          const float* raydirs, const float* cam, const float* weight, const float* confc, const void* wpack,
                                          const float* bias, const float* walpha, const float* balpha, const float* mul, float scale0,
                                          float scale2, float inv_act, int64_t Nv, int64_t K, float* sigma, float* X5, int32_t* status,
                                          void* stream) {
    HNR_CHECK_ARG(K == 8, "nbr_mlp_f16_forward_pp: K must be 8 (128-row tiles hold 16 whole samples)");
    HNR_CHECK_ARG(pp != nullptr && (reinterpret_cast<uintptr_t>(pp) & 15) == 0, "nbr_mlp_f16_forward_pp: 16-byte aligned per-point partial required");
    if (Nv == 0) return HNR_OK;
    F16Args A{};
    A.xyz = xyz; A.xyz_pers = xyz_pers; A.emb = nullptr; A.color = color; A.dir = dir; A.pidx = pidx; A.vlist = vlist;
    A.loc_w = loc_w; A.loc_pers = loc_pers; A.raydirs = raydirs; A.cam = cam; A.weight = weight; A.confc = confc;
    constexpr int NC0_PP = 4;                                  // the distance-encoding chunks are the last 4 of layer 0's 18
    A.wpack = (const uint8_t*)wpack + (size_t)(NC0 - NC0_PP) * W_STAGE; A.bias = bias; A.walpha = walpha; A.balpha = balpha;
    A.sigma = sigma; A.X5 = X5; A.dbg = nullptr;
    A.Nv = Nv;
    for (int l = 0; l < NLAYER; ++l) A.mul[l] = mul[l];
    A.scale0 = scale0; A.scale2 = scale2; A.inv_act = inv_act; A.araw = nullptr; A.status = status;
    A.pp = pp; A.pp_scale = 1.f / inv_act; A.nc0 = NC0_PP;
    return nbr_mlp_f16_launch(A, 0, stream);
}

// Training forward: same arithmetic, and everything the fused backward needs is saved as split images (img_common.cuh; every
// image has ceil(Nv*8 / 128) * 128 rows): x0img 288 columns (layer-0 input in kernel column order), eimg 16 columns (block3
// extras [colour 3, dir - view 3, <dir, view>, 9 x 0]), himg[0..3] 256 columns (outputs of the four layers); araw (Nv*8) density
// pre-activations.  Consumers: hnr_alpha_ksum_bwd_img, hnr_nbr_bwd_f16, hnr_wgrad_img.
extern "C" int hnr_nbr_mlp_f16_forward_train(const float* xyz, const float* xyz_pers, const float* emb, const float* color, const float* dir,
                                             const int32_t* pidx, const int32_t* vlist, const float* loc_w, const float* loc_pers,
                                             const float* raydirs, const float* cam, const float* weight, const float* confc,
                                             const void* wpack, const float* bias, const float* walpha, const float* balpha, const float* mul,
                                             float scale0, float scale2, float inv_act, int64_t Nv, int64_t K, float* sigma, float* X5,
                                             void* x0img, void* eimg, void* h0img, void* h1img, void* h2img, void* h3img, float* araw,
                                             int32_t* status, void* stream) {
    HNR_CHECK_ARG(K == 8, "nbr_mlp_f16_forward_train: K must be 8 (128-row tiles hold 16 whole samples)");
    HNR_CHECK_ARG(x0img && eimg && h0img && h1img && h2img && h3img && araw, "nbr_mlp_f16_forward_train: every image is required");
    if (Nv == 0) return HNR_OK;
    F16Args A{};
    A.xyz = xyz; A.xyz_pers = xyz_pers; A.emb = emb; A.color = color; A.dir = dir; A.pidx = pidx; A.vlist = vlist;
    A.loc_w = loc_w; A.loc_pers = loc_pers; A.raydirs = raydirs; A.cam = cam; A.weight = weight; A.confc = confc;
    A.wpack = (const uint8_t*)wpack; A.bias = bias; A.walpha = walpha; A.balpha = balpha; A.sigma = sigma; A.X5 = X5; A.dbg = nullptr;
    A.Nv = Nv;
    for (int l = 0; l < NLAYER; ++l) A.mul[l] = mul[l];
    A.scale0 = scale0; A.scale2 = scale2; A.inv_act = inv_act; A.araw = araw; A.status = status;
    A.x0img = (uint8_t*)x0img; A.eimg = (uint8_t*)eimg;
    A.himg[0] = (uint8_t*)h0img; A.himg[1] = (uint8_t*)h1img; A.himg[2] = (uint8_t*)h2img; A.himg[3] = (uint8_t*)h3img;
    A.inv_scale0 = 1.f / scale0; A.inv_scale2 = 1.f / scale2;
    A.nc0 = NC0;
    return nbr_mlp_f16_launch(A, 2, stream);
}
