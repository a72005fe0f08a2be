// Fused per-neighbour MLP on the 5th-gen tensor cores (SURVEY.md §8a rows G1 + A1-A3).
//
// One persistent CTA per SM walks tiles of 128 neighbour rows (16 shading samples x K=8).  For a
// tile it gathers the neighbours' attributes straight from the point tables, then runs
//     x0(284) -> 256 -> 256 -> [+colour, dir-view, <dir,view>](263) -> 256 -> 256 -> density head
// and the inverse-distance-weighted K-sum without the activations ever leaving the SM:
//   * accumulators live in TMEM (128 lanes x 256 fp32 columns), written by tcgen05.mma;
//   * the layer input is kept in shared memory as fp32 (128 x 256) and re-split on the fly into
//     TF32 "hi" and "lo" parts, K-chunk by K-chunk, straight into the canonical UMMA layout;
//     the layer-1 input (embedding, its positional encoding, the encoded distances) is never
//     materialised at all -- each chunk is generated from 38 floats per row;
//   * fp32 accuracy on TF32 tensor cores: every product a*w is issued as three MMAs
//     a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (3xTF32; the dropped lo*lo term is ~2^-22 relative);
//   * weights are pre-split and pre-tiled on the host into the exact shared-memory image of one
//     K-chunk (hi then lo, canonical no-swizzle K-major core matrices), so a chunk is ONE 16 KB
//     cp.async.bulk (TMA bulk copy, completes on an mbarrier) from L2;
//   * warp roles: 8 worker warps (gather, operand conversion, TMEM->register epilogue: bias,
//     LeakyReLU, density dot product, weighted K-sum), 1 MMA-issuing warp (one elected thread),
//     1 TMA warp; a 3-stage mbarrier ring between them.
//
// Replaces, for inference, nbr_features + 4 x linear_fwd + alpha_ksum_fwd (aggregate.cu,
// linear_simt.cu); the arithmetic it restates is models/aggregators/point_aggregators.py:921-972,
// :1002-1036 and models/helpers/networks.py:175-189 of the reference.
#include "common.cuh"
#include "hnr.h"
#define TRACE_SRC ((long long*)nullptr)
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int TM = 128;                 // rows (neighbours) per tile
constexpr int HID = 256;                // layer width == UMMA N
constexpr int KC = 8;                   // K elements per chunk == one tf32 MMA K-step (32 bytes)
constexpr int NSTAGE = 3;
constexpr int W_PART = HID * KC * 4;    // 8192 B: hi (or lo) weights of one chunk
constexpr int A_PART = TM * KC * 4;     // 4096 B
constexpr int STAGE_BYTES = 2 * W_PART + 2 * A_PART;   // 24576
constexpr int ACT_LD = 260;             // fp32 activation row stride (bank-conflict-free float4 rows)
constexpr int EMB_LD = 36;
constexpr int NWORKER = 256;
constexpr int NTHREADS = 320;           // 8 worker warps + MMA warp + TMA warp
constexpr int NLAYER = 4;
constexpr int FEAT = 32, NF_FEAT = 3, NF_DIST = 5, NF_VIEW = 4, X5_W = 280;

// shared memory map (bytes)
constexpr int OFF_STAGE = 0;
constexpr int OFF_ACT = OFF_STAGE + NSTAGE * STAGE_BYTES;            // 73728
constexpr int OFF_E = OFF_ACT + TM * ACT_LD * 4;                     // +133120
constexpr int OFF_WC = OFF_E + TM * 8 * 4;
constexpr int OFF_ARAW = OFF_WC + TM * 4;
constexpr int OFF_BIAS = OFF_ARAW + 2 * TM * 4;
constexpr int OFF_WALPHA = OFF_BIAS + NLAYER * HID * 4;
constexpr int OFF_BAR = OFF_WALPHA + HID * 4;
constexpr int SMEM_BYTES = OFF_BAR + 128;
// aliases inside the activation buffer (only live while layer 1 is being fed)
constexpr int OFF_EMB = OFF_ACT;
constexpr int OFF_DIST = OFF_EMB + TM * EMB_LD * 4;
static_assert(OFF_DIST + TM * 8 * 4 <= OFF_E, "alias overflow");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct TcParams {
    int nchunk[NLAYER];         // K chunks per layer
    int64_t w_off[NLAYER];      // byte offset of the layer's first chunk in the packed weights
    int nlayer;
};

// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=256
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr uint32_t A_LBO = (TM / 8) * 128, W_LBO = (HID / 8) * 128, SBO = 128;

__device__ __forceinline__ void store_split(uint8_t* stage, int r, int h, float4 v) {
    float4 hi, lo;
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo.x = v.x - hi.x;
    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo.y = v.y - hi.y;
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo.z = v.z - hi.z;
    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo.w = v.w - hi.w;
    uint8_t* a = stage + 2 * W_PART + h * (int)A_LBO + r * 16;
    *reinterpret_cast<float4*>(a) = hi;
    *reinterpret_cast<float4*>(a + A_PART) = lo;
}

__device__ __forceinline__ void rot3(const float* m, float x, float y, float z, float& ox, float& oy, float& oz) {
    ox = x * m[0] + y * m[3] + z * m[6];
    oy = x * m[1] + y * m[4] + z * m[7];
    oz = x * m[2] + y * m[5] + z * m[8];
}
__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }

struct TcArgs {
    // point tables + per-sample inputs (fused mode)
    const float *xyz, *xyz_pers, *emb, *color, *dir;
    const int32_t *pidx, *vlist;
    const float *loc_w, *loc_pers, *raydirs, *cam, *weight, *confc;
    const uint8_t* wpack;          // packed split weights
    const float *bias, *walpha, *balpha;   // bias: NLAYER x 256
    float *sigma, *X5;
    int64_t Nv;
    int K;
    // test mode
    const float* a_test;           // (rows, ldA) fp32
    float* out_test;               // (rows, 256) raw accumulators
    int64_t rows_test;
    int ld_test;
};

template <bool TEST>
__global__ void __launch_bounds__(NTHREADS, 1) mlp_tc_kernel(TcArgs A, TcParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NSTAGE), bar_acc = smem_u32(bars + 2 * NSTAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 1);
    float* act = reinterpret_cast<float*>(smem + OFF_ACT);
    float* emb_s = reinterpret_cast<float*>(smem + OFF_EMB);
    float* dist_s = reinterpret_cast<float*>(smem + OFF_DIST);
    float* E_s = reinterpret_cast<float*>(smem + OFF_E);
    float* wc_s = reinterpret_cast<float*>(smem + OFF_WC);
    float* araw_s = reinterpret_cast<float*>(smem + OFF_ARAW);
    float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
    float* walpha_s = reinterpret_cast<float*>(smem + OFF_WALPHA);

    const int64_t total_rows = TEST ? A.rows_test : A.Nv * A.K;
    const int64_t ntiles = (total_rows + TM - 1) / TM;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(bar_full + 8 * s, 8 + 1);     // 8 worker warps + the TMA thread's expect_tx arrive
            mbar_init(bar_empty + 8 * s, 1);        // one tcgen05.commit
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {    // TMEM allocation is a warp-wide instruction
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (!TEST && tid < NWORKER) {
        for (int i = tid; i < NLAYER * HID; i += NWORKER) bias_s[i] = A.bias[i];
        walpha_s[tid] = A.walpha[tid];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int nlayer = P.nlayer;

    if (warp == 9) {
        // ================= TMA producer: one 16 KB bulk copy per K chunk =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int l = 0; l < nlayer; ++l) {
                    const uint8_t* src = A.wpack + P.w_off[l];
                    for (int c = 0; c < P.nchunk[l]; ++c, ++it) {
                        const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1;
                        mbar_wait(bar_empty + 8 * s, ph ^ 1);
                        mbar_arrive_expect_tx(bar_full + 8 * s, 2 * W_PART);
                        bulk_g2s(smem_u32(smem + OFF_STAGE + s * STAGE_BYTES), src + (size_t)c * 2 * W_PART, 2 * W_PART, bar_full + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 8) {
        // ================= MMA issuer: one thread, 3 MMAs (3xTF32) per chunk =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int l = 0; l < nlayer; ++l) {
                    for (int c = 0; c < P.nchunk[l]; ++c, ++it) {
                        const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1;
                        mbar_wait(bar_full + 8 * s, ph);
                        tc_fence_after();
                        const uint32_t st = smem_u32(smem + OFF_STAGE + s * STAGE_BYTES);
                        const uint64_t w_hi = umma_desc(st, W_LBO, SBO), w_lo = umma_desc(st + W_PART, W_LBO, SBO);
                        const uint64_t a_hi = umma_desc(st + 2 * W_PART, A_LBO, SBO), a_lo = umma_desc(st + 2 * W_PART + A_PART, A_LBO, SBO);
                        tc_mma_tf32(tmem_base, a_hi, w_hi, IDESC, c > 0 ? 1u : 0u);
                        tc_mma_tf32(tmem_base, a_lo, w_hi, IDESC, 1u);
                        tc_mma_tf32(tmem_base, a_hi, w_lo, IDESC, 1u);
                        tc_commit(bar_empty + 8 * s);                 // frees the stage when these MMAs retire
                        if (c == P.nchunk[l] - 1) tc_commit(bar_acc); // accumulator of this layer complete
                    }
                }
            }
        }
    } else {
        // ================= workers: gather, operand conversion, epilogue =================
        const int r = tid & (TM - 1), h = tid >> 7;          // conversion role: row, K half of the chunk
        const int erow = 32 * (warp & 3) + lane, ehalf = warp >> 2;   // epilogue role: TMEM lane, column half
        uint32_t it = 0, acc_it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t row0 = tile * TM;
            float view[3] = {0.f, 0.f, 0.f};
            if (!TEST) {
                // ---- gather this tile's neighbours (row r; both halves share the address math) ----
                const int64_t row = row0 + r;
                const bool live = row < total_rows;
                const int64_t v = live ? row / A.K : 0;
                const int k = live ? (int)(row - v * A.K) : 0;
                const int64_t s = A.vlist[live ? v : 0];
                const int64_t g = live ? max(A.pidx[s * A.K + k], 0) : 0;
                const float4* e4 = reinterpret_cast<const float4*>(A.emb + g * FEAT) + h * 4;
                float4* d4 = reinterpret_cast<float4*>(emb_s + r * EMB_LD + h * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) d4[i] = e4[i];
                const float* rt = A.cam + 12;
                if (h == 0) {
                    const float px = A.xyz[g * 3], py = A.xyz[g * 3 + 1], pz = A.xyz[g * 3 + 2];
                    float d[8];
                    rot3(rt, px - A.loc_w[s * 3], py - A.loc_w[s * 3 + 1], pz - A.loc_w[s * 3 + 2], d[0], d[1], d[2]);
                    float qx, qy, qz;
                    if (A.xyz_pers) {
                        qx = A.xyz_pers[g * 3]; qy = A.xyz_pers[g * 3 + 1]; qz = A.xyz_pers[g * 3 + 2];
                    } else {
                        float cx, cy, cz;
                        rot3(A.cam + 3, px - A.cam[0], py - A.cam[1], pz - A.cam[2], cx, cy, cz);
                        qx = cx / cz; qy = cy / cz; qz = cz;
                    }
                    const float sx = A.loc_pers[s * 3], sy = A.loc_pers[s * 3 + 1], sz = A.loc_pers[s * 3 + 2];
                    d[3] = qx * qz - sx * sz; d[4] = qy * qz - sy * sz; d[5] = qz - sz; d[6] = 0.f; d[7] = 0.f;
                    float4* o = reinterpret_cast<float4*>(dist_s + r * 8);
                    o[0] = make_float4(d[0], d[1], d[2], d[3]);
                    o[1] = make_float4(d[4], d[5], 0.f, 0.f);
                } else {
                    float vx, vy, vz, rx, ry, rz;
                    rot3(rt, A.raydirs[s * 3], A.raydirs[s * 3 + 1], A.raydirs[s * 3 + 2], vx, vy, vz);
                    rot3(rt, A.dir[g * 3], A.dir[g * 3 + 1], A.dir[g * 3 + 2], rx, ry, rz);
                    float4* o = reinterpret_cast<float4*>(E_s + r * 8);
                    o[0] = make_float4(A.color[g * 3], A.color[g * 3 + 1], A.color[g * 3 + 2], rx - vx);
                    o[1] = make_float4(ry - vy, rz - vz, rx * vx + ry * vy + rz * vz, 0.f);
                    wc_s[r] = live ? A.weight[s * A.K + k] * (A.confc ? A.confc[s * A.K + k] : 1.f) : 0.f;
                }
                named_bar<NWORKER>();
            }
            for (int l = 0; l < nlayer; ++l) {
                // ---- feed the layer: one (hi, lo) operand chunk per stage ----
                float sn[4], cs[4];
                for (int c = 0; c < P.nchunk[l]; ++c, ++it) {
                    const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    uint8_t* stage = smem + OFF_STAGE + s * STAGE_BYTES;
                    float4 val;
                    if (TEST) {
                        const int64_t row = row0 + r;
                        const int kk = c * KC + h * 4;
                        val = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (row < total_rows) {
                            const float* ap = A.a_test + row * A.ld_test;
                            val.x = kk + 0 < A.ld_test ? ap[kk + 0] : 0.f;
                            val.y = kk + 1 < A.ld_test ? ap[kk + 1] : 0.f;
                            val.z = kk + 2 < A.ld_test ? ap[kk + 2] : 0.f;
                            val.w = kk + 3 < A.ld_test ? ap[kk + 3] : 0.f;
                        }
                    } else if (l == 0) {
                        if (c < 4) {                                   // raw embedding channels
                            val = *reinterpret_cast<const float4*>(emb_s + r * EMB_LD + c * 8 + h * 4);
                        } else if (c < 28) {                           // sin/cos(2^f e): [c-block of 8][f][sin|cos][8 channels]
                            const int q = c - 4, cblk = q / 6, fs = q - cblk * 6;
                            if (fs == 0) {
                                const float4 e = *reinterpret_cast<const float4*>(emb_s + r * EMB_LD + cblk * 8 + h * 4);
                                sincosf(e.x, &sn[0], &cs[0]); sincosf(e.y, &sn[1], &cs[1]);
                                sincosf(e.z, &sn[2], &cs[2]); sincosf(e.w, &sn[3], &cs[3]);
                            } else if ((fs & 1) == 0) {                // next octave by angle doubling
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float s2 = 2.f * sn[i] * cs[i], c2 = 1.f - 2.f * sn[i] * sn[i];
                                    sn[i] = s2; cs[i] = c2;
                                }
                            }
                            val = (fs & 1) ? make_float4(cs[0], cs[1], cs[2], cs[3]) : make_float4(sn[0], sn[1], sn[2], sn[3]);
                        } else {                                       // encoded distances: idx = f*12 + (sin|cos)*6 + component
                            float o[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int idx = (c - 28) * 8 + h * 4 + i;
                                float res = 0.f;
                                if (idx < 2 * NF_DIST * 6) {
                                    const int f = idx / 12, rem = idx - f * 12, sc = rem / 6, j = rem - sc * 6;
                                    float a, b;
                                    sincosf(dist_s[r * 8 + j] * (float)(1 << f), &a, &b);
                                    res = sc ? b : a;
                                }
                                o[i] = res;
                            }
                            val = make_float4(o[0], o[1], o[2], o[3]);
                        }
                    } else if (l == 2 && c == 32) {                    // block3 extras (colour, dir-view, <dir,view>, 0)
                        val = *reinterpret_cast<const float4*>(E_s + r * 8 + h * 4);
                    } else {
                        val = *reinterpret_cast<const float4*>(act + r * ACT_LD + c * KC + h * 4);
                    }
                    store_split(stage, r, h, val);
                    fence_proxy_async();          // generic-proxy stores -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + 8 * s);
                }
                // ---- epilogue of the layer: TMEM -> registers -> next layer's input ----
                mbar_wait(bar_acc, acc_it & 1);
                ++acc_it;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + ehalf * 128;
                if (TEST) {
                    const int64_t row = row0 + erow;
#pragma unroll 1
                    for (int cb = 0; cb < 4; ++cb) {
                        float v[32];
                        tmem_ld32(taddr + cb * 32, v);
                        if (row < total_rows) {
                            float4* o = reinterpret_cast<float4*>(A.out_test + row * HID + ehalf * 128 + cb * 32);
#pragma unroll
                            for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        }
                    }
                } else {
                    const bool last = (l == nlayer - 1);
                    const float wrow = last ? wc_s[erow] : 1.f;
                    float dot = 0.f;
#pragma unroll 1
                    for (int cb = 0; cb < 4; ++cb) {
                        float v[32];
                        tmem_ld32(taddr + cb * 32, v);
                        const int col0 = ehalf * 128 + cb * 32;
                        float4* o = reinterpret_cast<float4*>(act + erow * ACT_LD + col0);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float y[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float t = v[4 * i + j] + bias_s[l * HID + col0 + 4 * i + j];
                                t = t > 0.f ? t : 0.01f * t;
                                if (last) { dot = fmaf(t, walpha_s[col0 + 4 * i + j], dot); t *= wrow; }
                                y[j] = t;
                            }
                            o[i] = make_float4(y[0], y[1], y[2], y[3]);
                        }
                    }
                    if (last) araw_s[ehalf * TM + erow] = dot;
                }
                tc_fence_before();
                named_bar<NWORKER>();           // every row of the new activation is in place / TMEM drained
            }
            if (!TEST) {
                // ---- weighted K-sum over each sample's 8 neighbour rows + density head + view encoding ----
                const int nsamp = TM / A.K;
                const int64_t v0 = row0 / A.K;
                for (int sidx = 0; sidx < nsamp; ++sidx) {
                    const int64_t v = v0 + sidx;
                    if (v >= A.Nv) break;
                    float acc = 0.f;
                    for (int k = 0; k < A.K; ++k) acc += act[(sidx * A.K + k) * ACT_LD + tid];
                    A.X5[v * X5_W + tid] = acc;
                }
                if (tid < nsamp && v0 + tid < A.Nv) {
                    float sg = 0.f;
                    for (int k = 0; k < A.K; ++k) {
                        const int rr = tid * A.K + k;
                        const float raw = araw_s[rr] + araw_s[TM + rr] + A.balpha[0];
                        sg += wc_s[rr] * softplus_t(raw - 1.f);
                    }
                    A.sigma[v0 + tid] = sg;
                }
                if (tid >= 32 && tid < 32 + nsamp * 3 * NF_VIEW) {
                    const int idx = tid - 32, sidx = idx / (3 * NF_VIEW), q = idx - sidx * (3 * NF_VIEW);
                    const int64_t v = v0 + sidx;
                    if (v < A.Nv) {
                        const int64_t s = A.vlist[v];
                        float vx, vy, vz;
                        rot3(A.cam + 12, A.raydirs[s * 3], A.raydirs[s * 3 + 1], A.raydirs[s * 3 + 2], vx, vy, vz);
                        const int c = q / NF_VIEW, f = q - c * NF_VIEW;
                        const float val = c == 0 ? vx : (c == 1 ? vy : vz);
                        float a, b;
                        sincosf(val * (float)(1 << f), &a, &b);
                        A.X5[v * X5_W + HID + q] = a;
                        A.X5[v * X5_W + HID + 3 * NF_VIEW + q] = b;
                    }
                }
                named_bar<NWORKER>();           // act / wc / araw are reused by the next tile's gather
            }
            (void)view;
        }
    }
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

int launch_cfg_done = 0;

template <bool TEST>
int launch(const TcArgs& a, const TcParams& p, int64_t ntiles, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(mlp_tc_kernel<TEST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) {
            hnr_set_error(cudaGetErrorString(e));
            return HNR_ERR_CUDA;
        }
        configured = true;
    }
    int grid = (int)(ntiles < HNR_NUM_SMS ? ntiles : HNR_NUM_SMS);
    mlp_tc_kernel<TEST><<<grid, NTHREADS, SMEM_BYTES, st>>>(a, p);
    return HNR_OK;
}

}  // namespace

// bytes of the packed image of a layer with Kp (multiple of 8) input columns
extern "C" int64_t hnr_mlp_tc_packed_bytes(int64_t Kp) { return (Kp / KC) * 2 * (int64_t)W_PART; }

// Single-layer self test of the tensor-core machinery: out (rows,256) = A (rows, ldA>=K) . Wpacked^T, raw
// fp32 accumulators (no bias / activation).  Wpacked = image produced by the host packer for Kp columns.
extern "C" int hnr_mlp_tc_gemm_test(const float* a, int64_t rows, int64_t ld, const void* wpack, int64_t Kp, float* out, void* stream) {
    HNR_CHECK_ARG(Kp > 0 && Kp % KC == 0, "mlp_tc_gemm_test: Kp must be a multiple of 8");
    if (rows == 0) return HNR_OK;
    TcArgs A{};
    A.wpack = (const uint8_t*)wpack;
    A.a_test = a; A.out_test = out; A.rows_test = rows; A.ld_test = (int)ld;
    A.K = 8;
    TcParams P{};
    P.nlayer = 1; P.nchunk[0] = (int)(Kp / KC); P.w_off[0] = 0;
    int rc = launch<true>(A, P, hnr_cdiv(rows, TM), (cudaStream_t)stream);
    if (rc) return rc;
    HNR_CHECK_LAUNCH("mlp_tc_gemm_test");
    return HNR_OK;
}

// Fused per-neighbour MLP + density head + weighted K-sum for Nv valid samples (K must be 8).
// wpack: the four layers' packed images back to back (288, 256, 264, 256 input columns, in the K order
// documented in hybridneuralrendering_b200/mlp_tc.py); bias (4,256); outputs sigma (Nv), X5 (Nv,280).
extern "C" int hnr_mlp_tc_forward(const float* xyz, const float* xyz_pers, const float* emb, const float* color, const float* dir,
                                  const int32_t* pidx, const int32_t* vlist, const float* loc_w, const float* loc_pers,
                                  const float* raydirs, const float* cam, const float* weight, const float* confc, const void* wpack,
                                  const float* bias, const float* walpha, const float* balpha, int64_t Nv, int64_t K, float* sigma,
                                  float* X5, void* stream) {
    HNR_CHECK_ARG(K == 8, "mlp_tc_forward: K must be 8 (128-row tiles hold 16 whole samples)");
    if (Nv == 0) return HNR_OK;
    TcArgs A{};
    A.xyz = xyz; A.xyz_pers = xyz_pers; A.emb = emb; A.color = color; A.dir = dir; A.pidx = pidx; A.vlist = vlist;
    A.loc_w = loc_w; A.loc_pers = loc_pers; A.raydirs = raydirs; A.cam = cam; A.weight = weight; A.confc = confc;
    A.wpack = (const uint8_t*)wpack; A.bias = bias; A.walpha = walpha; A.balpha = balpha; A.sigma = sigma; A.X5 = X5;
    A.Nv = Nv; A.K = (int)K;
    TcParams P{};
    const int kp[NLAYER] = {288, 256, 264, 256};
    int64_t off = 0;
    P.nlayer = NLAYER;
    for (int l = 0; l < NLAYER; ++l) { P.nchunk[l] = kp[l] / KC; P.w_off[l] = off; off += hnr_mlp_tc_packed_bytes(kp[l]); }
    int rc = launch<false>(A, P, hnr_cdiv(Nv * K, TM), (cudaStream_t)stream);
    if (rc) return rc;
    HNR_CHECK_LAUNCH("mlp_tc_forward");
    return HNR_OK;
}
