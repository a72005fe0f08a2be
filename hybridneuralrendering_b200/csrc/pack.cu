// One-launch (re)packing of every tensor-core weight image of the training step.
//
// After each optimiser step all MLP weights change, and every fused kernel wants them as split hi/lo chunk images in the
// canonical K-major UMMA layout: fp16 x scale for the forward kernels (nbr_mlp_f16.cu, chain_f16.cu), bf16 of the TRANSPOSE for
// the data-gradient chains (nbr_bwd_f16.cu, chain_bwd_f16.cu).  Doing that with tensor ops costs ~270 tiny launches (and, before
// round 2, four host synchronisations) per step -- more host time than the whole backward pass.  Here a static job table
// (built once per model on the host: mlp_tc.py / chain.py describe the layouts) drives ONE kernel:
//   thread = one 16-byte piece (8 consecutive reduction elements of one operand row): gathers its 8 source weights (optional
//   column map = the kernel's input-column order, -1 = zero padding), scales, splits into hi = rn(v), lo = rn(v - hi) and writes
//   both planes.  Extra blocks refresh the pre-scaled bias tables.
// Image of a job: red_p/16 chunks x [hi | lo], each plane [k block (2)][row group (rows_p/8)][row (8)][8 elements].
#include <cuda_fp16.h>

#include "common.cuh"
#include "hnr.h"
#include "img_common.cuh"

namespace {

struct PackJob {               // keep in sync with mlp_tc.PackJob (ctypes)
    const float* W;            // source weight (n_src, k_src), row stride ldw
    const int32_t* colmap;     // optional: kernel-order input column -> source column (-1 = zero); map_len entries, beyond = -1
    uint8_t* dst;
    int64_t piece0;            // global index of this job's first piece
    float scale;
    int32_t ldw, n_src, k_src;
    int32_t rows_p, red_p;     // padded operand rows (multiple of 8) / reduction length (multiple of 16)
    int32_t transpose;         // 0: operand row = output unit n, reduction = input column k (forward); 1: the transpose (data gradient)
    int32_t fmt;               // 0 = fp16 (saturation reported through the status word), 1 = bf16
    int32_t map_len;
    int32_t pad_;
};
struct BiasJob {
    const float* src;
    float* dst;
    float scale;
    int32_t n;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

__global__ void __launch_bounds__(256) pack_weights_kernel(const PackJob* __restrict__ jobs, int njobs, int64_t total_pieces,
                                                          const BiasJob* __restrict__ bjobs, int nbias, int nblk_pieces,
                                                          int32_t* __restrict__ status) {
    if ((int)blockIdx.x >= nblk_pieces) {
        const int b = blockIdx.x - nblk_pieces;
        if (b < nbias) {
            const BiasJob J = bjobs[b];
            for (int i = threadIdx.x; i < J.n; i += blockDim.x) J.dst[i] = J.src[i] * J.scale;
        }
        return;
    }
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total_pieces) return;
    int ji = 0;
    for (int q = 1; q < njobs; ++q)
        if (gid >= jobs[q].piece0) ji = q;
    const PackJob J = jobs[ji];
    const int p = (int)(gid - J.piece0);
    const int c = p / (2 * J.rows_p), rem = p - c * 2 * J.rows_p;
    const int kb = rem / J.rows_p, row = rem - kb * J.rows_p;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int red = c * 16 + kb * 8 + e;
        // (n, kcol): output unit and kernel-order input column of this element
        const int n = J.transpose ? red : row, kc = J.transpose ? row : red;
        int col = kc;
        if (J.colmap) col = kc < J.map_len ? J.colmap[kc] : -1;
        v[e] = (n < J.n_src && col >= 0 && col < J.k_src) ? J.W[(int64_t)n * J.ldw + col] * J.scale : 0.f;
    }
    uint4 hi, lo;
    if (J.fmt == 0) {
        uint32_t h[4], l[4];
        bool sat = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a = v[2 * i], b = v[2 * i + 1];
            sat |= fabsf(a) > 60000.f || fabsf(b) > 60000.f;
            h[i] = pack_h2(a, b);
            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
            l[i] = pack_h2(a - hf.x, b - hf.y);
        }
        if (sat && status) atomicOr(status, 4);
        hi = make_uint4(h[0], h[1], h[2], h[3]);
        lo = make_uint4(l[0], l[1], l[2], l[3]);
    } else {
        img::split8_bf16(v, hi, lo);
    }
    uint8_t* d = J.dst + (int64_t)c * J.rows_p * 64 + (int64_t)kb * J.rows_p * 16 + (int64_t)row * 16;
    *reinterpret_cast<uint4*>(d) = hi;
    *reinterpret_cast<uint4*>(d + (int64_t)J.rows_p * 32) = lo;
}

}  // namespace

extern "C" int64_t hnr_pack_job_bytes(void) { return (int64_t)sizeof(PackJob); }
extern "C" int64_t hnr_bias_job_bytes(void) { return (int64_t)sizeof(BiasJob); }

// Re-pack every weight image described by the device-resident job tables (see the header of this file) in one launch.
// jobs: njobs x PackJob with ascending piece0, total_pieces = sum over jobs of (red_p/16) * 2 * rows_p; bias_jobs: nbias x BiasJob
// (dst[i] = src[i] * scale).  status (optional): |= 4 when a scaled fp16 weight exceeds the format's range.
extern "C" int hnr_pack_weights(const void* jobs, int64_t njobs, int64_t total_pieces, const void* bias_jobs, int64_t nbias,
                                int32_t* status, void* stream) {
    HNR_CHECK_ARG(njobs >= 0 && njobs <= 64 && nbias >= 0, "pack_weights: bad job count");
    if (njobs == 0 && nbias == 0) return HNR_OK;
    const int nblk = (int)hnr_cdiv(total_pieces, 256);
    pack_weights_kernel<<<nblk + (int)nbias, 256, 0, (cudaStream_t)stream>>>((const PackJob*)jobs, (int)njobs, total_pieces,
                                                                            (const BiasJob*)bias_jobs, (int)nbias, nblk, status);
    HNR_CHECK_LAUNCH("pack_weights");
    return HNR_OK;
}
