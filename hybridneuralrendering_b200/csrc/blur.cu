// Blur-handling module: pre-defined degradation kernels applied per patch,
// border-renormalised, best candidate per patch chosen by L1 against the ground truth.
// Replaces BaseRenderingModel.blur_update_output (models/base_rendering_model.py:677-786).
//
// One CTA per patch; a thread per patch pixel holds its 3 channels.  The patch (3 x PS x PS) and
// the current kernel are staged in shared memory; all Nk+1 candidates are evaluated in-CTA, so the
// only HBM traffic is the patch, its GT, the kernels (L2 resident) and the result.
#include "common.cuh"

namespace {

constexpr int MAX_PS = 16;   // patch edge
constexpr int MAX_KS = 15;   // blur kernel edge

template <bool BWD>
__global__ void blur_kernel(const float* __restrict__ pred,      // (S*S,3) raster
                            const float* __restrict__ gt,        // (S*S,3)            (fwd)
                            const float* __restrict__ kernels,   // (Nk,KS,KS)
                            const float* __restrict__ g_out,     // (S*S,3)            (bwd)
                            int32_t* __restrict__ select,        // (PN*PN)  out (fwd) / in (bwd)
                            float* __restrict__ out,             // (S*S,3) fwd result / bwd grad wrt pred
                            int PN, int PS, int Nk, int KS) {
    __shared__ float sp[3][MAX_PS][MAX_PS];
    __shared__ float sk[MAX_KS][MAX_KS];
    __shared__ float red[32];
    __shared__ float s_err;
    const int patch = blockIdx.x;
    const int pi = patch / PN, pj = patch % PN;
    const int S = PN * PS;
    const int tid = threadIdx.x;
    const int npix = PS * PS;
    const int y = tid / PS, x = tid % PS;
    const bool act = tid < npix;
    const int pad = KS / 2;
    const size_t gidx = act ? ((size_t)(pi * PS + y) * S + (pj * PS + x)) * 3 : 0;

    if (!BWD) {
        float g[3] = {0.f, 0.f, 0.f}, me[3] = {0.f, 0.f, 0.f};
        if (act) {
            for (int c = 0; c < 3; ++c) { me[c] = pred[gidx + c]; sp[c][y][x] = me[c]; g[c] = gt[gidx + c]; }
        }
        __syncthreads();
        float best_err = INFINITY; int best = Nk;
        float bestv[3] = {me[0], me[1], me[2]};
        for (int n = 0; n <= Nk; ++n) {
            float v[3] = {me[0], me[1], me[2]};
            if (n < Nk) {
                __syncthreads();
                for (int i = tid; i < KS * KS; i += blockDim.x) sk[i / KS][i % KS] = kernels[(size_t)n * KS * KS + i];
                __syncthreads();
                if (act) {
                    float acc[3] = {0.f, 0.f, 0.f}, norm = 0.f;
                    for (int ky = 0; ky < KS; ++ky) {
                        int yy = y + ky - pad;
                        if (yy < 0 || yy >= PS) continue;
                        for (int kx = 0; kx < KS; ++kx) {
                            int xx = x + kx - pad;
                            if (xx < 0 || xx >= PS) continue;
                            float w = sk[ky][kx];
                            norm += w;
                            acc[0] += w * sp[0][yy][xx]; acc[1] += w * sp[1][yy][xx]; acc[2] += w * sp[2][yy][xx];
                        }
                    }
                    v[0] = acc[0] / norm; v[1] = acc[1] / norm; v[2] = acc[2] / norm;
                }
            }
            float e = act ? fabsf(v[0] - g[0]) + fabsf(v[1] - g[1]) + fabsf(v[2] - g[2]) : 0.f;
            e = warp_sum(e);
            if ((tid & 31) == 0) red[tid >> 5] = e;
            __syncthreads();
            if (tid == 0) {
                float t = 0.f;
                for (int w = 0; w < (int)((blockDim.x + 31) / 32); ++w) t += red[w];
                s_err = t;
            }
            __syncthreads();
            float err = s_err;
            // strict <: first minimum wins, as torch.argmin; a NaN error (kernel whose border
            // normalisation is 0/0) counts as the minimum, again as torch.argmin does
            if (err < best_err || (err != err && best_err == best_err)) {
                best_err = err; best = n;
                bestv[0] = v[0]; bestv[1] = v[1]; bestv[2] = v[2];
            }
        }
        if (act) { out[gidx] = bestv[0]; out[gidx + 1] = bestv[1]; out[gidx + 2] = bestv[2]; }
        if (tid == 0) select[patch] = best;
    } else {
        const int n = select[patch];
        if (n >= Nk) {
            if (act) for (int c = 0; c < 3; ++c) out[gidx + c] = g_out[gidx + c];
            return;
        }
        for (int i = tid; i < KS * KS; i += blockDim.x) sk[i / KS][i % KS] = kernels[(size_t)n * KS * KS + i];
        __syncthreads();
        if (act) {
            float norm = 0.f;
            for (int ky = 0; ky < KS; ++ky) {
                int yy = y + ky - pad;
                if (yy < 0 || yy >= PS) continue;
                for (int kx = 0; kx < KS; ++kx) {
                    int xx = x + kx - pad;
                    if (xx < 0 || xx >= PS) continue;
                    norm += sk[ky][kx];
                }
            }
            for (int c = 0; c < 3; ++c) sp[c][y][x] = g_out[gidx + c] / norm;   // upstream / border norm
        }
        __syncthreads();
        if (act) {
            // d pred[y,x] = sum over outputs (oy,ox) that read (y,x): tap (y-oy+pad, x-ox+pad)
            float acc[3] = {0.f, 0.f, 0.f};
            for (int oy = 0; oy < PS; ++oy) {
                int ky = y - oy + pad;
                if (ky < 0 || ky >= KS) continue;
                for (int ox = 0; ox < PS; ++ox) {
                    int kx = x - ox + pad;
                    if (kx < 0 || kx >= KS) continue;
                    float w = sk[ky][kx];
                    acc[0] += w * sp[0][oy][ox]; acc[1] += w * sp[1][oy][ox]; acc[2] += w * sp[2][oy][ox];
                }
            }
            out[gidx] = acc[0]; out[gidx + 1] = acc[1]; out[gidx + 2] = acc[2];
        }
    }
}

}  // namespace

extern "C" int hnr_blur_select_fwd(const float* pred, const float* gt, const float* kernels, int64_t patch_num, int64_t patch_size,
                                   int64_t num_kernels, int64_t kernel_size, float* out, int32_t* select, void* stream) {
    HNR_CHECK_ARG(patch_size > 0 && patch_size <= MAX_PS, "blur: patch_size must be in 1..16");
    HNR_CHECK_ARG(kernel_size > 0 && kernel_size <= MAX_KS && (kernel_size & 1), "blur: kernel_size must be odd and <= 15");
    HNR_CHECK_ARG(patch_num > 0 && num_kernels >= 0, "blur: bad shape");
    int threads = (int)(((patch_size * patch_size + 31) / 32) * 32);
    blur_kernel<false><<<(unsigned)(patch_num * patch_num), threads, 0, (cudaStream_t)stream>>>(
        pred, gt, kernels, nullptr, select, out, (int)patch_num, (int)patch_size, (int)num_kernels, (int)kernel_size);
    HNR_CHECK_LAUNCH("blur_select_fwd");
    return HNR_OK;
}

extern "C" int hnr_blur_select_bwd(const float* g_out, const float* kernels, const int32_t* select, int64_t patch_num,
                                   int64_t patch_size, int64_t num_kernels, int64_t kernel_size, float* g_pred, void* stream) {
    HNR_CHECK_ARG(patch_size > 0 && patch_size <= MAX_PS, "blur: patch_size must be in 1..16");
    HNR_CHECK_ARG(kernel_size > 0 && kernel_size <= MAX_KS && (kernel_size & 1), "blur: kernel_size must be odd and <= 15");
    int threads = (int)(((patch_size * patch_size + 31) / 32) * 32);
    blur_kernel<true><<<(unsigned)(patch_num * patch_num), threads, 0, (cudaStream_t)stream>>>(
        nullptr, nullptr, kernels, g_out, const_cast<int32_t*>(select), g_pred, (int)patch_num, (int)patch_size, (int)num_kernels,
        (int)kernel_size);
    HNR_CHECK_LAUNCH("blur_select_bwd");
    return HNR_OK;
}
