// Learnable blur-kernel branch: the per-patch predicted degradation kernel applied to the rendered
// patch.  Replaces the body of BaseRenderingModel.learnable_blur_update_output
// (models/base_rendering_model.py:827-1020) around the predictor MLP (which runs through the dense-layer kernels):
//
//   blur_gray_*      patch raster -> predictor input rows [mean_c gt | mean_c pred]                 (:883-889)
//   blur_learn_*     raw predictor output (sigmoid) -> normalised kernel (/sum or softmax, :895-899), mode 4 mix with the
//                    identity kernel by the predicted weight + renormalisation (:904-909), grouped per-patch KSxKS
//                    cross-correlation with zero padding and the three boundary modes (:915-923), written back on the
//                    patch raster (:999-1005); backward w.r.t. the rendered patch and the raw predictor output.
//
// One CTA per patch.  Everything a patch needs (3 x PS x PS pixels, KS x KS taps) lives in shared memory; HBM traffic is
// the patch, the raw kernel row and the result: 24*PS^2 + 4*(KS^2+1) bytes forward.
#include "common.cuh"

namespace {

constexpr int MAX_PS = 16;
constexpr int MAX_KS = 15;
constexpr int MAX_KK = MAX_KS * MAX_KS;

// block-wide sum; every thread gets the result.  `red` holds >= 32 floats.
__device__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) t += red[w];
    return t;
}
__device__ float block_max(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = -INFINITY;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) t = fmaxf(t, red[w]);
    return t;
}

struct KernelBuild {
    float s;     // norm 0: sum of the raw taps
    float s2;    // mode 4: sum of the mixed kernel
    float wc;    // mode 4: combine weight
};

// raw (KK [+1]) -> k1 (normalised) and k (final) in shared memory
__device__ KernelBuild build_kernel(const float* __restrict__ raw, int KK, int centre, int norm_mode, int mix_mode, float* k1, float* k,
                                    float* red) {
    const int tid = threadIdx.x;
    KernelBuild kb{1.f, 1.f, 0.f};
    if (norm_mode == 0) {
        float a = 0.f;
        for (int i = tid; i < KK; i += blockDim.x) a += raw[i];
        kb.s = block_sum(a, red);
        for (int i = tid; i < KK; i += blockDim.x) k1[i] = raw[i] / kb.s;
    } else {
        float m = -INFINITY;
        for (int i = tid; i < KK; i += blockDim.x) m = fmaxf(m, raw[i]);
        m = block_max(m, red);
        float a = 0.f;
        for (int i = tid; i < KK; i += blockDim.x) { float e = expf(raw[i] - m); k1[i] = e; a += e; }
        kb.s = block_sum(a, red);
        for (int i = tid; i < KK; i += blockDim.x) k1[i] = k1[i] / kb.s;
    }
    __syncthreads();
    if (mix_mode == 4) {
        kb.wc = raw[KK];
        float a = 0.f;
        for (int i = tid; i < KK; i += blockDim.x) {
            float v = kb.wc * k1[i] + (1.f - kb.wc) * (i == centre ? 1.f : 0.f);
            k[i] = v; a += v;
        }
        kb.s2 = block_sum(a, red);
        for (int i = tid; i < KK; i += blockDim.x) k[i] = k[i] / kb.s2;
    } else {
        for (int i = tid; i < KK; i += blockDim.x) k[i] = k1[i];
    }
    __syncthreads();
    return kb;
}

template <bool BWD>
__global__ void blur_learn_kernel(const float* __restrict__ pred,     // (S*S,3) patch raster
                                  const float* __restrict__ raw,      // (N, KK [+1]) predictor output, row stride ld_raw
                                  int ld_raw,
                                  const float* __restrict__ g_out,    // (S*S,3)           (bwd)
                                  float* __restrict__ out,            // fwd: (S*S,3) result;  bwd: grad wrt pred
                                  float* __restrict__ g_raw,          // bwd: (N, ld_raw) grad wrt raw
                                  int PN, int PS, int KS, int norm_mode, int mix_mode, int boundary_mode) {
    __shared__ float sp[3][MAX_PS][MAX_PS];      // rendered patch
    __shared__ float sg[3][MAX_PS][MAX_PS];      // bwd: d conv
    __shared__ float sdm[MAX_PS][MAX_PS];        // bwd: d mask_out
    __shared__ float k1[MAX_KK], k[MAX_KK], dk[MAX_KK];
    __shared__ float red[32];
    const int patch = blockIdx.x;
    const int pi = patch / PN, pj = patch % PN;
    const int S = PN * PS;
    const int tid = threadIdx.x;
    const int npix = PS * PS, KK = KS * KS, pad = KS / 2, centre = pad * KS + pad;
    const int y = tid / PS, x = tid % PS;
    const bool act = tid < npix;
    const size_t gidx = act ? ((size_t)(pi * PS + y) * S + (pj * PS + x)) * 3 : 0;
    const float* rawp = raw + (size_t)patch * ld_raw;

    float me[3] = {0.f, 0.f, 0.f};
    if (act) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { me[c] = pred[gidx + c]; sp[c][y][x] = me[c]; }
    }
    KernelBuild kb = build_kernel(rawp, KK, centre, norm_mode, mix_mode, k1, k, red);   // syncs inside (covers sp)

    // cross-correlation with zero padding (F.conv2d) and the in-bounds tap mass (conv of the ones mask)
    float conv[3] = {0.f, 0.f, 0.f}, m = 0.f;
    if (act) {
        for (int ky = 0; ky < KS; ++ky) {
            int yy = y + ky - pad;
            if (yy < 0 || yy >= PS) continue;
            for (int kx = 0; kx < KS; ++kx) {
                int xx = x + kx - pad;
                if (xx < 0 || xx >= PS) continue;
                float w = k[ky * KS + kx];
                m += w;
                conv[0] += w * sp[0][yy][xx]; conv[1] += w * sp[1][yy][xx]; conv[2] += w * sp[2][yy][xx];
            }
        }
    }
    if (!BWD) {
        if (act) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
                out[gidx + c] = boundary_mode == 0 ? conv[c] / (m + 1e-10f) : conv[c] + (1.f - m) * me[c];
        }
        return;
    }

    // ---- backward ----
    float direct[3] = {0.f, 0.f, 0.f};           // gradient reaching pred without passing through the taps
    if (act) {
        float g[3] = {g_out[gidx], g_out[gidx + 1], g_out[gidx + 2]};
        float dm = 0.f;
        if (boundary_mode == 0) {
            float inv = 1.f / (m + 1e-10f);
#pragma unroll
            for (int c = 0; c < 3; ++c) { sg[c][y][x] = g[c] * inv; dm -= g[c] * conv[c] * inv * inv; }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) { sg[c][y][x] = g[c]; dm -= g[c] * me[c]; direct[c] = g[c] * (1.f - m); }
            if (boundary_mode == 2) dm = 0.f;    // mask_outputs computed from a detached kernel (:921-922)
        }
        sdm[y][x] = dm;
    }
    __syncthreads();
    // d pred[y,x] = direct + sum over outputs (oy,ox) that read (y,x) through tap (y-oy+pad, x-ox+pad)
    if (act) {
        float acc[3] = {direct[0], direct[1], direct[2]};
        for (int oy = 0; oy < PS; ++oy) {
            int ky = y - oy + pad;
            if (ky < 0 || ky >= KS) continue;
            for (int ox = 0; ox < PS; ++ox) {
                int kx = x - ox + pad;
                if (kx < 0 || kx >= KS) continue;
                float w = k[ky * KS + kx];
                acc[0] += w * sg[0][oy][ox]; acc[1] += w * sg[1][oy][ox]; acc[2] += w * sg[2][oy][ox];
            }
        }
        out[gidx] = acc[0]; out[gidx + 1] = acc[1]; out[gidx + 2] = acc[2];
    }
    // d k[ky,kx] = sum over output pixels whose tap is in bounds of (sum_c dconv_c * p_c(shifted) + dm)
    for (int t = tid; t < KK; t += blockDim.x) {
        int ky = t / KS, kx = t % KS;
        float a = 0.f;
        for (int oy = 0; oy < PS; ++oy) {
            int yy = oy + ky - pad;
            if (yy < 0 || yy >= PS) continue;
            for (int ox = 0; ox < PS; ++ox) {
                int xx = ox + kx - pad;
                if (xx < 0 || xx >= PS) continue;
                a += sg[0][oy][ox] * sp[0][yy][xx] + sg[1][oy][ox] * sp[1][yy][xx] + sg[2][oy][ox] * sp[2][yy][xx] + sdm[oy][ox];
            }
        }
        dk[t] = a;
    }
    __syncthreads();
    float* g_rawp = g_raw + (size_t)patch * ld_raw;
    if (mix_mode == 4) {
        // k = k2 / s2,  k2 = wc k1 + (1 - wc) identity
        float a = 0.f;
        for (int i = tid; i < KK; i += blockDim.x) a += dk[i] * k[i];
        float dot = block_sum(a, red);
        float b = 0.f;
        for (int i = tid; i < KK; i += blockDim.x) {
            float dk2 = (dk[i] - dot) / kb.s2;
            b += dk2 * (k1[i] - (i == centre ? 1.f : 0.f));
            dk[i] = kb.wc * dk2;                                  // now d k1
        }
        float dwc = block_sum(b, red);
        if (tid == 0) g_rawp[KK] = dwc;
        __syncthreads();
    }
    {
        float a = 0.f;
        for (int i = tid; i < KK; i += blockDim.x) a += dk[i] * k1[i];
        float dot = block_sum(a, red);
        for (int i = tid; i < KK; i += blockDim.x)
            g_rawp[i] = norm_mode == 0 ? (dk[i] - dot) / kb.s : k1[i] * (dk[i] - dot);
    }
}

// predictor input rows: feat[n] = [mean_c gt(patch n) (PS^2) | mean_c pred(patch n) (PS^2)], pixels in patch-raster order
__global__ void blur_gray_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, float* __restrict__ feat, int PN, int PS) {
    const int patch = blockIdx.x, pi = patch / PN, pj = patch % PN, S = PN * PS, npix = PS * PS;
    for (int t = threadIdx.x; t < npix; t += blockDim.x) {
        int y = t / PS, x = t % PS;
        size_t g = ((size_t)(pi * PS + y) * S + (pj * PS + x)) * 3;
        feat[(size_t)patch * 2 * npix + t] = (gt[g] + gt[g + 1] + gt[g + 2]) / 3.f;
        feat[(size_t)patch * 2 * npix + npix + t] = (pred[g] + pred[g + 1] + pred[g + 2]) / 3.f;
    }
}
__global__ void blur_gray_bwd_kernel(const float* __restrict__ g_feat, float* __restrict__ g_pred, int PN, int PS) {
    const int patch = blockIdx.x, pi = patch / PN, pj = patch % PN, S = PN * PS, npix = PS * PS;
    for (int t = threadIdx.x; t < npix; t += blockDim.x) {
        int y = t / PS, x = t % PS;
        size_t g = ((size_t)(pi * PS + y) * S + (pj * PS + x)) * 3;
        float v = g_feat[(size_t)patch * 2 * npix + npix + t] / 3.f;
        g_pred[g] = v; g_pred[g + 1] = v; g_pred[g + 2] = v;
    }
}

int check_geom(int64_t patch_num, int64_t patch_size, int64_t kernel_size, int norm_mode, int mix_mode, int boundary_mode, int64_t ld_raw) {
    HNR_CHECK_ARG(patch_num > 0 && patch_size > 0 && patch_size <= MAX_PS, "blur_learn: patch_size must be in 1..16");
    HNR_CHECK_ARG(kernel_size > 0 && kernel_size <= MAX_KS && (kernel_size & 1), "blur_learn: kernel_size must be odd and <= 15");
    HNR_CHECK_ARG(norm_mode == 0 || norm_mode == 1, "blur_learn: learnable_blur_kernel_norm must be 0 (/sum) or 1 (softmax)");
    HNR_CHECK_ARG(mix_mode == 0 || mix_mode == 4, "blur_learn: learnable_blur_kernel_mode must be 0 or 4");
    HNR_CHECK_ARG(boundary_mode >= 0 && boundary_mode <= 2, "blur_learn: boundary_mode must be 0, 1 or 2");
    HNR_CHECK_ARG(ld_raw >= kernel_size * kernel_size + (mix_mode == 4 ? 1 : 0), "blur_learn: predictor rows too short");
    return HNR_OK;
}

}  // namespace

extern "C" int hnr_blur_gray_fwd(const float* pred, const float* gt, int64_t patch_num, int64_t patch_size, float* feat, void* stream) {
    HNR_CHECK_ARG(patch_num > 0 && patch_size > 0, "blur_gray: bad shape");
    blur_gray_fwd_kernel<<<(unsigned)(patch_num * patch_num), 64, 0, (cudaStream_t)stream>>>(pred, gt, feat, (int)patch_num, (int)patch_size);
    HNR_CHECK_LAUNCH("blur_gray_fwd");
    return HNR_OK;
}

extern "C" int hnr_blur_gray_bwd(const float* g_feat, int64_t patch_num, int64_t patch_size, float* g_pred, void* stream) {
    HNR_CHECK_ARG(patch_num > 0 && patch_size > 0, "blur_gray: bad shape");
    blur_gray_bwd_kernel<<<(unsigned)(patch_num * patch_num), 64, 0, (cudaStream_t)stream>>>(g_feat, g_pred, (int)patch_num, (int)patch_size);
    HNR_CHECK_LAUNCH("blur_gray_bwd");
    return HNR_OK;
}

extern "C" int hnr_blur_learn_fwd(const float* pred, const float* raw, int64_t ld_raw, int64_t patch_num, int64_t patch_size,
                                  int64_t kernel_size, int norm_mode, int mix_mode, int boundary_mode, float* out, void* stream) {
    int rc = check_geom(patch_num, patch_size, kernel_size, norm_mode, mix_mode, boundary_mode, ld_raw);
    if (rc != HNR_OK) return rc;
    int threads = (int)(((patch_size * patch_size + 31) / 32) * 32);
    if (threads < 96) threads = 96;
    blur_learn_kernel<false><<<(unsigned)(patch_num * patch_num), threads, 0, (cudaStream_t)stream>>>(
        pred, raw, (int)ld_raw, nullptr, out, nullptr, (int)patch_num, (int)patch_size, (int)kernel_size, norm_mode, mix_mode, boundary_mode);
    HNR_CHECK_LAUNCH("blur_learn_fwd");
    return HNR_OK;
}

extern "C" int hnr_blur_learn_bwd(const float* pred, const float* raw, int64_t ld_raw, const float* g_out, int64_t patch_num,
                                  int64_t patch_size, int64_t kernel_size, int norm_mode, int mix_mode, int boundary_mode, float* g_pred,
                                  float* g_raw, void* stream) {
    int rc = check_geom(patch_num, patch_size, kernel_size, norm_mode, mix_mode, boundary_mode, ld_raw);
    if (rc != HNR_OK) return rc;
    int threads = (int)(((patch_size * patch_size + 31) / 32) * 32);
    if (threads < 96) threads = 96;
    blur_learn_kernel<true><<<(unsigned)(patch_num * patch_num), threads, 0, (cudaStream_t)stream>>>(
        pred, raw, (int)ld_raw, g_out, g_pred, g_raw, (int)patch_num, (int)patch_size, (int)kernel_size, norm_mode, mix_mode, boundary_mode);
    HNR_CHECK_LAUNCH("blur_learn_bwd");
    return HNR_OK;
}
