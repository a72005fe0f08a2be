// Frame-dict producer on the device: the per-item arithmetic of ScannetFtDataset.__getitem__
// (data/scannet_ft_dataset.py:736-976) from the decoded images on.  The scene's frames stay resident in HBM as uint8
// (a ScanNet scene is ~1 GB of 180 GB); an item is then two launches instead of 1+V host image conversions and a
// 30 MB host->device copy:
//   frame_rays_kernel   pixel grid of the dilated patches / one patch / the full frame (:887-945), get_dtu_raydir
//                       (data/data_utils.py:57-71) and the ground-truth lookup gt_image_full[py, px] (:957)
//   frame_views_kernel  images_nearest (V,H,W,3) fp32 = uint8 / 255 of the chosen frames (T.ToTensor, :743 / :823-826)
// Both are HBM-bound streaming kernels; the first one is tiny (R <= 4096 rays when training).
#include "common.cuh"
#include "hnr.h"

namespace {

// One thread per ray.  patches == nullptr: full frame inside the margin, row-major (px = margin + col, py = margin + row).
// Otherwise ray (row, col) of the (PN*PS)^2 grid belongs to patch (row / PS) * PN + col / PS with origin (x0, y0) and
// dilation d:  px = x0 + d * (col % PS), py = y0 + d * (row % PS)   (np.meshgrid is 'xy'-indexed: px varies along columns).
__global__ void __launch_bounds__(256) frame_rays_kernel(const int32_t* __restrict__ patches, int PN, int PS, int cols, int64_t n,
                                                         int margin, int W, int H, const float* __restrict__ intrinsic,
                                                         const float* __restrict__ c2w, int dir_norm,
                                                         const unsigned char* __restrict__ frame_u8, float* __restrict__ pixel_idx,
                                                         float* __restrict__ raydir, float* __restrict__ gt_image) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int row = (int)(r / cols), col = (int)(r % cols);
    int ix, iy;
    if (patches == nullptr) {
        ix = margin + col;
        iy = margin + row;
    } else {
        const int32_t* p = patches + 3 * ((row / PS) * PN + col / PS);
        ix = p[0] + p[2] * (col % PS);
        iy = p[1] + p[2] * (row % PS);
    }
    const float px = (float)ix, py = (float)iy;
    pixel_idx[2 * r + 0] = px;
    pixel_idx[2 * r + 1] = py;
    // get_dtu_raydir: same operation order as the numpy fp32 expression; IEEE division (no fast-math in this build)
    const float fx = __ldg(intrinsic + 0), cx = __ldg(intrinsic + 2), fy = __ldg(intrinsic + 4), cy = __ldg(intrinsic + 5);
    const float x = __fdiv_rn(__fsub_rn(__fadd_rn(px, 0.5f), cx), fx);
    const float y = __fdiv_rn(__fsub_rn(__fadd_rn(py, 0.5f), cy), fy);
    float d[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {       // dirs @ rot.T : d_j = x r_j0 + y r_j1 + 1 r_j2   (c2w is 4x4 row-major)
        const float r0 = __ldg(c2w + 4 * j + 0), r1 = __ldg(c2w + 4 * j + 1), r2 = __ldg(c2w + 4 * j + 2);
        d[j] = __fadd_rn(__fadd_rn(__fmul_rn(x, r0), __fmul_rn(y, r1)), r2);
    }
    if (dir_norm) {
        const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
        const float den = __fadd_rn(nrm, 1e-5f);
#pragma unroll
        for (int j = 0; j < 3; ++j) d[j] = __fdiv_rn(d[j], den);
    }
    raydir[3 * r + 0] = d[0];
    raydir[3 * r + 1] = d[1];
    raydir[3 * r + 2] = d[2];
    if (gt_image != nullptr) {
        const unsigned char* s = frame_u8 + 3 * ((int64_t)iy * W + ix);
#pragma unroll
        for (int c = 0; c < 3; ++c) gt_image[3 * r + c] = __fdiv_rn((float)s[c], 255.f);
    }
}

// images_nearest: V frames of the resident bank -> fp32 in [0,1].  4 bytes in, 16 bytes out per thread step.
__global__ void __launch_bounds__(256) frame_views_kernel(const unsigned char* __restrict__ bank, const int32_t* __restrict__ view_ids,
                                                          int64_t frame_bytes, int64_t quads_per_frame, int64_t total_quads,
                                                          float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total_quads; q += stride) {
        const int64_t v = q / quads_per_frame, i = q % quads_per_frame;
        const unsigned char* src = bank + (int64_t)__ldg(view_ids + v) * frame_bytes;
        float* dst = out + v * frame_bytes;
        const int64_t b0 = 4 * i;
        if (b0 + 4 <= frame_bytes && ((frame_bytes & 3) == 0)) {
            const uchar4 u = *reinterpret_cast<const uchar4*>(src + b0);
            const float4 f = make_float4(__fdiv_rn((float)u.x, 255.f), __fdiv_rn((float)u.y, 255.f), __fdiv_rn((float)u.z, 255.f),
                                         __fdiv_rn((float)u.w, 255.f));
            *reinterpret_cast<float4*>(dst + b0) = f;
        } else {
            for (int64_t b = b0; b < b0 + 4 && b < frame_bytes; ++b) dst[b] = __fdiv_rn((float)src[b], 255.f);
        }
    }
}

}  // namespace

extern "C" int hnr_frame_rays(const int32_t* patches, int64_t patch_num, int64_t patch_size, int64_t width, int64_t height,
                              int64_t margin, const float* intrinsic, const float* c2w, int dir_norm, const uint8_t* frame_u8,
                              float* pixel_idx, float* raydir, float* gt_image, void* stream) {
    HNR_CHECK_ARG(intrinsic && c2w && pixel_idx && raydir, "frame_rays: null pointer");
    HNR_CHECK_ARG(width > 0 && height > 0 && margin >= 0, "frame_rays: bad frame size");
    HNR_CHECK_ARG(gt_image == nullptr || frame_u8 != nullptr, "frame_rays: gt_image requested without a frame");
    int64_t cols, rows;
    if (patches == nullptr) {
        cols = width - 2 * margin;
        rows = height - 2 * margin;
    } else {
        HNR_CHECK_ARG(patch_num > 0 && patch_size > 0, "frame_rays: bad patch grid");
        cols = rows = patch_num * patch_size;
    }
    const int64_t n = cols * rows;
    if (n <= 0) return HNR_OK;
    HNR_CHECK_ARG(n <= 0x7fffffffLL * 256, "frame_rays: too many rays");
    frame_rays_kernel<<<(unsigned)hnr_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(patches, (int)patch_num, (int)patch_size, (int)cols, n,
                                                                                   (int)margin, (int)width, (int)height, intrinsic, c2w,
                                                                                   dir_norm, frame_u8, pixel_idx, raydir, gt_image);
    HNR_CHECK_LAUNCH("frame_rays");
    return HNR_OK;
}

extern "C" int hnr_frame_views(const uint8_t* bank, const int32_t* view_ids, int64_t n_views, int64_t frame_bytes, float* out,
                               void* stream) {
    HNR_CHECK_ARG(bank && view_ids && out, "frame_views: null pointer");
    HNR_CHECK_ARG(n_views >= 0 && frame_bytes > 0, "frame_views: bad sizes");
    if (n_views == 0) return HNR_OK;
    const int64_t qpf = hnr_cdiv(frame_bytes, 4), total = qpf * n_views;
    const int64_t blocks = hnr_cdiv(total, 256);
    const int g = (int)(blocks < 16 * HNR_NUM_SMS ? blocks : 16 * HNR_NUM_SMS);
    frame_views_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(bank, view_ids, frame_bytes, qpf, total, out);
    HNR_CHECK_LAUNCH("frame_views");
    return HNR_OK;
}
