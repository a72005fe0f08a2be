"""Blur-handling module -- host-side mirror of BaseRenderingModel.blur_update_output
(models/base_rendering_model.py:677-786) and learnable_blur_update_output (:827-1020), plus the pre-defined
kernel bank the dataset builds (data/scannet_ft_dataset.py:184-242)."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops


def blur_select(coarse_raycolor, gt_image, blur_kernels, patch_num: int, patch_size: int):
    """coarse_raycolor, gt_image (1,S*S,3) on the S x S patch raster; blur_kernels (1,Nk,k,k).
    Returns (new coarse_raycolor (1,S*S,3), select_index (patch_num^2,) int32)."""
    out, sel = ops.BlurSelectFn.apply(coarse_raycolor.reshape(-1, 3), gt_image.reshape(-1, 3), blur_kernels.reshape((-1,) + tuple(blur_kernels.shape[-2:])),
                                      int(patch_num), int(patch_size))
    return out.view(1, -1, 3), sel


def blur_update_output(model, faster_version=True):
    """drop-in body for the method: reads model.output["coarse_raycolor"], model.gt_image,
    model.blur_kernels, model.dilation_PatchNum/Size and replaces model.output["coarse_raycolor"]."""
    if not faster_version or int(model.dilation_PatchNum) <= 0:
        raise NotImplementedError
    out, _ = blur_select(model.output["coarse_raycolor"], model.gt_image, model.blur_kernels.to(model.output["coarse_raycolor"].device),
                         int(model.dilation_PatchNum), int(model.dilation_PatchSize))
    model.output["coarse_raycolor"] = out


def blur_predictor_forward(blur_predictor, feat: torch.Tensor) -> torch.Tensor:
    """learn_blur_kernel_block (point_aggregators.py:715-749: Linear 2*ps^2 ->128->128->128-> ks^2[+1], LeakyReLU, final Sigmoid)
    through the dense-layer kernels; feat (N, 2*ps^2)."""
    if isinstance(blur_predictor, (list, tuple)):
        raise NotImplementedError("learnable_blur_kernel_conv=1 (conv stem before the predictor MLP) is not selected by any shipped "
                                  "script and is not implemented; there is no fallback path")
    lins = [m for m in blur_predictor if isinstance(m, torch.nn.Linear)]
    x = feat
    for i, l in enumerate(lins):
        x = ops.linear([x], l.weight, l.bias, ops.ACT_SIGMOID if i == len(lins) - 1 else ops.ACT_LRELU)
    return x


def learnable_blur(coarse_raycolor, gt_image, blur_predictor, patch_num: int, patch_size: int, kernel_size: int = 9, kernel_mode: int = 4,
                   kernel_norm: int = 0, boundary_mode: int = 0):
    """coarse_raycolor, gt_image (1,S*S,3) on the S x S patch raster -> (new coarse_raycolor (1,S*S,3), raw predictor output
    (N, ks^2 [+1]))."""
    if kernel_mode not in (0, 4) or kernel_norm not in (0, 1) or boundary_mode not in (0, 1, 2):
        raise NotImplementedError(f"learnable blur: mode {kernel_mode} / norm {kernel_norm} / boundary {boundary_mode}")
    pred = coarse_raycolor.reshape(-1, 3)
    feat = ops.BlurGrayFn.apply(pred, gt_image.reshape(-1, 3), int(patch_num), int(patch_size))
    raw = blur_predictor_forward(blur_predictor, feat)
    out = ops.BlurLearnFn.apply(pred, raw, int(patch_num), int(patch_size), int(kernel_size), int(kernel_norm), int(kernel_mode),
                                int(boundary_mode))
    return out.view(1, -1, 3), raw


def learnable_blur_update_output(model, blur_predictor, visualize=False, faster_version=True):
    """drop-in body for BaseRenderingModel.learnable_blur_update_output (models/base_rendering_model.py:827-1020): reads
    model.output["coarse_raycolor"], model.gt_image, model.dilation_PatchNum/Size and model.opt.{learnable_blur_kernel_size,
    learnable_blur_kernel_mode, learnable_blur_kernel_norm, learnable_blur_kernel_conv, boundary_mode}; replaces
    model.output["coarse_raycolor"]."""
    if not faster_version or int(model.dilation_PatchNum) <= 0:
        raise NotImplementedError
    opt = model.opt
    if getattr(opt, "learnable_blur_kernel_conv", 0):
        raise NotImplementedError("learnable_blur_kernel_conv=1 is not implemented (no shipped script selects it)")
    out, _ = learnable_blur(model.output["coarse_raycolor"], model.gt_image.to(model.output["coarse_raycolor"].device), blur_predictor,
                            int(model.dilation_PatchNum), int(model.dilation_PatchSize), int(getattr(opt, "learnable_blur_kernel_size", 9)),
                            int(getattr(opt, "learnable_blur_kernel_mode", 4)), int(getattr(opt, "learnable_blur_kernel_norm", 0)),
                            int(getattr(opt, "boundary_mode", 0)))
    model.output["coarse_raycolor"] = out


def _rotate_bilinear(img: np.ndarray, angle_deg: float) -> np.ndarray:
    """rotation about (w//2, h//2) with bilinear sampling and zero border -- the arithmetic of
    cv2.warpAffine(getRotationMatrix2D(...), INTER_LINEAR) that imutils.rotate wraps, up to OpenCV's
    fixed-point interpolation weights (1/32 px coordinate quantisation)."""
    h, w = img.shape
    cx, cy = w // 2, h // 2
    a = math.radians(angle_deg)
    al, be = math.cos(a), math.sin(a)
    M = np.array([[al, be, (1 - al) * cx - be * cy], [-be, al, be * cx + (1 - al) * cy]])
    A = np.vstack([M, [0, 0, 1]])
    Ai = np.linalg.inv(A)
    out = np.zeros_like(img, dtype=np.float64)
    for y in range(h):
        for x in range(w):
            sx = Ai[0, 0] * x + Ai[0, 1] * y + Ai[0, 2]
            sy = Ai[1, 0] * x + Ai[1, 1] * y + Ai[1, 2]
            sx, sy = round(sx * 32) / 32.0, round(sy * 32) / 32.0     # INTER_TAB_SIZE quantisation
            x0, y0 = math.floor(sx), math.floor(sy)
            fx, fy = sx - x0, sy - y0
            v = 0.0
            for dy, wy in ((0, 1 - fy), (1, fy)):
                for dx, wx in ((0, 1 - fx), (1, fx)):
                    xx, yy = x0 + dx, y0 + dy
                    if 0 <= xx < w and 0 <= yy < h:
                        v += wx * wy * img[yy, xx]
            out[y, x] = v
    return out


def _rotate(img: np.ndarray, angle_deg: float) -> np.ndarray:
    """imutils.rotate == cv2.warpAffine(getRotationMatrix2D((w//2,h//2), angle, 1)).  cv2 is used when
    importable (bit-identical to the reference's dataset code); otherwise the restatement above,
    which differs from OpenCV's fixed-point interpolation by < 1e-2 per tap."""
    try:
        import cv2
        h, w = img.shape
        M = cv2.getRotationMatrix2D((w // 2, h // 2), angle_deg, 1.0)
        return cv2.warpAffine(img, M, (w, h))
    except ImportError:
        return _rotate_bilinear(img, angle_deg)


def predefined_blur_kernels(version: int = 3, k_size: int = 9, num_dirs: int = 8, move_dists=(1, 2, 4)) -> np.ndarray:
    """kernel bank of the blur module: line kernels of length dist+1 (asymmetric, all num_dirs
    directions) and 2*dist+1 (symmetric, half the directions), each normalised to sum 1."""
    c = k_size // 2
    dirs = list(np.linspace(0, 360, num_dirs + 1)[:num_dirs])
    out = []
    if version in (1, 3):
        for d in move_dists:
            k = np.zeros((k_size, k_size))
            k[c - d:c + 1, c] = 255
            for ang in dirs:
                r = _rotate(k, ang)
                out.append(r / r.sum())
    if version in (2, 3):
        for d in move_dists:
            k = np.zeros((k_size, k_size))
            k[c - d:c + d + 1, c] = 255
            for ang in dirs[:num_dirs // 2]:
                r = _rotate(k, ang)
                out.append(r / r.sum())
    return np.stack(out).astype(np.float32)
