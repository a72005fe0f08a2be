"""The hot-path conductor -- host-side mirror of NeuralPointsRayMarching
(models/neural_points_volumetric_model.py:225-427) and of ``fill_invalid`` (:87-126).

``forward(**frame_dict)`` takes the reference's frame dict and returns the
reference's output dict (Appendix A.3).  Internally it is the fused B200 path: voxel query ->
aggregation kernels reading the point tables directly -> compositing with the ray_dist prologue
fused in.  One device->host readback per forward (the query's output counts).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import diff_ray_marching as drm
from .neural_points import NeuralPoints
from .point_aggregators import PointAggregator


class NeuralPointsRayMarching(nn.Module):
    def __init__(self, tonemap_func=None, render_func=None, blend_func=None, aggregator=None, is_compute_depth=False,
                 neural_points=None, opt=None, num_pos_freqs=0, num_viewdir_freqs=0, **kwargs):
        super().__init__()
        self.aggregator = aggregator
        self.num_pos_freqs = num_pos_freqs
        self.num_viewdir_freqs = num_viewdir_freqs
        self.render_func = render_func if render_func is not None else drm.radiance_render
        self.blend_func = blend_func if blend_func is not None else drm.alpha_blend
        drm._check_funcs(self.render_func, self.blend_func)
        self.tone_map = tonemap_func if tonemap_func is not None else drm.no_tone_map
        self.return_depth = is_compute_depth
        if self.return_depth:
            raise NotImplementedError("is_compute_depth is off in every shipped config")
        self.return_color = True
        self.opt = opt
        self.neural_points = neural_points
        self.near_far: Optional[tuple] = None     # set to (near, far) floats to skip the per-call D2H read
        # one-shot callback run after the query and before the first kernel that reads the point tables: a data-parallel training
        # loop parks its deferred point-table optimiser step here (parallel.train_step), so that the gradient all-reduce of the
        # previous step overlaps with this forward's packing, pyramid and query
        self.before_point_read = None

    def forward(self, campos, raydir, gt_image=None, bg_color=None, camrotc2w=None, pixel_idx=None, near=None, far=None, focal=None,
                h=None, w=None, intrinsic=None, aux_image=None, c2w=None, c2w_nearest=None, images_nearest=None, campos_nearest=None,
                intrinsic_nearest=None, vid_angle_nearest=None, frame_weight_nearest=None, **kargs):
        opt = self.opt
        output: Dict[str, torch.Tensor] = {}
        nf = self.near_far
        inputs = {"pixel_idx": pixel_idx, "camrotc2w": camrotc2w, "campos": campos, "near": near, "far": far, "h": h, "w": w,
                  "intrinsic": intrinsic, "raydir": raydir}
        if getattr(opt, "dynamic_nearest", 0):
            opt.use_nearest = c2w_nearest.shape[1]
        V = int(opt.use_nearest)
        # everything that does not depend on the query is issued BEFORE it: the query ends with the one device->host read-back of a
        # forward, and whatever the host still has to launch after that read-back runs with the GPU idle
        self.aggregator.prepack()
        views = self.aggregator.prepare_views(images_nearest, c2w_nearest[0, :V]) if V > 0 else None
        sample_pidx, sample_loc, sample_loc_w, sample_ray_dirs, ray_mask_tensor, vsize, extras = self.neural_points.query(
            inputs, near=None if nf is None else nf[0], far=None if nf is None else nf[1])
        if self.before_point_read is not None:
            cb, self.before_point_read = self.before_point_read, None
            cb()
        decoded, ray_valid, weight, conf_coefficient = self.aggregator.forward_fused(
            self.neural_points, sample_pidx, sample_loc, sample_loc_w, sample_ray_dirs, campos, camrotc2w, extras=extras,
            img_n=images_nearest, c2w_n=None if V == 0 else c2w_nearest[0, :V], intrinsic_n=None if V == 0 else intrinsic_nearest[0],
            campos_n=None if V == 0 else campos_nearest[0, :V], views=views)
        output["blur_predictor"] = self.aggregator.learn_blur_kernel_block if getattr(opt, "is_train", False) else None
        keep_w = not ((opt.sparse_loss_weight <= 0) and ("conf_coefficient" not in opt.zero_one_loss_items) and opt.prob == 0)
        if not keep_w:
            weight, conf_coefficient = None, None
        output["queried_shading"] = torch.logical_not(torch.any(ray_valid, dim=-1, keepdims=True)).repeat(1, 1, 3).to(torch.float32)
        if "bg_ray" in kargs:
            bg_color = None
        ray_color, opacity, acc_transmission, blend_weight, background_transmission, _ = drm.ray_march_from_depth(
            sample_loc, ray_valid, decoded, float(vsize[2]), int(opt.raydist_mode_unit > 0), bg_color)
        ray_color = self.tone_map(ray_color)
        output["coarse_raycolor_patch"] = ray_color
        output["coarse_raycolor"] = ray_color
        output["coarse_point_opacity"] = opacity
        output["coarse_is_background"] = background_transmission
        output["ray_mask"] = ray_mask_tensor
        output["ray_ids"] = None if extras is None else extras.ray_ids       # (R'',) ids of the kept rays (extra key, not in the reference's dict)
        if weight is not None:
            output["weight"] = weight.detach()
            output["blend_weight"] = blend_weight.detach()
            output["conf_coefficient"] = conf_coefficient
        if opt.prob == 1 and output["coarse_point_opacity"].shape[1] > 0:
            self._probe_outputs(output, weight, conf_coefficient, sample_pidx, sample_loc_w)
        self.last_extras = extras
        return output

    def prefetch_query(self, campos=None, raydir=None, camrotc2w=None, **kargs) -> None:
        """Software pipelining for training loops: enqueue the voxel query of the NEXT frame (pass its frame dict) before the
        current frame's backward pass is issued.  Needs `near_far` to be set (no device read of near / far).  The next forward()
        with the same tensors consumes the result without a host wait; the reference has no counterpart (its query blocks)."""
        if self.near_far is None:
            raise RuntimeError("prefetch_query needs net.near_far = (near, far) as python floats")
        self.neural_points.prefetch({"raydir": raydir, "campos": campos, "camrotc2w": camrotc2w}, self.near_far[0], self.near_far[1])

    def _probe_outputs(self, output, weight, conf_coefficient, sample_pidx, sample_loc_w):
        """the eight extra tensors point growing reads when opt.prob == 1 (:394-425); plain tensor ops
        on our kernels' outputs -- not a hot path (runs every prob_freq=10000 iterations)."""
        np_ = self.neural_points
        idx = torch.clamp(sample_pidx, min=0).long()
        mo, oi = torch.max(output["coarse_point_opacity"], dim=-1, keepdim=True)
        output["ray_max_shading_opacity"] = mo
        oi = oi[..., None]
        loc = torch.gather(sample_loc_w, 2, oi.expand(-1, -1, -1, 3)).squeeze(2)
        output["ray_max_sample_loc_w"] = loc
        w = torch.gather(weight * conf_coefficient, 2, oi.expand(-1, -1, -1, weight.shape[-1])).squeeze(2)[..., None]
        pid = torch.gather(idx, 2, oi.expand(-1, -1, -1, idx.shape[-1])).squeeze(2)                   # (1,R,K)
        xyz = np_.xyz[pid]
        output["ray_max_far_dist"] = torch.min(torch.norm(xyz - loc[..., None, :], dim=-1), dim=-1, keepdim=True)[0]
        take = lambda t: None if t is None else t[0][pid]
        col, dr, cf, em = take(np_.points_color), take(np_.points_dir), take(np_.points_conf), take(np_.points_embeding)
        output["shading_avg_color"] = torch.sum(col * w, dim=-2) if col is not None else None
        output["shading_avg_dir"] = torch.sum(dr * w, dim=-2) if dr is not None else None
        output["shading_avg_conf"] = torch.sum(cf * w, dim=-2) if cf is not None else None
        output["shading_avg_embedding"] = torch.sum(em * w, dim=-2)


def fill_invalid(output: Dict[str, torch.Tensor], bg_color: Optional[torch.Tensor], ray_ids: Optional[torch.Tensor] = None,
                 bg_ray: Optional[torch.Tensor] = None, tonemap_func=None):
    """scatter the R'' valid-ray results back to all R rays (reference models/neural_points_volumetric_model.py:87-126): misses get the
    (tone-mapped) background colour, transmittance 1, opacity 0; also `coarse_mask` = 1 - coarse_is_background, the patch colours,
    the `bg_ray` variant (per-ray background added under the kept rays' colours) and -- with opt.prob == 1 -- the unmasked probe
    tensors point growing reads (`unmask`, :128-140: rays that hit nothing get zeros).  With `ray_ids` (from the query) no host
    sync is needed; otherwise falls back to nonzero(ray_mask)."""
    ray_mask = output["ray_mask"]
    B, R = ray_mask.shape
    if ray_ids is None:
        ray_ids = torch.nonzero(ray_mask[0] > 0, as_tuple=False).view(-1)
    idx = ray_ids.long()
    dev = ray_mask.device
    tm = tonemap_func if tonemap_func is not None else (lambda t: t)
    out = dict(output)
    c = output["coarse_raycolor"]
    isbg = torch.ones((B, R, 1), device=dev, dtype=output["coarse_is_background"].dtype).index_copy(1, idx, output["coarse_is_background"])
    out["coarse_is_background"] = isbg
    out["coarse_mask"] = 1 - isbg
    if bg_ray is not None:
        out["coarse_raycolor"] = (isbg * bg_ray.to(c)).index_add(1, idx, c)
    else:
        full = tm((bg_color.reshape(1, 1, 3).to(c).expand(B, R, 3) if bg_color is not None else torch.zeros((B, R, 3), device=dev)).clone())
        out["coarse_raycolor"] = full.index_copy(1, idx, c)
        cp = output.get("coarse_raycolor_patch")
        if cp is not None:
            out["coarse_raycolor_patch"] = full.repeat(1, 1, cp.shape[-1] // 3).index_copy(1, idx, cp)
    op = output["coarse_point_opacity"]
    out["coarse_point_opacity"] = torch.zeros((B, R, op.shape[-1]), device=dev, dtype=op.dtype).index_copy(1, idx, op)
    out["queried_shading"] = torch.ones((B, R, 3), device=dev).index_copy(1, idx, output["queried_shading"])
    for k in ("weight", "blend_weight", "conf_coefficient"):
        if k in output and output[k] is not None:
            t = output[k]
            out[k] = torch.zeros((B, R) + tuple(t.shape[2:]), device=dev, dtype=t.dtype).index_copy(1, idx, t)
    for k in ("ray_max_sample_loc_w", "ray_max_shading_opacity", "shading_avg_color", "shading_avg_dir", "shading_avg_conf",
              "shading_avg_embedding", "ray_max_far_dist"):
        if k in output and output[k] is not None:
            t = output[k]
            out[k] = torch.zeros((B, R) + tuple(t.shape[2:]), device=dev, dtype=t.dtype).index_copy(1, idx, t)
    return out
