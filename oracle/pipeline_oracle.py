"""End-to-end CPU oracle: query oracle -> gathers -> aggregation oracle -> compositing oracle.

TEST INFRASTRUCTURE ONLY (see render_oracle.py / query_oracle.py headers).  Restates
NeuralPointsRayMarching.forward of the reference (models/neural_points_volumetric_model.py:257-391)
for one frame dict; used by tests/, smoke() and bench.py's CPU arm.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import query_oracle as qo
from . import render_oracle as ro

T = torch.from_numpy


def render_from_query(P, cfg: ro.AggCfg, pts: Dict[str, np.ndarray], q: Dict[str, np.ndarray], frame: Dict[str, np.ndarray],
                      vsize_z: float, dtype=torch.float32, params_grad: bool = False, unit_mode: bool = True, xy_override=None,
                      kink_eps: float = 0.0):
    """pts: xyz (N,3), emb (N,32), conf (N,1), color, dir (N,3).  q: output of query_oracle.query (or
    the product's query tensors as numpy).  Returns dict with ray_color (1,R'',3) etc. and the leaf
    tensors (for gradients).
    xy_override (V,R'',SR,2): pixel projections to use instead of the oracle's own (tests pass the PRODUCT's projections after
    checking them against `xy_own` to sub-pixel accuracy: the nearest-pixel lookup is discontinuous, so both sides must truncate the
    same floats).  kink_eps > 0: also return `kink_free` (R'',) -- rays without a hidden unit on a LeakyReLU kink."""
    c = lambda a: T(np.ascontiguousarray(a)).to(dtype)
    pidx = T(np.ascontiguousarray(q["sample_pidx"])).long()
    mask = pidx >= 0
    idx = pidx.clamp(min=0)
    leaf = {k: c(pts[k]).requires_grad_(params_grad) for k in ("emb", "conf", "color", "dir")}
    Pd = {k: v.to(dtype).clone().requires_grad_(params_grad) for k, v in P.items()}
    xyz = c(pts["xyz"])
    campos, camrot = c(frame["campos"])[0], c(frame["camrotc2w"])[0]
    shift = xyz - campos
    xc = shift @ camrot                               # cam_j = sum_i shift_i R[i,j]
    xyz_pers = torch.stack([xc[:, 0] / xc[:, 2], xc[:, 1] / xc[:, 2], xc[:, 2]], -1)
    loc_w = c(q["sample_loc_w"])
    V = cfg.use_nearest
    xy = dv = img = None
    xy_own = None
    if V > 0:
        xy = xy_own = ro.project_to_views(loc_w[0], c(frame["intrinsic_nearest"])[0], c(frame["c2w_nearest"])[0, :V])
        if xy_override is not None:
            xy = (xy_override if torch.is_tensor(xy_override) else T(np.ascontiguousarray(xy_override))).to(dtype).reshape(xy_own.shape)
        dv = ro.delta_viewdirs(loc_w[0], campos, c(frame["campos_nearest"])[0, :V])
        img = c(frame["images_nearest"])[:, :V]
    if kink_eps > 0:
        ro.KINK_TAP = []
    decoded, valid, w, cc = ro.aggregate(Pd, cfg, leaf["color"][idx], torch.eye(3, dtype=dtype), leaf["dir"][idx], leaf["conf"][idx],
                                         leaf["emb"][idx], xyz_pers[idx], xyz[idx], mask, c(q["sample_loc"]), loc_w,
                                         c(q["sample_ray_dirs"]), img_n=img, sample_loc_i_n=xy, delta_viewdir_n=dv)
    kink_free = None
    if kink_eps > 0:
        taps, ro.KINK_TAP = ro.KINK_TAP, None
        kink_free = ro.kink_free_rays(taps, mask, valid, kink_eps)
    dist = ro.ray_dist_from_depth(c(q["sample_loc"])[..., 2], valid, vsize_z, unit_mode)
    bg = c(frame["bg_color"]) if "bg_color" in frame else None
    color, _, opacity, accT, bw, bgT, _ = ro.ray_march(dist, valid, decoded, bg)
    return dict(ray_color=color, opacity=opacity, bg_T=bgT, decoded=decoded, ray_valid=valid, weight=w, conf_coefficient=cc,
                leaf=leaf, params=Pd, xy_own=xy_own, delta_view=dv, kink_free=kink_free)


def query(pts, frame, opt, ts, skip_cell=None):
    return qo.query(pts["xyz"], frame["campos"], frame["camrotc2w"], frame["raydir"], ts, vsize=opt.vsize, vscale=opt.vscale,
                    kernel_size=opt.kernel_size, query_size=opt.query_size, ranges=opt.ranges, radius_limit_scale=opt.radius_limit_scale,
                    SR=opt.SR, K=opt.K, P=opt.P, max_o=getattr(opt, "max_o", None), skip_cell=skip_cell)


def render(P, cfg, pts, frame, opt, ts, dtype=torch.float32, params_grad=False, skip_cell=None, xy_override=None, kink_eps=0.0, q=None):
    q = q if q is not None else qo.query(pts["xyz"], frame["campos"], frame["camrotc2w"], frame["raydir"], ts, vsize=opt.vsize, vscale=opt.vscale,
                 kernel_size=opt.kernel_size, query_size=opt.query_size, ranges=opt.ranges, radius_limit_scale=opt.radius_limit_scale,
                 SR=opt.SR, K=opt.K, P=opt.P, max_o=getattr(opt, "max_o", None), skip_cell=skip_cell)
    out = render_from_query(P, cfg, pts, q, frame, float(opt.vsize[2]), dtype, params_grad, unit_mode=opt.raydist_mode_unit > 0,
                            xy_override=xy_override, kink_eps=kink_eps)
    out["query"] = q
    return out


def training_loss(out, gt_masked: torch.Tensor):
    """MSE on ray-masked colours + 1e-4 * zero-one regulariser on conf_coefficient
    (models/base_rendering_model.py:1114-1118, :1229-1240)."""
    mse = torch.nn.functional.mse_loss(out["ray_color"], gt_masked.to(out["ray_color"].dtype))
    v = out["conf_coefficient"].clamp(1e-3, 1 - 1e-3)
    return mse + 1e-4 * torch.mean(torch.log(v) + torch.log(1 - v))
