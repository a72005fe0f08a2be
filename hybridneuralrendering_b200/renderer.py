"""Full-frame renderer and training step over the fused hot path.

The reference renders a frame in chunks of 2,304 rays and rebuilds the occupancy grid for every
chunk (run/train_ft.py:282-351 + query_point_indices_worldcoords.py:616); here the grid is built
once per point set, chunks are as large as memory allows, and every chunk's colours are written
straight into the (H*W,3) device image through the query's ray-id list (no nonzero(), no per-chunk
device->host copy)."""
from __future__ import annotations

from typing import Dict, Optional

import torch


def render_rays(net, frame: Dict[str, torch.Tensor], chunk_rays: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """frame: the reference's frame dict on the device (raydir (1,R,3), campos, camrotc2w, near, far,
    nearest-view tensors...).  Returns colours (R,3) with bg colour for rays that hit nothing.
    chunk_rays=None: as few ray chunks as the query's int32 index space allows (normally ONE per frame, i.e.
    one host readback per frame; the aggregator bounds activation memory by itself)."""
    raydir = frame["raydir"]
    R = raydir.shape[1]
    dev = raydir.device
    bg = frame.get("bg_color")
    if out is None:
        out = torch.empty((R, 3), device=dev, dtype=torch.float32)
    out[:] = bg.reshape(1, 3).to(out) if bg is not None else 0.0
    if chunk_rays is None:
        opt = net.opt
        cap = (2 ** 31 - 1) // (int(opt.SR) * int(opt.K))
        n = -(-R // cap)
        chunk_rays = -(-R // n)
    static = {k: v for k, v in frame.items() if k not in ("raydir", "pixel_idx", "gt_image")}
    with torch.no_grad():
        for r0 in range(0, R, chunk_rays):
            r1 = min(R, r0 + chunk_rays)
            o = net(raydir=raydir[:, r0:r1], pixel_idx=None, **static)
            ids = net.last_extras.ray_ids.long()
            if ids.numel():
                out.index_copy_(0, ids + r0, o["coarse_raycolor"][0])
    # range guard of the split-fp16 kernels: the queries above only checked the frames BEFORE them; check this frame's last
    # chunk too (one event wait -- the caller reads the image next anyway), so that a saturated image is never returned silently
    from . import ops
    ops.status_fetch_async(dev)
    ev = torch.cuda.Event()
    ev.record()
    ev.synchronize()
    ops.status_check(dev)
    return out


def training_loss(output: Dict[str, torch.Tensor], gt_image: torch.Tensor, zero_one_weight: float = 1e-4, frame_weight: float = 1.0):
    """MSE on ray-masked colours (+1e-6) scaled by frame_weight, plus the zero-one regulariser on
    conf_coefficient (reference models/base_rendering_model.py:1114-1118, :1198-1240).
    `output` is the un-filled output of NeuralPointsRayMarching (R'' kept rays)."""
    if output.get("ray_ids") is not None and output["coarse_raycolor"].is_cuda:
        # fused: one launch computes the value and both gradients (csrc/loss.cu); gt is looked up through the query's kept-ray list
        from . import ops
        cc = output.get("conf_coefficient") if zero_one_weight > 0 else None
        return ops.TrainLossFn.apply(output["coarse_raycolor"], cc, gt_image, output["ray_ids"], float(frame_weight), float(zero_one_weight))
    if output.get("ray_ids") is not None:
        gt = gt_image.index_select(1, output["ray_ids"].long())       # kept-ray list from the query: no nonzero(), no host sync
    else:
        gt = gt_image[:, output["ray_mask"][0] > 0]
    loss = (torch.nn.functional.mse_loss(output["coarse_raycolor"], gt) + 1e-6) * frame_weight
    if "conf_coefficient" in output and zero_one_weight > 0:
        v = output["conf_coefficient"].clamp(1e-3, 1 - 1e-3)
        loss = loss + zero_one_weight * torch.mean(torch.log(v) + torch.log(1 - v))
    return loss
