"""Benchmark bodies shared by bench.py and scripts/: the training step of BASELINE.json configs[2]
(ScanNet scene0241_01-shaped hybrid training step: 640x480 frames, 4096-ray batch = 8x8 dilated patches of 8x8,
V=8 reference-view feature maps, 2M neural points, SR=24, K=8) on synthetic data."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import ops

TRAIN_WORKLOAD = "ScanNet scene0241_01-shaped hybrid training step: 640x480 frames, 4096-ray batch, 8 reference-view feature maps, 2M points"


def build_train_case(dev, points: int = 2_000_000, views: int = 8, seed: int = 0):
    from . import NeuralPoints, NeuralPointsRayMarching, PointAggregator, make_opt
    from . import synthetic as syn
    opt = make_opt("scannet", use_nearest=views, SR=24, is_train=True, drop_ratio=0.5, dilation_setup="8_8_1_8",
                   max_o=1_000_000)       # >= occupied voxels of the 2M-point room (generator must respect max_o)
    xyz = syn.room_scene(points, 0)                        # the point cloud is replicated: same on every rank
    att = syn.point_attributes(np.random.default_rng(0), len(xyz))
    fr = syn.room_frame(H=480, W=640, V=views, patch_num=8, patch_size=8, seed=seed)
    c = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pts = NeuralPoints(32, len(xyz), opt, dev)
    pts.set_points(c(xyz), c(att["emb"])[None], points_color=c(att["color"])[None], points_dir=c(att["dir"])[None],
                   points_conf=c(att["conf"])[None], parameter=True)
    torch.manual_seed(0)
    agg = PointAggregator(opt).to(dev)
    net = NeuralPointsRayMarching(aggregator=agg, neural_points=pts, opt=opt).to(dev)
    net.near_far = (0.1, 8.0)
    frame = {k: (c(v) if isinstance(v, np.ndarray) and v.dtype.kind == "f" else v) for k, v in fr.items()}
    return net, frame


def train_step_benchmark(dev, steps: int = 5, warmup: int = 3, world: int = 1, points: int = 2_000_000, views: int = 8,
                         rank: int = 0, stage_split: bool = True, prefetch: bool = True) -> Dict:
    """fwd + loss + bwd (+ gradient all-reduce over NCCL when world > 1) timed with CUDA events; Adam timed separately.
    Returns a dict; 'value' is THIS rank's rays/s -- the caller aggregates over ranks with the max-over-ranks time."""
    import torch.distributed as dist
    from . import parallel
    from .renderer import training_loss
    # weak scaling: every rank gets the SAME amount of work (the same 4096-ray batch; gradients are still all-reduced), so that the
    # max-over-ranks time measures the collective and not the spread of valid-sample counts between different batches
    net, frame = build_train_case(dev, points, views, seed=0)
    agg = net.aggregator
    R = frame["raydir"].shape[1]
    params = [p for p in net.parameters() if p.requires_grad]
    opt_net = torch.optim.Adam([p for n, p in net.named_parameters() if p.requires_grad and not n.startswith("neural_points.")], lr=5e-4)
    from .optim import FusedAdam
    opt_pts = FusedAdam([p for n, p in net.named_parameters() if p.requires_grad and n.startswith("neural_points.")], lr=2e-3)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    ar = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))

    def fwd_bwd():
        for p in params:
            p.grad = None
        out = net(**frame)
        loss = training_loss(out, frame["gt_image"])
        if prefetch:
            net.prefetch_query(**frame)       # the NEXT step's voxel query (one query per step, software-pipelined one step ahead)
        with ops.tag("backward"):
            loss.backward()
        if world > 1:
            ar[0].record()
            parallel.allreduce_gradients(params, (out["ray_mask"] > 0).sum())
            ar[1].record()
        return out, loss

    for _ in range(warmup):
        fwd_bwd()
        opt_net.step(); opt_pts.step()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(steps)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ops.LAUNCHES = 0
    for s, m, e in ev:
        flush.zero_()
        s.record()
        out, loss = fwd_bwd()
        m.record()
        opt_net.step(); opt_pts.step()
        e.record()
    torch.cuda.synchronize()
    launches = ops.LAUNCHES
    t_fb = sum(s.elapsed_time(m) for s, m, e in ev) / steps
    t_opt = sum(m.elapsed_time(e) for s, m, e in ev) / steps
    stages = {}
    if stage_split:
        ops.TIMERS = []
        fwd_bwd()
        torch.cuda.synchronize()
        for tag, s, e in ops.TIMERS:
            tag = tag.split("[")[0]
            stages[tag] = stages.get(tag, 0.0) + s.elapsed_time(e)
        ops.TIMERS = None
    ex = net.last_extras
    if world > 1:
        torch.cuda.synchronize()
        stages["allreduce (NCCL + scaling, device time of the last step)"] = ar[0].elapsed_time(ar[1])
    return {"metric": "train rays/s (fwd+bwd)", "value": R / (t_fb * 1e-3), "unit": "rays/s", "ms_fwd_bwd": t_fb, "ms_adam": t_opt,
            "rays": R, "kept_rays": int(ex.n_rays), "valid_samples": int(ex.n_valid), "valid_neighbours": agg.last_valid_neighbours(),
            "points": points, "views": views, "loss": float(loss.detach()), "launches_per_step": launches // steps,
            "stage_ms": {k: round(v, 3) for k, v in sorted(stages.items())}, "config": TRAIN_WORKLOAD}


BLUR_WORKLOAD = "full model with blur handling: 32x32 patch rays (4x4 dilated patches of 8x8) + pre-defined degradation-kernel convolution, fwd+bwd"


def blur_train_step_benchmark(dev, steps: int = 5, warmup: int = 3, points: int = 2_000_000, views: int = 8, learnable: bool = False) -> Dict:
    """BASELINE.json configs[3]: as the training step, but 1,024 rays on a 4x4 grid of 8x8 patches, the output filled back to
    the patch raster and passed through the blur module (36 pre-defined 9x9 kernels + identity, per-patch best match,
    models/base_rendering_model.py:677-786) before the loss; forward + backward through blur and render."""
    from . import NeuralPoints, NeuralPointsRayMarching, PointAggregator, make_opt
    from . import synthetic as syn
    from .blur import blur_select, learnable_blur, predefined_blur_kernels
    from .neural_points_volumetric_model import fill_invalid
    opt = make_opt("scannet", use_nearest=views, SR=24, is_train=True, drop_ratio=0.5, dilation_setup="4_8_1_8", max_o=1_000_000,
                   learnable_blur_kernel=int(learnable), boundary_mode=1)
    xyz = syn.room_scene(points, 0)
    att = syn.point_attributes(np.random.default_rng(0), len(xyz))
    fr = syn.room_frame(H=480, W=640, V=views, patch_num=4, patch_size=8, seed=0)
    c = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pts = NeuralPoints(32, len(xyz), opt, dev)
    pts.set_points(c(xyz), c(att["emb"])[None], points_color=c(att["color"])[None], points_dir=c(att["dir"])[None],
                   points_conf=c(att["conf"])[None], parameter=True)
    torch.manual_seed(0)
    net = NeuralPointsRayMarching(aggregator=PointAggregator(opt).to(dev), neural_points=pts, opt=opt).to(dev)
    net.near_far = (0.1, 8.0)
    frame = {k: (c(v) if isinstance(v, np.ndarray) and v.dtype.kind == "f" else v) for k, v in fr.items()}
    kernels = c(predefined_blur_kernels(3))[None]
    R = fr["raydir"].shape[1]
    params = [p for p in net.parameters() if p.requires_grad]

    def fwd_bwd():
        for p in params:
            p.grad = None
        out = net(**frame)
        full = fill_invalid(out, frame["bg_color"], net.last_extras.ray_ids)
        if learnable:
            blurred, _ = learnable_blur(full["coarse_raycolor"], frame["gt_image"], out["blur_predictor"], 4, 8, 9, 4, 0, 1)
        else:
            blurred, _ = blur_select(full["coarse_raycolor"], frame["gt_image"], kernels, 4, 8)
        loss = torch.nn.functional.mse_loss(blurred, frame["gt_image"])
        loss.backward()
        return loss

    for _ in range(warmup):
        fwd_bwd()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    for s, e in ev:
        s.record()
        loss = fwd_bwd()
        e.record()
    torch.cuda.synchronize()
    ms = sum(s.elapsed_time(e) for s, e in ev) / steps
    return {"metric": "train rays/s (fwd+bwd, with blur module)", "value": R / (ms * 1e-3), "unit": "rays/s", "ms_fwd_bwd": ms, "rays": R,
            "valid_samples": int(net.last_extras.n_valid), "loss": float(loss.detach()),
            "config": BLUR_WORKLOAD.replace("pre-defined degradation-kernel convolution", "learnable per-patch blur kernel") if learnable else BLUR_WORKLOAD}
