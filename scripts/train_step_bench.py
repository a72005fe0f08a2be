#!/usr/bin/env python
"""Training-step timing on BASELINE.json configs[2]: ScanNet-shaped hybrid training step
(640x480 frames, 4096-ray batch = 8x8 dilated patches of 8x8, V=8 reference-view feature maps,
2M neural points, SR=24, K=8), forward + loss + backward (+ Adam reported separately).

    python scripts/train_step_bench.py [--steps 5] [--warmup 3] [--points 2000000] [--json out.json]

Prints one JSON line: train rays/s for fwd+bwd, stage split from CUDA events on the launching stream.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=2_000_000)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    from hybridneuralrendering_b200 import NeuralPoints, NeuralPointsRayMarching, PointAggregator, make_opt, ops
    from hybridneuralrendering_b200 import synthetic as syn
    from hybridneuralrendering_b200.renderer import training_loss

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    V = args.views
    opt = make_opt("scannet", use_nearest=V, SR=24, is_train=True, drop_ratio=0.5, dilation_setup="8_8_1_8")
    xyz = syn.room_scene(args.points, 0)
    att = syn.point_attributes(np.random.default_rng(0), len(xyz))
    fr = syn.room_frame(H=480, W=640, V=V, patch_num=8, patch_size=8, seed=0)
    c = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pts = NeuralPoints(32, len(xyz), opt, dev)
    pts.set_points(c(xyz), c(att["emb"])[None], points_color=c(att["color"])[None], points_dir=c(att["dir"])[None],
                   points_conf=c(att["conf"])[None], parameter=True)
    torch.manual_seed(0)
    agg = PointAggregator(opt).to(dev)
    net = NeuralPointsRayMarching(aggregator=agg, neural_points=pts, opt=opt).to(dev)
    frame = {k: (c(v) if isinstance(v, np.ndarray) and v.dtype.kind == "f" else v) for k, v in fr.items()}
    R = fr["raydir"].shape[1]
    params = [p for p in net.parameters() if p.requires_grad]
    opt_net = torch.optim.Adam([p for n, p in net.named_parameters() if p.requires_grad and not n.startswith("neural_points.")], lr=5e-4)
    opt_pts = torch.optim.Adam([p for n, p in net.named_parameters() if p.requires_grad and n.startswith("neural_points.")], lr=2e-3)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    def fwd_bwd():
        for p in params:
            p.grad = None
        out = net(**frame)
        loss = training_loss(out, frame["gt_image"])
        with ops.tag('backward'):
            loss.backward()
        return out, loss

    for _ in range(args.warmup):
        fwd_bwd()
        opt_net.step(); opt_pts.step()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
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
    t_fb = sum(s.elapsed_time(m) for s, m, e in ev) / args.steps
    t_opt = sum(m.elapsed_time(e) for s, m, e in ev) / args.steps
    # stage split (per-launch events)
    ops.TIMERS = []
    fwd_bwd()
    torch.cuda.synchronize()
    stages = {}
    for tag, s, e in ops.TIMERS:
        stages[tag] = stages.get(tag, 0.0) + s.elapsed_time(e)
    ops.TIMERS = None
    ex = net.last_extras
    line = {"metric": "train rays/s (fwd+bwd)", "value": R / (t_fb * 1e-3), "unit": "rays/s", "ms_fwd_bwd": t_fb, "ms_adam": t_opt,
            "rays": R, "kept_rays": int(ex.n_rays), "valid_samples": int(ex.n_valid), "valid_neighbours": agg.last_valid_neighbours(),
            "points": len(xyz), "views": V, "loss": float(loss), "launches_per_step": launches // args.steps,
            "stage_ms": {k: round(v, 3) for k, v in sorted(stages.items())},
            "config": "ScanNet scene0241_01-shaped hybrid training step: 640x480, 4096-ray batch, 8 reference views, 2M points"}
    print(json.dumps(line))
    if args.json:
        with open(args.json, "w") as f:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
