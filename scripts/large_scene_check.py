#!/usr/bin/env python
"""BASELINE.json configs[4] shape check on one GPU: 8M neural points (12 x 10 x 3 m room), 1296x968 frame geometry,
4096-ray training batch with V=8 reference views: grid build, query, fwd+bwd, and a full-frame render."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridneuralrendering_b200 import NeuralPoints, NeuralPointsRayMarching, PointAggregator, make_opt
from hybridneuralrendering_b200 import synthetic as syn
from hybridneuralrendering_b200.renderer import render_rays, training_loss

dev = torch.device("cuda:0")
N, H, W, V = 8_000_000, 968, 1296, 8
size = (12.0, 10.0, 3.0)
opt = make_opt("scannet", use_nearest=V, SR=24, is_train=True, drop_ratio=0.5, dilation_setup="8_8_1_8", max_o=4_000_000)
t0 = time.time()
xyz = syn.room_scene(N, 0, size=size)
att = syn.point_attributes(np.random.default_rng(0), len(xyz))
fr = syn.room_frame(H=H, W=W, V=V, patch_num=8, patch_size=8, seed=0, size=size)
print(f"scene built in {time.time() - t0:.1f} s: {len(xyz)} points", flush=True)
c = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
pts = NeuralPoints(32, len(xyz), opt, dev)
pts.set_points(c(xyz), c(att["emb"])[None], points_color=c(att["color"])[None], points_dir=c(att["dir"])[None],
               points_conf=c(att["conf"])[None], parameter=True)
torch.manual_seed(0)
agg = PointAggregator(opt).to(dev)
net = NeuralPointsRayMarching(aggregator=agg, neural_points=pts, opt=opt).to(dev)
net.near_far = (0.1, 8.0)
frame = {k: (c(v) if isinstance(v, np.ndarray) and v.dtype.kind == "f" else v) for k, v in fr.items()}
params = [p for p in net.parameters() if p.requires_grad]

def step():
    for p in params:
        p.grad = None
    out = net(**frame)
    loss = training_loss(out, frame["gt_image"])
    loss.backward()
    return out, loss

for _ in range(3):
    out, loss = step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    out, loss = step()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
ex = net.last_extras
res = {"points": len(xyz), "train_ms_fwd_bwd": ms, "train_rays_per_s": 4096 / ms * 1e3, "kept_rays": int(ex.n_rays), "valid_samples": int(ex.n_valid),
       "loss": float(loss.detach()), "grad_finite": bool(all(torch.isfinite(p.grad).all() for p in params if p.grad is not None)),
       "mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}
# full frame render (inference) at 1296x968
opt.is_train = False
px, py = syn.full_frame_pixels(H, W)
Kmat = fr["intrinsic"][0]
rd = syn.rays_for_pixels(px, py, Kmat, fr["c2w"][0])[None]
fframe = {k: frame[k] for k in ("campos", "camrotc2w", "near", "far", "intrinsic", "bg_color", "images_nearest", "c2w_nearest", "campos_nearest", "intrinsic_nearest")}
fframe["raydir"] = c(rd.astype(np.float32))
img = render_rays(net, fframe); torch.cuda.synchronize()
s.record(); img = render_rays(net, fframe); e.record(); torch.cuda.synchronize()
res.update({"frame_ms": s.elapsed_time(e), "frame_Mpix_s": H * W / s.elapsed_time(e) / 1e3, "frame_finite": bool(torch.isfinite(img).all()),
            "frame_valid_samples": int(net.last_extras.n_valid), "mem_GB_total": torch.cuda.max_memory_allocated() / 2 ** 30})
print(json.dumps(res))
