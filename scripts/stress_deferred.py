"""Stress check of the deferred gradient tails (ops.defer_weight_gradients, parallel.train_step): the gradients of a train_step whose
weight-gradient / image-branch tails were parked and issued on two streams must equal those of a plain loss.backward(), run after
run.  Prints, per iteration, the worst parameter and its max |a-b| / max |b|; `--poison` fills the caching allocator's free blocks
with NaN between runs, so that a read of uninitialised memory shows up as NaN instead of passing by luck.

    python scripts/stress_deferred.py [--iters 30] [--poison]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hybridneuralrendering_b200 import NeuralPoints, NeuralPointsRayMarching, PointAggregator, make_opt, parallel  # noqa: E402
from hybridneuralrendering_b200 import synthetic as syn  # noqa: E402
from hybridneuralrendering_b200.optim import FusedAdam  # noqa: E402
from hybridneuralrendering_b200.renderer import training_loss  # noqa: E402


def poison():
    """hand every cached free block out once, filled with NaN, and give it back"""
    held = []
    for nbytes in (256 << 20, 32 << 20, 4 << 20, 1 << 20, 512 << 10, 64 << 10, 8 << 10, 512):
        for _ in range(4096):
            before = torch.cuda.memory_reserved()
            t = torch.empty(nbytes // 4, device="cuda", dtype=torch.float32)
            if torch.cuda.memory_reserved() > before:         # served by a new segment: the cache has no block of this size left
                del t
                break
            t.fill_(float("nan"))
            held.append(t)
    del held
    torch.cuda.synchronize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--poison", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda")
    c = lambda a: torch.from_numpy(a).to(dev)
    opt = make_opt("scannet", use_nearest=2, SR=24, is_train=True, drop_ratio=0.5, dilation_setup="4_4_1_8")
    xyz = syn.room_scene(30000, 9)
    att = syn.point_attributes(np.random.default_rng(9), len(xyz))
    fr = syn.room_frame(H=48, W=64, V=2, patch_num=4, patch_size=4, seed=5)
    P = syn.random_aggregator_params(10)
    frame = {k: (c(v) if isinstance(v, np.ndarray) and v.dtype.kind == "f" else v) for k, v in fr.items()}

    def build():
        pts = NeuralPoints(32, len(xyz), opt, dev)
        pts.set_points(c(xyz), c(att["emb"])[None], points_color=c(att["color"])[None], points_dir=c(att["dir"])[None],
                       points_conf=c(att["conf"])[None], parameter=True)
        agg = PointAggregator(opt).to(dev)
        agg.load_state_dict(P, strict=False)
        net = NeuralPointsRayMarching(aggregator=agg, neural_points=pts, opt=opt).to(dev)
        net.near_far = (0.1, 8.0)
        return net

    def run(deferred):
        net = build()
        torch.manual_seed(3)
        if args.poison:
            poison()
        if deferred:
            loss, _ = parallel.train_step(net, frame, [FusedAdam([p for p in net.parameters() if p.requires_grad], lr=0.0)])
        else:
            loss = training_loss(net(**frame), frame["gt_image"])
            loss.backward()
        torch.cuda.synchronize()
        return float(loss), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}

    lb, gb = run(False)
    worst_all = 0.0
    for it in range(args.iters):
        la, ga = run(True)
        l2, g2 = run(False)
        rows = []
        for k in gb:
            a, b, b2 = ga[k].double(), gb[k].double(), g2[k].double()
            den = float(b.abs().max()) + 1e-30
            rows.append((float((a - b).abs().max()) / den, float((b2 - b).abs().max()) / den, bool(torch.isnan(a).any()), k))
        rows.sort(reverse=True)
        worst_all = max(worst_all, rows[0][0])
        nan = [r[3] for r in rows if r[2]]
        print(f"it {it:3d} loss deferred/plain/first {la!r} {l2!r} {lb!r} | worst deferred-vs-plain {rows[0][0]:.3e} ({rows[0][3]}), "
              f"plain-vs-plain worst {max(r[1] for r in rows):.3e} | second {rows[1][0]:.3e} ({rows[1][3]}) | NaN in {nan}", flush=True)
    print(f"worst over {args.iters} iterations: {worst_all:.3e}  (test tolerance 1e-5)")


if __name__ == "__main__":
    main()
