#!/usr/bin/env python
"""Where a data-parallel training step spends its time on the main stream (configs[2], parallel.train_step's `timeline` marks):
forward, backward, gradient tails, small bucket, and -- inside the next forward -- how long the main stream waits for the
point-table all-reduce (the exposed part of the collective).   torchrun --nproc-per-node N scripts/dp_timeline.py"""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridneuralrendering_b200 import benchmarks, parallel

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
world = int(os.environ.get("WORLD_SIZE", 1))
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank() if world > 1 else 0
net, frames = benchmarks.build_train_case(dev, n_frames=benchmarks.FRAME_SET)
opts = benchmarks.make_optimizers(net)
steps = 12
tl = []
for i in range(steps + 3):
    cur, nx = frames[(rank + i) % len(frames)], frames[(rank + i + 1) % len(frames)]
    parallel.train_step(net, cur, opts, next_frame_shard=nx, timeline=tl if i >= 3 else None)
parallel.flush_pending(net)
torch.cuda.synchronize()
# intervals between consecutive marks, averaged over the steps
acc, cnt = {}, {}
for (n0, e0), (n1, e1) in zip(tl[:-1], tl[1:]):
    k = f"{n0} -> {n1}"
    acc[k] = acc.get(k, 0.0) + e0.elapsed_time(e1)
    cnt[k] = cnt.get(k, 0) + 1
rows = [f"{k:45s} {acc[k] / cnt[k]:7.3f} ms  (x{cnt[k]})" for k in acc]
total = tl[0][1].elapsed_time(tl[-1][1]) / steps
msg = f"rank {rank}/{world}: {total:.3f} ms per step\n  " + "\n  ".join(rows)
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, msg)
    if rank == 0:
        print("\n".join(out[:2] + out[-1:]))
    dist.destroy_process_group()
else:
    print(msg)
