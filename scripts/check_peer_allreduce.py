#!/usr/bin/env python
"""torchrun --nproc-per-node N scripts/check_peer_allreduce.py -- the own NVLink peer-memory all-reduce (csrc/peer.cu,
parallel.PeerBucket) against NCCL's all_reduce on the same data, and their device times at the small-bucket size."""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridneuralrendering_b200 import parallel

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(rank)
shapes = [(256, 284), (256,), (256, 256), (256,), (256, 263), (256,), (256, 256), (256,), (1, 256), (1,)] + [(128, 280), (128,)] + [(64, 176), (64,)] * 3 + [(45, 90), (45,)] * 3
grads = [torch.randn(s, device=dev) for s in shapes]
ref = [g.clone() for g in grads]
for g in ref:
    dist.all_reduce(g)
pb = parallel.peer_bucket(sum(g.numel() for g in grads), dev)
assert pb is not None, "peer memory unavailable"
res = {}
for mode in ("peer", "multimem"):
    if mode == "multimem" and not pb.multicast:
        continue
    mine = [g.clone() for g in grads]
    pb.allreduce_(mine, use_multimem=(mode == "multimem"))
    torch.cuda.synchronize()
    err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(mine, ref))
    # replicas must agree bit for bit: compare against rank 0's result
    flat = torch.cat([m.reshape(-1) for m in mine])
    r0 = flat.clone()
    dist.broadcast(r0, 0)
    same = bool(torch.equal(flat, r0))
    def run():
        pb.allreduce_(mine, use_multimem=(mode == "multimem"))
    for _ in range(5):
        run()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        run()
    e.record(); torch.cuda.synchronize()
    res[mode] = (err, same, s.elapsed_time(e) / 20)
flatn = torch.cat([g.reshape(-1) for g in grads])
for _ in range(5):
    dist.all_reduce(flatn)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    dist.all_reduce(flatn)
e.record(); torch.cuda.synchronize()
t_nccl = s.elapsed_time(e) / 20
allres = [None] * world
dist.all_gather_object(allres, res)
if rank == 0:
    print(f"world {world}, {flatn.numel()} floats, multicast {'yes' if pb.multicast else 'no'}; NCCL all_reduce (flat, back to back): {t_nccl * 1e3:.1f} us")
    for mode in res:
        print(f"  {mode:9s}: max rel. diff vs NCCL {max(r[mode][0] for r in allres):.2e}, replicas bit-identical: {all(r[mode][1] for r in allres)}, "
              f"{max(r[mode][2] for r in allres) * 1e3:.1f} us per call (copy in + barrier + kernel + barrier + copy out)")
dist.destroy_process_group()
