#!/usr/bin/env python
"""Timing of parallel.allreduce_gradients on the training parameter set (2M points: 312 MB of point gradients + 1.8 MB of MLP /
conv gradients), per phase.  torchrun --nproc-per-node N scripts/bench_allreduce.py"""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridneuralrendering_b200 import parallel

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 2_000_000
shapes = [(1, N, 32), (1, N, 1), (1, N, 3), (1, N, 3)] + [(256, 284), (256,), (256, 256), (256,), (256, 263), (256,), (256, 256), (256,)] + [(128, 280), (128,)] * 3 + [(64, 176), (64,)] * 4 + [(45, 90), (45,)] * 6
params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in shapes]
def grads():
    for p in params:
        p.grad = torch.ones_like(p)
nv = torch.tensor(4000, device=dev)
def timeit(fn, n=5):
    for _ in range(2):
        grads(); fn()
    ts = []
    for _ in range(n):
        grads(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]
t_all = timeit(lambda: parallel.allreduce_gradients(params, nv))
big = params[0].grad
t_big = timeit(lambda: dist.all_reduce(params[0].grad))
t_mul = timeit(lambda: [p.grad.mul_(0.5) for p in params])
if dist.get_rank() == 0:
    print(f"world {dist.get_world_size()}: allreduce_gradients {t_all:.2f} ms | raw all_reduce of the 256 MB embedding gradient {t_big:.2f} ms "
          f"({2 * (dist.get_world_size() - 1) / dist.get_world_size() * big.numel() * 4 / t_big / 1e6:.0f} GB/s bus) | scaling pass {t_mul:.2f} ms")
dist.destroy_process_group()
