#!/usr/bin/env python
"""Raw NCCL all-reduce timings on this box (device time, CUDA events, median of 7) for the message sizes of the gradient exchange."""
import os, sys
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W = dist.get_world_size()
def timeit(fn, n=7):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]
out = []
for mb in (1.8, 8, 24, 256, 312, 1248):
    t = torch.ones(int(mb * 1e6 / 4), device=dev)
    ms = timeit(lambda: dist.all_reduce(t))
    out.append(f"{mb} MB: {ms:.3f} ms (algbw {mb / ms:.0f} GB/s, busbw {2 * (W - 1) / W * mb / ms:.0f} GB/s)")
    del t
# four tensors back to back (the point tables as separate parameters) vs one flat buffer
ts = [torch.ones(int(mb * 1e6 / 4), device=dev) for mb in (256, 8, 24, 24)]
def four():
    hs = [dist.all_reduce(t, async_op=True) for t in ts]
    for h in hs: h.wait()
out.append(f"4 tensors (256+8+24+24 MB) async: {timeit(four):.3f} ms")
if dist.get_rank() == 0:
    print(f"world {W}:\n  " + "\n  ".join(out))
dist.destroy_process_group()
