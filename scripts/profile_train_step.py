#!/usr/bin/env python
"""torch.profiler view of one training step (configs[2]): which CUDA kernels -- ours and torch's -- take the time, and where
the GPU idles.  Writes a kernel table to stdout and a gzipped chrome trace to gpurun_out/train_trace.json.gz."""
import gzip, os, shutil, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridneuralrendering_b200.benchmarks import build_train_case
from hybridneuralrendering_b200.renderer import training_loss

dev = torch.device("cuda:0")
net, frame = build_train_case(dev)
params = [p for p in net.parameters() if p.requires_grad]
def step():
    for p in params: p.grad = None
    out = net(**frame)
    loss = training_loss(out, frame["gt_image"])
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); step(); e.record(); torch.cuda.synchronize()
print("unprofiled step ms:", s.elapsed_time(e))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=60))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
tr = os.path.join(ROOT, "gpurun_out", "train_trace.json")
prof.export_chrome_trace(tr)
with open(tr, "rb") as f, gzip.open(tr + ".gz", "wb") as g:
    shutil.copyfileobj(f, g)
os.remove(tr)
