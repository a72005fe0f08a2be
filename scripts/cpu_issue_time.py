#!/usr/bin/env python
"""Host-side view of the training step (configs[2]): how long the CPU needs to ISSUE one fwd+bwd (no device synchronisation inside the
loop), which calls synchronise with the device (torch's sync debug mode), and the GPU-side step time next to it.  If issue time >= GPU
time the step is host-bound and faster kernels do not show."""
import os, sys, time, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridneuralrendering_b200.benchmarks import build_train_case
from hybridneuralrendering_b200.renderer import training_loss
from hybridneuralrendering_b200.optim import FusedAdam

dev = torch.device("cuda:0")
net, frame = build_train_case(dev)
params = [p for p in net.parameters() if p.requires_grad]
opt_net = torch.optim.Adam([p for n, p in net.named_parameters() if p.requires_grad and not n.startswith("neural_points.")], lr=5e-4)
opt_pts = FusedAdam([p for n, p in net.named_parameters() if p.requires_grad and n.startswith("neural_points.")], lr=2e-3)
prefetch = "--no-prefetch" not in sys.argv


def step():
    for p in params:
        p.grad = None
    out = net(**frame)
    loss = training_loss(out, frame["gt_image"])
    if prefetch:
        net.prefetch_query(**frame)
    loss.backward()
    opt_net.step(); opt_pts.step()


for _ in range(4):
    step()
torch.cuda.synchronize()
N = 10
t0 = time.perf_counter()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(N):
    step()
t_issue = (time.perf_counter() - t0) / N * 1e3
e.record()
torch.cuda.synchronize()
print(f"host issue time per step (incl. optimiser): {t_issue:.2f} ms; device time per step: {s.elapsed_time(e) / N:.2f} ms; prefetch={prefetch}")
torch.cuda.set_sync_debug_mode(1)
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    step()
torch.cuda.set_sync_debug_mode(0)
torch.cuda.synchronize()
print("synchronising calls in one step:", len(w))
for x in w:
    print("  ", str(x.message)[:100], "@", x.filename.split("/")[-1], x.lineno)
