#!/usr/bin/env python
"""Micro-benchmark of the dense-layer kernels at the training-step size (602,192 x 256 x 256): forward, data gradient,
weight gradient."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridneuralrendering_b200 import ops

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

M, N, K = 602192, 256, 256
x = torch.randn(M, K, device="cuda", requires_grad=True)
W = (torch.randn(N, K, device="cuda") * 0.06).requires_grad_(True)
b = torch.zeros(N, device="cuda", requires_grad=True)
gy = torch.randn(M, N, device="cuda")
with torch.no_grad():
    t = timeit(lambda: ops.linear([x], W, b, 1))
print(f"fwd   {t:.3f} ms  {2*M*N*K/t/1e9:.1f} TFLOP/s")
y = ops.linear([x], W, b, 1)
def bwd():
    x.grad = None; W.grad = None; b.grad = None
    y.backward(gy, retain_graph=True)
t = timeit(bwd)
print(f"bwd (data+weight) {t:.3f} ms  {4*M*N*K/t/1e9:.1f} TFLOP/s")
