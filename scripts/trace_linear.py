#!/usr/bin/env python
"""clock64 event trace of CTA 0 of the tensor-core dense layer (linear_tc.cu) at the training size."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridneuralrendering_b200 import ops
from hybridneuralrendering_b200._lib import lib, ptr

M, N, K = 602192, 256, 256
x = torch.randn(M, K, device="cuda")
W = torch.randn(N, K, device="cuda") * 0.06
b = torch.zeros(N, device="cuda")
mode = sys.argv[2] if len(sys.argv) > 2 else "fwd"
buf = torch.zeros(3 * 2 * 4096, dtype=torch.int64, device="cuda")
if mode == "fwd":
    with torch.no_grad():
        ops.linear([x], W, b, 1); torch.cuda.synchronize()
        lib().hnr_chain_f16_set_trace(ptr(buf))
        ops.linear([x], W, b, 1); torch.cuda.synchronize()
        lib().hnr_chain_f16_set_trace(None)
elif mode == "wgrad":
    with torch.no_grad():
        y = ops.linear([x], W, b, 1)
        gy = torch.randn_like(y)
        ops.linear_backward(W, y, [x], (), gy, 1, [False], need_w=True); torch.cuda.synchronize()
        lib().hnr_chain_f16_set_trace(ptr(buf))
        ops.linear_backward(W, y, [x], (), gy, 1, [False], need_w=True); torch.cuda.synchronize()
        lib().hnr_chain_f16_set_trace(None)
else:                                   # data gradient only (gated A operand)
    with torch.no_grad():
        y = ops.linear([x], W, b, 1)
        gy = torch.randn_like(y)
        ops.linear_backward(W, y, [x], (), gy, 1, [True], need_w=False); torch.cuda.synchronize()
        lib().hnr_chain_f16_set_trace(ptr(buf))
        ops.linear_backward(W, y, [x], (), gy, 1, [True], need_w=False); torch.cuda.synchronize()
        lib().hnr_chain_f16_set_trace(None)
t = buf.cpu().numpy().reshape(3, 4096, 2)
ev = []
for role in range(3):
    for c, tag in t[role]:
        if c:
            ev.append((int(c), role, int(tag >> 32), int((tag >> 16) & 0xffff), int(tag & 0xffff)))
ev.sort()
t0 = ev[0][0]
names = {1: "mma.A", 3: "mma.acc", 4: "mma.W", 10: "cv.start", 11: "cv.free", 12: "cv.deliv", 20: "epi.acc", 22: "epi.done"}
if mode == "wgrad":
    names = {1: "mma.step", 10: "gen.super", 11: "gen.loaded", 12: "gen.deliv"}
for e in ev[:int(sys.argv[1]) if len(sys.argv) > 1 else 300]:
    print(f"{e[0] - t0:8d} {'  ' * e[1]}{names.get(e[2], e[2]):9s} {e[3]:4d}")
