"""stress the generic tensor-core layer at bench-sized shapes (run each case under `timeout`)"""
import sys, time
import torch
sys.path.insert(0, ".")
from hybridneuralrendering_b200 import ops

case = sys.argv[1]
Nv = 200_000
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g)
with torch.no_grad():
    if case == "cfb0":
        srcs, W, b, act, kw = [r(Nv, 280)], r(128, 280) * 0.1, r(128), 1, {}
    elif case == "cfb1":
        srcs, W, b, act, kw = [r(Nv, 128)], r(128, 128) * 0.1, r(128), 1, {}
    elif case == "amw0":
        srcs, W, b, act, kw = [r(4 * Nv, 45), r(Nv, 128), r(4 * Nv, 3)], r(64, 176) * 0.1, r(64), 1, dict(mods=(0, Nv, 0), M=4 * Nv)
    elif case == "amw1":
        srcs, W, b, act, kw = [r(4 * Nv, 64)], r(64, 64) * 0.1, r(64), 1, {}
    elif case == "mix0":
        gg = r(Nv, 128)
        srcs, W, b, act, kw = [gg[:, :45], r(Nv, 45)], r(45, 90) * 0.1, r(45), 1, {}
    elif case == "mix2":
        gg = r(Nv, 128)
        srcs, W, b, act, kw = [r(Nv, 45)], r(45, 45) * 0.1, r(45), 0, dict(res=gg[:, :45])
    elif case == "amw3":
        srcs, W, b, act, kw = [r(4 * Nv, 64)], r(1, 64) * 0.1, r(1), 0, {}
    elif case == "final":
        gg = r(Nv, 128)
        srcs, W, b, act, kw = [r(Nv, 45), gg[:, 45:]], r(3, 128) * 0.1, r(3), 0, {}
    elif case == "nbr":
        srcs, W, b, act, kw = [r(8 * Nv, 256), r(8 * Nv, 7)], r(256, 263) * 0.1, r(256), 1, {}
    y = ops.linear(srcs, W, b, act, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = ops.linear(srcs, W, b, act, **kw)
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) / 5e3
    M = y.shape[0]
    idx = torch.randint(0, M, (2000,), device="cuda")
    full = torch.cat([s if s.shape[0] == M else s.repeat(M // s.shape[0], 1) for s in srcs], 1)
    ref = torch.nn.functional.linear(full[idx].double(), W.double(), b.double())
    ref = torch.nn.functional.leaky_relu(ref, 0.01) if act == 1 else ref
    if "res" in kw:
        ref = ref + kw["res"][idx].double()
    err = float((y[idx].double() - ref).abs().max())
    print(case, "M", M, "ms", round(dt * 1e3, 2), "max err", err, "TFLOP/s(3x)", round(2 * M * W.shape[0] * W.shape[1] / dt / 1e12, 1))
