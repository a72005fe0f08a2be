#!/usr/bin/env python
"""Micro-benchmark of the fused per-sample chains (chain_f16.cu) and the small-N colour head at the
per-pass sizes of the lego frame (262,144 valid samples, V=4)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridneuralrendering_b200 import chain, ops  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def mk(widths, kin):
    layers = []
    for w in widths:
        layers.append(torch.nn.Linear(kin, w).cuda())
        kin = w
    return layers


def main():
    Nv, V = 262144, 4
    g = torch.randn(Nv, 128, device="cuda")
    X5 = torch.randn(Nv, 280, device="cuda")
    aux = torch.randn(V * Nv, 45, device="cuda")
    dv = torch.randn(V * Nv, 3, device="cuda")
    merged = torch.randn(Nv, 45, device="cuda")
    with torch.no_grad():
        cf = chain.PackedChain(mk((128, 128, 128), 280), [1, 1, 1], 280)
        am_l = mk((64, 64, 64), 176)
        am = chain.PackedChain(am_l, [1, 1, 1], 176)
        hw, hb = torch.randn(1, 64, device="cuda"), torch.zeros(1, device="cuda")
        cm = chain.PackedChain(mk((45, 45, 45), 90), [1, 1, 0], 90)
        fin = torch.nn.Linear(128, 3).cuda()
        mix = torch.randn(Nv, 45, device="cuda")
        t = timeit(lambda: chain.chain_forward(cf, [X5]))
        fl = 2 * (280 * 128 + 2 * 128 * 128) * Nv
        print(f"cf  chain: {t:.3f} ms  {fl / t / 1e9:.1f} TFLOP/s-equiv  in+out {(280 + 128) * 4 * Nv / t / 1e6:.0f} GB/s")
        t = timeit(lambda: chain.chain_forward(am, [g, aux, dv], M=V * Nv, mods=(Nv, 0, 0), out=False, head=(hw, hb, 2)))
        fl = 2 * (176 * 64 + 2 * 64 * 64 + 64) * V * Nv
        print(f"am  chain: {t:.3f} ms  {fl / t / 1e9:.1f} TFLOP/s-equiv  in {(176) * 4 * V * Nv / t / 1e6:.0f} GB/s")
        aux48 = torch.randn(V * Nv, 48, device="cuda")
        t = timeit(lambda: chain.chain_forward(am, [g, aux48], M=V * Nv, mods=(Nv, 0), out=False, head=(hw, hb, 2)))
        print(f"am  chain (48-wide aligned rows, inference layout): {t:.3f} ms  {fl / t / 1e9:.1f} TFLOP/s-equiv")
        t = timeit(lambda: chain.chain_forward(cm, [g[:, :45], merged], res=g[:, :45]))
        fl = 2 * (90 * 45 + 2 * 45 * 45) * Nv
        print(f"cm  chain: {t:.3f} ms  {fl / t / 1e9:.1f} TFLOP/s-equiv")
        t = timeit(lambda: ops.linear([mix, g[:, 45:]], fin.weight, fin.bias, 3))
        print(f"colour head (128->3): {t:.3f} ms  {128 * 4 * Nv / t / 1e6:.0f} GB/s")
        t = timeit(lambda: X5.clone())
        print(f"copy of X5 (ref): {t:.3f} ms  {2 * 280 * 4 * Nv / t / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
