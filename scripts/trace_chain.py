#!/usr/bin/env python
"""Event trace (clock64) of CTA 0 of the cf chain: where does a tile's time go?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridneuralrendering_b200 import chain
from hybridneuralrendering_b200._lib import lib, ptr

which = sys.argv[1] if len(sys.argv) > 1 else "cf"
Nv = 262144
with torch.no_grad():
    if which == "cf":
        layers, kin = [], 280
        for w in (128, 128, 128):
            layers.append(torch.nn.Linear(kin, w).cuda()); kin = w
        pc = chain.PackedChain(layers, [1, 1, 1], 280)
        X = torch.randn(Nv, 280, device="cuda")
        run = lambda: chain.chain_forward(pc, [X])
    elif which == "am":
        V = 4
        layers, kin = [], 176
        for w in (64, 64, 64):
            layers.append(torch.nn.Linear(kin, w).cuda()); kin = w
        head = torch.nn.Linear(64, 1).cuda()
        pc = chain.PackedChain(layers, [1, 1, 1], 176)
        g = torch.randn(Nv, 128, device="cuda"); aux = torch.randn(V * Nv, 48, device="cuda")
        run = lambda: chain.chain_forward(pc, [g, aux], M=V * Nv, mods=(Nv, 0), out=False, head=(head.weight, head.bias, 2))
    else:
        layers, kin = [], 90
        for w in (45, 45, 45):
            layers.append(torch.nn.Linear(kin, w).cuda()); kin = w
        pc = chain.PackedChain(layers, [1, 1, 0], 90)
        g = torch.randn(Nv, 128, device="cuda"); mg = torch.randn(Nv, 45, device="cuda")
        run = lambda: chain.chain_forward(pc, [g[:, :45], mg], res=g[:, :45])
    run(); torch.cuda.synchronize()
    buf = torch.zeros(3 * 2 * 4096, dtype=torch.int64, device="cuda")
    lib().hnr_chain_f16_set_trace(ptr(buf))
    run(); torch.cuda.synchronize()
    lib().hnr_chain_f16_set_trace(None)
t = buf.cpu().numpy().reshape(3, 4096, 2)
ev = []
for role in range(3):
    for c, tag in t[role]:
        if c:
            ev.append((int(c), role, int(tag >> 32), int((tag >> 16) & 0xffff), int(tag & 0xffff)))
ev.sort()
t0 = ev[0][0]
names = {1: "mma.ready", 2: "mma.issued", 10: "gen.data", 11: "gen.free", 12: "gen.deliv", 20: "epi.acc", 21: "epi.ld", 22: "epi.done"}
print("events", len(ev))
# per-tile summary: time between successive 'epi.done' of the last layer
last = max(e[3] for e in ev if e[2] == 22)
done = [e[0] - t0 for e in ev if e[2] == 22 and e[3] == last]
print("tile completion times (cycles):", done[:12], "mean period", (done[-1] - done[0]) / max(1, len(done) - 1))
for e in ev[:int(sys.argv[2]) if len(sys.argv) > 2 else 400]:
    print(f"{e[0] - t0:8d} {'  ' * e[1]}{names.get(e[2], e[2]):11s} {e[3]:4d} {e[4]:6d}")
