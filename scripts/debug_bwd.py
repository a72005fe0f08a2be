import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridneuralrendering_b200 import ops
T = torch.from_numpy
for (M, N, ks, act) in [(4096, 256, (256, 7), 1), (4096, 256, (263,), 1), (4096, 256, (256,), 1), (1000, 256, (284,), 1)]:
    for trial in range(3):
        rng = np.random.default_rng(M + N)
        K = sum(ks)
        srcs = [T(rng.standard_normal((M, k)).astype(np.float32)).cuda().requires_grad_(True) for k in ks]
        W = T((rng.standard_normal((N, K)) * 0.1).astype(np.float32)).cuda().requires_grad_(True)
        b = T(rng.standard_normal(N).astype(np.float32)).cuda().requires_grad_(True)
        gy = T(rng.standard_normal((M, N)).astype(np.float32)).cuda()
        y = ops.linear(srcs, W, b, act)
        y.backward(gy)
        sd = [s.detach().double().requires_grad_(True) for s in srcs]
        Wd, bd = W.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
        yd = torch.nn.functional.leaky_relu(torch.nn.functional.linear(torch.cat(sd, 1), Wd, bd), 0.01)
        yd.backward(gy.double())
        out = []
        for name, a, r in [("y", y, yd)] + [(f"dsrc{i}", s.grad, d.grad) for i, (s, d) in enumerate(zip(srcs, sd))] + [("dW", W.grad, Wd.grad), ("db", b.grad, bd.grad)]:
            err = (a.detach().double() - r.detach()).abs()
            bad = err > 1e-4 * r.detach().abs().max()
            loc = torch.nonzero(bad)
            out.append(f"{name}: max {float(err.max()):.2e} bad {int(bad.sum())}" + (f" rows {sorted(set(loc[:,0].tolist()))[:6]} cols {sorted(set(loc[:,1].tolist()))[:8] if loc.shape[1] > 1 else ''}" if int(bad.sum()) else ""))
        print((M, N, ks), trial, " | ".join(out), flush=True)
