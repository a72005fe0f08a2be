#!/usr/bin/env python
"""Dense TF32 tensor throughput of this GPU, measured the way MEASURED_PEAKS.json measures bf16 (SURVEY.md 8d asks for it: the
3xTF32 backward kernels are compared against bf16_sustained / 2 "derived" until this number exists).

    python scripts/measure_tf32_peak.py [--json profiles/r2_tf32_peak.json]      (the file bench.py reads the measured TF32 peak from)

torch.matmul fp32 8192^3 with TF32 enabled: best of 10 (burst) and back to back for 4 s (sustained), CUDA events."""
import argparse
import json
import time

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--n", type=int, default=8192)
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = True
    dev = torch.device("cuda:0")
    n = args.n
    a, b = torch.randn(n, n, device=dev), torch.randn(n, n, device=dev)
    c = torch.empty(n, n, device=dev)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    flop = 2.0 * n ** 3
    best = 1e9
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); torch.matmul(a, b, out=c); e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, it = time.perf_counter(), 0
    s.record()
    while time.perf_counter() - t0 < 4.0:
        for _ in range(20):
            torch.matmul(a, b, out=c)
        it += 20
        torch.cuda.synchronize()
    e.record()
    torch.cuda.synchronize()
    out = {"tf32_tflops": flop / (best * 1e-3) / 1e12, "tf32_tflops_sustained": flop * it / (s.elapsed_time(e) * 1e-3) / 1e12, "n": n,
           "how": "torch.matmul fp32 with allow_tf32, best of 10 (burst) and back to back for 4 s (sustained), CUDA events",
           "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    print(json.dumps(out))
    if args.json:
        with open(args.json, "w") as f:
            f.write(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
