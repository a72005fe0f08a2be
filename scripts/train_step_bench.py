#!/usr/bin/env python
"""Training-step timing on BASELINE.json configs[2] (see hybridneuralrendering_b200/benchmarks.py).

    python scripts/train_step_bench.py [--steps 5] [--warmup 3] [--points 2000000] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=2_000_000)
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--json", default=None)
    ap.add_argument("--no-prefetch", action="store_true", help="do not launch the next step's voxel query ahead (A/B of the software pipelining)")
    args = ap.parse_args()
    from hybridneuralrendering_b200.benchmarks import train_step_benchmark
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    line = train_step_benchmark(dev, args.steps, args.warmup, points=args.points, views=args.views, prefetch=not args.no_prefetch)
    print(json.dumps(line))
    if args.json:
        with open(args.json, "w") as f:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
