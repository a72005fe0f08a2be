#!/usr/bin/env python
"""cProfile of the host side of the training step (configs[2]): where do the ~8 ms of Python / dispatcher time per step go?"""
import cProfile, os, pstats, sys, io
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridneuralrendering_b200 import benchmarks, parallel
dev = torch.device("cuda:0")
net, frame = benchmarks.build_train_case(dev)
opts = benchmarks.make_optimizers(net)
for _ in range(5):
    parallel.train_step(net, frame, opts, next_frame_shard=frame)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    parallel.train_step(net, frame, opts, next_frame_shard=frame)
pr.disable()
torch.cuda.synchronize()
for key in ("tottime", "cumtime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue()[:9000])
