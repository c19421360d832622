"""Time of one optimisation step (GraphedTrainStep: zero_grad + loss + backward + fused AdamW, one graph replay)
at small batch sizes: python scripts/opt_step_time.py [n ...]"""
import argparse
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch

import bench
import mentflow_b200 as mf

sizes = [int(a) for a in sys.argv[1:]] or [25_000, 100_000]
dev = torch.device("cuda")
for n in sizes:
    args = argparse.Namespace(particles=n, ndim=6, num_proj=100, bins=64)
    model, _ = bench.build_model(args, dev)
    opt = torch.optim.AdamW(model.generator.parameters(), lr=1e-3, capturable=True, fused=True)
    step = mf.graphs.GraphedTrainStep(model, opt, n)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            step()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b) / 20)
    print(f"n={n}: {best:.4f} ms per optimisation step, loss {float(step()[0]):.5f}")
