"""Two eager optimisation steps at the reference batch size (25,000 particles) for an ncu launch list."""
import argparse
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch

import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25_000
args = argparse.Namespace(particles=n, ndim=6, num_proj=100, bins=64)
dev = torch.device("cuda")
model, _ = bench.build_model(args, dev)
opt = torch.optim.AdamW(model.generator.parameters(), lr=1e-3, capturable=True)
for _ in range(3):
    opt.zero_grad(set_to_none=True)
    L, H, D = model.loss(n)
    L.backward()
    opt.step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
opt.zero_grad(set_to_none=True)
L, H, D = model.loss(n)
L.backward()
opt.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
