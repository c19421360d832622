"""The bench workload launched eagerly for ncu: two forward steps, then two training steps
(forward + backward to all flow parameters) of 1e6 particles.  Kernel order per forward step:
5 x nsf_tc_layer_kernel<..,0>, moments, kde1d_deposit, kde1d_finish; per training step additionally
kde1d_finish_bwd, kde1d_bwd and per layer (last to first) nsf_tc_layer_kernel<..,1>, dgrad, wgrad."""
import argparse
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch

import bench

args = argparse.Namespace(particles=1_000_000, ndim=6, num_proj=100, bins=64)
dev = torch.device("cuda")
model, _ = bench.build_model(args, dev)
z = torch.randn(args.particles, args.ndim, device=dev)
with torch.no_grad():
    for _ in range(2):
        x, lq = model.generator.forward_and_log_prob(z)
        model.loss_from_particles(x, lq)
torch.cuda.synchronize()
for _ in range(2):
    for p in model.generator.parameters():
        p.grad = None
    x, lq = model.generator.forward_and_log_prob(z)
    L, H, D = model.loss_from_particles(x, lq)
    L.backward()
torch.cuda.synchronize()
print("loss", float(L))
