"""One classical-MENT measurement update at the C5 size for an ncu launch list (bench.ment_step_measure's model)."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch

import mentflow_b200 as mf
from mentflow_b200 import workloads

dev = torch.device("cuda")
d, xmax, num_proj, bins, res, n = 6, 3.5, 25, 64, 16, 12_500_000
wl = workloads.isotropic_1d(d, num_proj, bins, xmax)
tfs = [mf.simulate.LinearTransform(m.to(dev)) for m in wl["matrices"]]
diag = mf.diagnostics.Histogram1D(axis=0, edges=wl["edges"], bandwidth=0.5).to(dev)
diags = [[diag] for _ in tfs]
truth = workloads.gaussian_mixture(200_000, ndim=d, seed=1, device=dev)
with torch.no_grad():
    meas = [[p[0]] for p in mf.simulate.forward(truth, tfs, diags)]
sampler = mf.sample.GridSampler(limits=d * [(-xmax, xmax)], shape=tuple(d * [res]), device=dev)
m = mf.ment.MENT(ndim=d, transforms=tfs, diagnostics=diags, measurements=meas, prior=mf.prior.Gaussian(ndim=d, scale=3.0),
                 mode="sample", sampler=sampler, n_samples=n, device=dev)
torch.manual_seed(77)
m.gauss_seidel_update(lr=0.9)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.simulate(0, 0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
