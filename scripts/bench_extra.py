"""Secondary measurements on one B200 (not the contract line of bench.py): the other BASELINE
configs through the public API, CUDA events, 5 timed repeats after 2 warm-ups, median.

  C1  rec_2d/linear : 2-D flow, 7 rotations, 85 bins (forward step, 1e6 particles)
  C3  rec_nd_1d     : 6-D flow alone (sample + log q) and the density direction log_prob(x)
  C4  rec_nd_2d     : 6-D flow + 15 two-dimensional KDE screens 85 x 85
  C5  classical MENT: 6-D, 25 projections, grid 16^6, one sample-mode update of one table and
                      the density on the grid
  train             : forward + backward of the C3 loss
Writes one JSON object per line.
"""
import json
import os
import statistics
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch

import mentflow_b200 as mf
from mentflow_b200 import workloads

dev = torch.device("cuda")


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def trained_like(m, f=3.0):
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(f)
    return m


def flow_model(wl, d, kind, n_truth=200_000):
    torch.manual_seed(0)
    gen = trained_like(mf.generate.NSFGenerator(d)).to(dev)
    tfs = wl.get("transforms") or [mf.simulate.LinearTransform(m.to(dev)) for m in wl["matrices"]]
    if kind == "1d":
        diag = mf.diagnostics.Histogram1D(axis=0, edges=wl["edges"], bandwidth=0.5).to(dev)
    else:
        diag = mf.diagnostics.Histogram2D(axis=wl["axis"], edges=wl["edges"], bandwidth=(0.5, 0.5)).to(dev)
    diags = [[diag] for _ in tfs]
    truth = workloads.gaussian_mixture(n_truth, ndim=d, seed=1, device=dev)
    with torch.no_grad():
        meas = mf.simulate.forward(truth, tfs, diags)
    meas = [[m[0].detach()] for m in meas]
    prior = mf.prior.Gaussian(ndim=d, scale=3.0)
    return mf.MENTFlow(transforms=tfs, diagnostics=diags, measurements=meas, generator=gen, prior=prior,
                       entropy_estimator=mf.entropy.MonteCarloEntropyEstimator(prior=prior),
                       discrepancy_function=mf.loss.kl_divergence, penalty_parameter=25.0)


def out(name, n, ms, **kw):
    print(json.dumps({"case": name, "particles": n, "ms": round(ms, 4), "particles_per_s": n / ms * 1e3, **kw}), flush=True)


N = 1_000_000
# ---- C1
m1 = flow_model(workloads.rotations_2d(7, 85, 3.5), 2, "1d")
with torch.no_grad():
    out("C1 rec_2d/linear: 2-D NSF + 7 x KDE-1D(85) + KL, forward", N, timed(lambda: m1.loss(N)))
# ---- C1n: rec_2d/nonlinear (experiments/rec_2d/nonlinear/setup.py:24-44, config rec_2d_nonlinear_flow.yaml):
#      sextupole kicks of strength linspace(-1.5, 1.5, num) followed by a 90 degree rotation
import math
import numpy as np
rot = mf.simulate.rotation_matrix(math.radians(90.0)).float().to(dev)
for num in (4, 7):
    wl_n = {"edges": torch.linspace(-3.5, 3.5, 86), "transforms": [
        mf.simulate.CompositeTransform(mf.simulate.MultipoleTransform(order=3, strength=float(st)),
                                       mf.simulate.LinearTransform(rot)) for st in np.linspace(-1.5, 1.5, num)]}
    m1n = flow_model(wl_n, 2, "1d")
    with torch.no_grad():
        out(f"C1n rec_2d/nonlinear: 2-D NSF + {num} x (sextupole kick -> rotation -> KDE-1D(85)) + KL, forward", N,
            timed(lambda: m1n.loss(N)))
    del m1n
# ---- C3 pieces
m3 = flow_model(workloads.isotropic_1d(6, 100, 64, 3.5), 6, "1d")
g3 = m3.generator
z = torch.randn(N, 6, device=dev)
with torch.no_grad():
    out("C3 flow only: 6-D NSF sample + log q", N, timed(lambda: g3.forward_and_log_prob(z)))
    x = g3.forward(z)
    n_inv = 200_000
    out("C3 density direction: log_prob(x), 6 sweeps per layer", n_inv, timed(lambda: g3.log_prob(x[:n_inv])))
    m25 = flow_model(workloads.isotropic_1d(6, 25, 64, 3.5), 6, "1d")
    out("C3 25 projections: 6-D NSF + 25 x KDE-1D(64) + KL, forward", N, timed(lambda: m25.loss(N)))
    out("C3 100 projections: 6-D NSF + 100 x KDE-1D(64) + KL, forward (eager launches)", N, timed(lambda: m3.loss(N)))
    for nb in (25_000, 100_000):
        from mentflow_b200.graphs import GraphedLoss
        gl = GraphedLoss(m3, nb)
        out(f"C3 100 projections at the reference batch size, CUDA graph replay", nb, timed(lambda: gl(None), reps=9))
        out(f"C3 100 projections at the reference batch size, eager", nb, timed(lambda: m3.loss(nb), reps=9))
# ---- train step
params = list(m3.parameters())


def train():
    for p in params:
        p.grad = None
    L, H, D = m3.loss(N)
    L.backward()


out("C3 training step: forward + backward to all flow parameters", N, timed(train, reps=3, warm=1))
from mentflow_b200.graphs import GraphedTrainStep
for nb in (25_000, 100_000, 1_000_000):
    opt = torch.optim.AdamW(m3.parameters(), lr=1e-5, weight_decay=0.0, capturable=True, fused=True)
    gts = GraphedTrainStep(m3, opt, nb)
    out("C3 optimisation step (zero_grad + loss + backward + AdamW) as one CUDA-graph replay", nb, timed(gts, reps=7))

    def eager_step():
        opt.zero_grad(set_to_none=True)
        L, H, D = m3.loss(nb)
        L.backward()
        opt.step()

    out("C3 optimisation step, eager launches", nb, timed(eager_step, reps=7))
# ---- C4
m4 = flow_model(workloads.corner_2d(6, 85, 3.5), 6, "2d", n_truth=100_000)
with torch.no_grad():
    out("C4 rec_nd_2d: 6-D NSF + 15 x KDE-2D(85x85) + KL, forward", N, timed(lambda: m4.loss(N), reps=3, warm=1))
    x4 = m4.generator.forward(z)
    out("C4 screens only: 15 x KDE-2D(85x85)", N,
        timed(lambda: mf.simulate.forward(x4, m4.transforms, m4.diagnostics), reps=3, warm=1))
from mentflow_b200.graphs import GraphedLoss as _GL
for nb in (30_000, 1_000_000):      # 30,000: the reference's rec_nd_2d batch size
    g4 = _GL(m4, nb)
    out("C4 rec_nd_2d forward as one CUDA-graph replay", nb, timed(lambda: g4(None), reps=5))
    del g4


def train4():
    for p in m4.parameters():
        p.grad = None
    L, H, D = m4.loss(N)
    L.backward()


out("C4 training step: forward + backward to all flow parameters", N, timed(train4, reps=3, warm=1))
# ---- C5
d, K, res, xmax = 6, 25, 16, 3.5
wl5 = workloads.isotropic_1d(d, K, 64, xmax)
tfs = [mf.simulate.LinearTransform(mm.to(dev)) for mm in wl5["matrices"]]
diag = mf.diagnostics.Histogram1D(axis=0, edges=wl5["edges"], bandwidth=0.5).to(dev)
truth = workloads.gaussian_mixture(200_000, ndim=d, seed=1, device=dev)
with torch.no_grad():
    meas = [[p[0]] for p in mf.simulate.forward(truth, tfs, [[diag] for _ in tfs])]
sampler = mf.sample.GridSampler(limits=d * [(-xmax, xmax)], shape=tuple(d * [res]), device="cuda")
n5 = 10_000_000
ment = mf.ment.MENT(ndim=d, transforms=tfs, diagnostics=[[diag] for _ in tfs], measurements=meas,
                    prior=mf.prior.Gaussian(ndim=d, scale=3.0), mode="sample", sampler=sampler, n_samples=n5,
                    device="cuda")
out("C5 MENT density on the 16^6 grid (25 tables)", res ** d, timed(lambda: ment.prob_on_grid(sampler), reps=3, warm=1),
    unit="grid points")
out("C5 MENT one measurement update: grid density + 1e7 samples + projection + KDE", n5,
    timed(lambda: ment.simulate(0, 0), reps=3, warm=1))
out("C5 MENT gauss_seidel_update, all 25 measurements", n5 * K, timed(lambda: ment.gauss_seidel_update(lr=0.9), reps=2, warm=1),
    unit="sampled particles")
