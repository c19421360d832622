"""GPU parity: classical MENT (density, integration, grid sampling, Gauss-Seidel update) vs
goldens produced by the reference's own ment.py / sample.py."""
import numpy as np
import pytest
import torch

import mentflow_b200 as mf
from mentflow_b200 import ops
from mfb_testutil import cuda, t32
from oracle import hotpath as hp

pytestmark = pytest.mark.gpu


def _model(g, mode="sample", n_samples=30000, res=None):
    mats, edges, meas = t32(g["matrices"]), t32(g["edges"]), t32(g["meas"])
    d = mats.shape[1]
    tfs = [mf.simulate.LinearTransform(m.cuda()) for m in mats]
    diag = mf.diagnostics.Histogram1D(axis=0, edges=edges, bandwidth=0.5).to("cuda")
    kw = {}
    if mode == "sample":
        xmax = float(g["grid_xmax"])
        res = int(g["grid_res"]) if res is None else res
        kw["sampler"] = mf.sample.GridSampler(limits=d * [(-xmax, xmax)], shape=tuple(d * [res]), device="cuda")
        kw["n_samples"] = n_samples
    else:
        lo, hi = [float(v) for v in g["int_limits"]]
        kw["integration_limits"] = [[[(lo, hi)]] for _ in tfs]
        kw["integration_shape"] = [[(int(g["int_shape"]),)] for _ in tfs]
    model = mf.ment.MENT(ndim=d, transforms=tfs, diagnostics=[[diag] for _ in tfs],
                         measurements=[[m.cuda()] for m in meas],
                         prior=mf.prior.Gaussian(ndim=d, scale=float(g["prior_scale"])), mode=mode, device="cuda", **kw)
    return model


def test_prob_matches_reference(golden):
    g = golden("ment_4d")
    model = _model(g)
    for i, t in enumerate(t32(g["tables0"])):
        model.lagrange_functions[i][0].set_values(t.cuda())
    prob = model.prob(cuda(g["xq"])).cpu()
    ref = t32(g["prob_q"])
    assert torch.allclose(prob, ref, rtol=1e-4, atol=1e-9 * float(ref.max()))
    assert (prob[:4] == 0).all() or torch.allclose(prob[:4], ref[:4], rtol=1e-4, atol=1e-12)
    grid = model.prob_on_grid(model.sampler).cpu()
    refg = t32(g["prob_grid"])
    # grid points come from the index (lo + i*step) instead of fp32 linspace centres: 1-ulp shifts
    assert torch.allclose(grid, refg, rtol=1e-4, atol=2e-6 * float(refg.max()))
    # generic prob_func path of the sampler (materialised grid points) gives the same density
    pts = model.sampler.get_grid_points()
    assert torch.allclose(model.prob(pts).cpu(), grid, rtol=1e-4, atol=2e-6 * float(refg.max()))
    # a single Lagrange function called like the reference's interpolator
    u = torch.linspace(-4.5, 4.5, 1001)
    h = model.lagrange_functions[1][0](u.cuda()).cpu()
    want = hp.lagrange_interp(t32(g["tables0"])[1], hp.centres(t32(g["edges"])), u).float()
    assert torch.allclose(h, want, rtol=1e-6, atol=1e-7)


def test_grid_sampler_statistics(golden):
    g = golden("ment_4d")
    res, xmax = int(g["grid_res"]), float(g["grid_xmax"])
    rho = cuda(g["prob_grid"])
    n = 4_000_000
    cell = 2 * xmax / res
    x = ops.cdf_sample(rho, [res] * 4, [-xmax] * 4, [cell] * 4, n, seed=7)
    assert x.shape == (n, 4) and float(x.min()) >= -xmax and float(x.max()) <= xmax
    idx = torch.floor((x.double() + xmax) / cell).long().clamp_(0, res - 1)
    flat = ((idx[:, 0] * res + idx[:, 1]) * res + idx[:, 2]) * res + idx[:, 3]
    freq = torch.bincount(flat, minlength=res ** 4).double().cpu()
    pmf = hp.cell_pmf(t32(g["prob_grid"]).double())
    big = pmf * n > 50
    zscore = (freq[big] - pmf[big] * n) / torch.sqrt(pmf[big] * n)
    assert float(zscore.abs().max()) < 6.5 and abs(float(zscore.mean())) < 0.1 and abs(float(zscore.std()) - 1) < 0.1
    assert float(freq[pmf * n < 1e-6].sum()) == 0
    frac = ((x.double() + xmax) / cell) % 1.0
    assert (frac.mean(dim=0).cpu() - 0.5).abs().max() < 2e-3            # uniform inside the cell
    again = ops.cdf_sample(rho, [res] * 4, [-xmax] * 4, [cell] * 4, n, seed=7)
    assert torch.equal(x, again)                                        # Philox stream: reproducible
    other = ops.cdf_sample(rho, [res] * 4, [-xmax] * 4, [cell] * 4, n, seed=8)
    assert not torch.equal(x, other)
    # same moments as the reference's sampler (golden drew 30000 particles)
    assert (x.double().mean(dim=0).cpu() - torch.from_numpy(g["xs_mean"])).abs().max() < 0.05
    assert (torch.cov(x.double().T).cpu() - torch.from_numpy(g["xs_cov"])).abs().max() < 0.08


def test_sample_mode_update_is_statistically_consistent(golden):
    g = golden("ment_4d")
    torch.manual_seed(0)
    model = _model(g, n_samples=2_000_000)
    for i, t in enumerate(t32(g["tables0"])):
        model.lagrange_functions[i][0].set_values(t.cuda())
    pred = model.simulate(0, 0).cpu()
    ref = t32(g["pred0"])                      # 30000-particle estimate from the reference
    assert (pred - ref).abs().max() < 0.06 * ref.max()
    width = float(t32(g["edges"])[1] - t32(g["edges"])[0])
    assert abs(float(pred.sum()) * width - 1.0) < 1e-5
    # one full Gauss-Seidel sweep at high statistics on both sides: the oracle (pinned to the
    # reference's sweep by tests/test_oracle_golden.py) with torch's CPU generator, the CUDA path with
    # its Philox stream.  Sampling noise ~ sqrt(32 bins / 1e6) = 0.6 %, compounded over 6 updates.
    n = 1_000_000
    model.n_samples = n
    model.gauss_seidel_update(lr=float(g["lr"]), thresh=float(g["thresh"]))
    got = torch.stack([model.lagrange_functions[i][0].values.cpu() for i in range(6)])
    mats, edges, meas = t32(g["matrices"]), t32(g["edges"]), t32(g["meas"])
    res, xmax, s = int(g["grid_res"]), float(g["grid_xmax"]), float(g["prior_scale"])
    scr = [[hp.Screen1D(edges=edges)] for _ in mats]
    gedges = [torch.linspace(-xmax, xmax, res + 1) for _ in range(4)]
    pts = hp.grid_points([hp.centres(e) for e in gedges])
    cur = [t.clone() for t in t32(g["tables0"])]
    gen = torch.Generator().manual_seed(5)
    preds = []
    for i in range(6):
        pgi = hp.ment_prob(pts, list(mats), scr, [[t] for t in cur], s)
        xs = hp.sample_grid(pgi.reshape(4 * [res]), gedges, n, generator=gen)
        pr = hp.kde_profile_1d(hp.linear_map(xs, mats[i])[:, 0], edges, scr[i][0].sigma, chunk=100000)
        pr = hp.normalize_projection(pr, edges[1] - edges[0])
        preds.append(pr)
        cur[i] = hp.gauss_seidel_table(cur[i], meas[i], pr, float(g["lr"]), float(g["thresh"]))
    want = torch.stack(cur)
    preds = torch.stack(preds)
    solid = (meas > 0) & (preds > 0.02 * preds.max(dim=1, keepdim=True).values)
    rel = ((got - want).abs() / want.abs().clamp_min(1e-12))[solid]
    assert float(rel.median()) < 0.02 and float(rel.max()) < 0.25
    assert torch.equal(got == 0, want == 0)
    assert model.epoch == 1


def test_integrate_mode_matches_reference(golden):
    g = golden("ment_2d_integrate")
    model = _model(g, mode="integrate")
    pred = model.simulate(1, 0).cpu()
    assert torch.allclose(pred, t32(g["pred_1_0"]), rtol=1e-4, atol=1e-7)
    model.gauss_seidel_update(lr=1.0, thresh=1.0e-10)
    model.gauss_seidel_step(lr=1.0, thresh=1.0e-10)
    got = torch.stack([model.lagrange_functions[i][0].values.cpu() for i in range(5)])
    assert torch.allclose(got, t32(g["tables_after_2"]), rtol=2e-4, atol=1e-6)


def test_gs_update_kernel_matches_rule():
    torch.manual_seed(1)
    table, meas, pred = torch.rand(64) + 0.1, torch.rand(64), torch.rand(64)
    meas[:7] = 0.0
    pred[5:12] = 1e-12
    want = hp.gauss_seidel_table(table, meas, pred, 0.9, 1e-10)
    t = table.clone().cuda()
    ops.gs_update(t, meas.cuda(), pred.cuda(), 0.9, 1e-10)
    assert torch.allclose(t.cpu(), want, rtol=1e-6)
