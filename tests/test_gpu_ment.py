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


@pytest.mark.parametrize("shape", [(3,), (5,), (37,), (1000,), (7, 13), (65, 63), (4, 4, 4, 5), (17, 16, 16)])
def test_cdf_sample_on_ragged_grid_sizes(shape):
    """The pivot-level search of cdf_sample (four keys per sector, partial groups at every level) picks cells with the
    probabilities of the density for cell counts that are not powers of four, never a zero-probability cell, and the
    last cell is reachable."""
    torch.manual_seed(sum(shape))
    g = 1
    for s_ in shape:
        g *= s_
    rho = torch.rand(g) ** 3
    rho[torch.rand(g) < 0.3] = 0.0
    rho[-1] = rho.max() * 2            # the last cell carries weight: the descent must reach it
    rho[0] = 0.0
    n = 2_000_000
    d = len(shape)
    x = ops.cdf_sample(rho.cuda(), list(shape), [0.0] * d, [1.0] * d, n, seed=3)
    idx = torch.floor(x.double()).long()
    flat = torch.zeros(n, dtype=torch.long, device=x.device)
    for i, s_ in enumerate(shape):
        flat = flat * s_ + idx[:, i].clamp_(0, s_ - 1)
    freq = torch.bincount(flat, minlength=g).double().cpu()
    pmf = (rho.double() / rho.double().sum())
    assert float(freq[pmf == 0].sum()) == 0
    big = pmf * n > 50
    z = (freq[big] - pmf[big] * n) / torch.sqrt(pmf[big] * n * (1 - pmf[big]))
    assert float(z.abs().max()) < 6.0, float(z.abs().max())
    assert freq[-1] > 0


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


# --------------------------------------------------------------------------------------------------
# two-dimensional screens (ment.py:20-52 N-D tables; experiments/config/rec_nd_2d_ment.yaml)
# --------------------------------------------------------------------------------------------------
def _model_2d(g, mode="sample", with_1d=False, n_samples=1000):
    mats, ex, ey, meas = t32(g["matrices"]), t32(g["edges_x"]), t32(g["edges_y"]), t32(g["meas"])
    d = mats.shape[1]
    tfs = [mf.simulate.LinearTransform(m.cuda()) for m in mats]
    diag = mf.diagnostics.Histogram2D(axis=(0, 2), edges=(ex, ey), bandwidth=(0.5, 0.5)).to("cuda")
    diags = [[diag] for _ in tfs]
    ms = [[m.cuda()] for m in meas]
    if with_1d:
        mats1, e1, meas1 = t32(g["matrices1"]), t32(g["edges1"]), t32(g["meas1"])
        diag1 = mf.diagnostics.Histogram1D(axis=0, edges=e1, bandwidth=0.5).to("cuda")
        tfs += [mf.simulate.LinearTransform(m.cuda()) for m in mats1]
        diags += [[diag1] for _ in mats1]
        ms += [[m.cuda()] for m in meas1]
    kw = {}
    xmax, res = float(g["grid_xmax"]), int(g["grid_res"])
    if mode == "sample":
        kw["sampler"] = mf.sample.GridSampler(limits=d * [(-xmax, xmax)], shape=tuple(d * [res]), device="cuda")
        kw["n_samples"] = n_samples
    else:
        lim = [tuple(float(v) for v in row) for row in g["int_limits"]]
        kw["integration_limits"] = [[lim] for _ in tfs]
        kw["integration_shape"] = [[tuple(int(v) for v in g["int_shape"])] for _ in tfs]
    model = mf.ment.MENT(ndim=d, transforms=tfs, diagnostics=diags, measurements=ms,
                         prior=mf.prior.Gaussian(ndim=d, scale=float(g["prior_scale"])), mode=mode, device="cuda", **kw)
    for i, t in enumerate(t32(g["tables0"])):
        model.lagrange_functions[i][0].set_values(t.cuda())
    return model


def test_prob_with_2d_screens_matches_reference(golden):
    g = golden("ment_2d_screens")
    model = _model_2d(g)
    assert model.lagrange_functions[0][0].values.shape == (16, 14)
    prob = model.prob(cuda(g["xq"])).cpu()
    ref = t32(g["prob_q"])
    assert torch.allclose(prob, ref, rtol=1e-4, atol=1e-9 * float(ref.max()))
    assert int((ref > 0).sum()) > 500                     # the comparison is not vacuous
    grid = model.prob_on_grid(model.sampler).cpu()
    refg = t32(g["prob_grid"])
    assert torch.allclose(grid, refg, rtol=1e-4, atol=2e-6 * float(refg.max()))
    # one 2-D Lagrange function called like the reference's interpolator: bilinear, zero outside the centres
    lf = model.lagrange_functions[3][0]
    cx, cy = hp.centres(t32(g["edges_x"])), hp.centres(t32(g["edges_y"]))
    torch.manual_seed(2)
    u = torch.rand(5000, 2) * torch.tensor([9.0, 8.0]) - torch.tensor([4.5, 4.0])
    got = lf(u.cuda()).cpu()
    tab = t32(g["tables0"])[3].double()
    ix = (torch.searchsorted(cx.double(), u[:, 0].double()) - 1).clamp(0, cx.numel() - 2)
    iy = (torch.searchsorted(cy.double(), u[:, 1].double()) - 1).clamp(0, cy.numel() - 2)
    wx = (u[:, 0].double() - cx.double()[ix]) / (cx.double()[ix + 1] - cx.double()[ix])
    wy = (u[:, 1].double() - cy.double()[iy]) / (cy.double()[iy + 1] - cy.double()[iy])
    want = (tab[ix, iy] * (1 - wx) * (1 - wy) + tab[ix, iy + 1] * (1 - wx) * wy + tab[ix + 1, iy] * wx * (1 - wy)
            + tab[ix + 1, iy + 1] * wx * wy)
    inside = (u[:, 0] >= cx[0]) & (u[:, 0] <= cx[-1]) & (u[:, 1] >= cy[0]) & (u[:, 1] <= cy[-1])
    want = torch.where(inside, want, torch.zeros_like(want)).float()
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)


def test_prob_with_mixed_1d_and_2d_screens_matches_reference(golden):
    g = golden("ment_2d_screens")
    model = _model_2d(g, with_1d=True)
    for i, t in enumerate(t32(g["tables1d"])):
        model.lagrange_functions[6 + i][0].set_values(t.cuda())
    prob = model.prob(cuda(g["xq"])).cpu()
    ref = t32(g["prob_q_mixed"])
    assert torch.allclose(prob, ref, rtol=1e-4, atol=1e-9 * float(ref.max()))


def test_integrate_mode_with_2d_screens_matches_reference(golden):
    g = golden("ment_2d_screens")
    model = _model_2d(g, mode="integrate")
    pred = model.simulate(2, 0).cpu()
    ref = t32(g["pred_2_0"])
    assert pred.shape == ref.shape == (16, 14)
    assert torch.allclose(pred, ref, rtol=1e-4, atol=1e-7 * float(ref.max()))
    model.gauss_seidel_update(lr=float(g["lr"]), thresh=1.0e-10)
    got = torch.stack([model.lagrange_functions[i][0].values.cpu() for i in range(6)])
    want = t32(g["tables_after_gs"])
    assert torch.allclose(got, want, rtol=5e-4, atol=1e-6)
    assert torch.equal(got == 0, want == 0)


def test_sample_mode_sweep_with_2d_screens_runs_and_is_consistent(golden):
    """rec_nd_2d_ment.yaml runs in sample mode: a sweep at high statistics must agree with the integration-mode
    sweep of the same model on a fine grid (both estimate the same projections)."""
    g = golden("ment_2d_screens")
    torch.manual_seed(3)
    model = _model_2d(g, n_samples=4_000_000)
    model.sampler = mf.sample.GridSampler(limits=4 * [(-4.0, 4.0)], shape=(24, 24, 24, 24), device="cuda")
    pred_s = model.simulate(2, 0).cpu()
    fine = _model_2d(g, mode="integrate")
    fine.integration_limits = [[[(-4.0, 4.0), (-4.0, 4.0)]] for _ in fine.transforms]
    fine.integration_shape = [[(160, 160)] for _ in fine.transforms]
    pred_i = fine.simulate(2, 0).cpu()
    cell = float((t32(g["edges_x"])[1] - t32(g["edges_x"])[0]) * (t32(g["edges_y"])[1] - t32(g["edges_y"])[0]))
    assert abs(float(pred_s.sum()) * cell - 1.0) < 1e-4
    # the sampled profile is the integrated one smoothed by the KDE kernel and the grid cells: compare totals of
    # coarse blocks rather than pixels
    bs, bi = pred_s[:16, :12].reshape(4, 4, 3, 4).sum(dim=(1, 3)), pred_i[:16, :12].reshape(4, 4, 3, 4).sum(dim=(1, 3))
    assert float((bs - bi).abs().max()) < 0.12 * float(bi.max())
    model.gauss_seidel_update(lr=0.5)
    assert model.epoch == 1 and all(torch.isfinite(model.lagrange_functions[i][0].values).all() for i in range(6))


def test_sharded_draws_are_slices_of_one_stream(golden):
    """What rank r of W draws (offset = first index of its slice) is the slice of the single-GPU draw."""
    g = golden("ment_4d")
    model = _model(g, n_samples=100_003)
    for i, t in enumerate(t32(g["tables0"])):
        model.lagrange_functions[i][0].set_values(t.cuda())
    torch.manual_seed(17)
    full = model.sample(100_003)
    parts = []
    for r in range(3):
        model.shard = (r, 3)
        torch.manual_seed(17)
        parts.append(model.sample(100_003))
    model.shard = None
    assert [p.shape[0] for p in parts] == [33_335, 33_334, 33_334]
    assert torch.equal(torch.cat(parts), full)
