"""Pins oracle/hotpath.py to outputs of the REAL reference code (tests/golden/*.npz made by
oracle/make_goldens.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import hotpath as hp
from mfb_testutil import t32


def _screens1d(edges, k, bandwidth=0.5, direction=None):
    return [[hp.Screen1D(edges=edges, bandwidth=bandwidth, axis=0, direction=direction)] for _ in range(k)]


def test_kde1d_6d_matches_reference(golden):
    g = golden("kde1d_6d")
    x, mats, edges = t32(g["x"]), t32(g["matrices"]), t32(g["edges"])
    out = hp.simulate(x, list(mats), _screens1d(edges, len(mats)))
    got = torch.stack([o[0] for o in out])
    assert torch.allclose(got, t32(g["kde"]), rtol=0, atol=2e-7)
    # chunked evaluation = same arithmetic per row
    out = hp.simulate(x, list(mats), _screens1d(edges, len(mats)), chunk=512)
    got = torch.stack([o[0] for o in out])
    assert torch.allclose(got, t32(g["kde"]), rtol=1e-5, atol=1e-7)


def test_hard_hist_1d_matches_reference_bitwise(golden):
    g = golden("kde1d_6d")
    x, mats, edges = t32(g["x"]), t32(g["matrices"]), t32(g["edges"])
    out = hp.simulate(x, list(mats), _screens1d(edges, len(mats)), kde=False)
    got = torch.stack([o[0] for o in out])
    assert torch.equal(got, t32(g["hard"]))
    # integer counts -> density reproduces torch.histogram(density=True) bit for bit
    for k in range(len(mats)):
        u = t32(g["uproj"][k])
        counts = hp.hist_counts_1d(u, edges)
        assert torch.equal(hp.density_from_counts_1d(counts, edges), t32(g["hard"][k]))


def test_kde1d_2d_rotations_and_direction(golden):
    g = golden("kde1d_2d")
    x, mats, edges = t32(g["x"]), t32(g["matrices"]), t32(g["edges"])
    got = torch.stack([o[0] for o in hp.simulate(x, list(mats), _screens1d(edges, len(mats)))])
    assert torch.allclose(got, t32(g["kde"]), rtol=0, atol=2e-7)
    scr = _screens1d(edges, len(mats), bandwidth=float(g["bandwidth_dir"]), direction=t32(g["direction"]))
    got = torch.stack([o[0] for o in hp.simulate(x, list(mats), scr)])
    assert torch.allclose(got, t32(g["kde_dir"]), rtol=0, atol=2e-7)


def test_kde2d_and_hard2d(golden):
    g = golden("kde2d_4d")
    x, mats = t32(g["x"]), t32(g["matrices"])
    ex, ey = t32(g["edges_x"]), t32(g["edges_y"])
    bw = tuple(float(b) for b in g["bandwidth"])
    scr = [[hp.Screen2D(axis=(0, 2), edges_x=ex, edges_y=ey, bandwidth=bw)] for _ in mats]
    got = torch.stack([o[0] for o in hp.simulate(x, list(mats), scr)])
    assert torch.allclose(got, t32(g["kde"]), rtol=1e-6, atol=1e-8)
    hard = torch.stack([o[0] for o in hp.simulate(x, list(mats), scr, kde=False)])
    assert torch.equal(hard, t32(g["hard"]))
    # counts rule == np.histogramdd
    u = hp.linear_map(x, mats[0])[:, [0, 2]]
    counts = hp.hist_counts_2d(u, ex, ey)
    dens = counts.double() / counts.sum() / (torch.diff(ex).double()[:, None] * torch.diff(ey).double()[None, :])
    assert torch.allclose(dens.float(), t32(g["hard"][0]), rtol=1e-6, atol=0)
    kl = torch.stack([hp.kl_div(p, m) for p, m in zip(got, t32(g["meas"]))])
    assert torch.allclose(kl, t32(g["kl"]), rtol=1e-5, atol=1e-8)


def test_discrepancies_and_gradient(golden):
    g = golden("kde1d_6d")
    kde, meas = t32(g["kde"]), t32(g["meas"])
    assert torch.allclose(torch.stack([hp.kl_div(p, m) for p, m in zip(kde, meas)]), t32(g["kl"]), rtol=1e-6)
    assert torch.allclose(torch.stack([hp.mae(p, m) for p, m in zip(kde, meas)]), t32(g["mae"]), rtol=1e-6)
    assert torch.allclose(torch.stack([hp.mse(p, m) for p, m in zip(kde, meas)]), t32(g["mse"]), rtol=1e-6)
    x = t32(g["x"]).requires_grad_(True)
    mats, edges = t32(g["matrices"]), t32(g["edges"])
    out = hp.simulate(x, list(mats), _screens1d(edges, len(mats)))
    loss = sum(hp.kl_div(o[0], m) for o, m in zip(out, meas)) / len(mats)
    loss.backward()
    assert abs(float(loss.detach()) - float(g["mean_kl"])) <= 1e-6 * abs(float(g["mean_kl"]))
    ref = t32(g["grad_x"])
    assert (x.grad - ref).abs().max() <= 1e-5 * ref.abs().max()


def test_prior_entropy_loss(golden):
    g = golden("entropy_loss")
    x, logq = t32(g["x"]), t32(g["logq"])
    s = float(g["prior_scale"])
    assert torch.allclose(hp.gaussian_log_prob(x, s), t32(g["prior_log_prob"]), rtol=1e-6, atol=1e-6)
    assert abs(float(hp.mc_entropy(x, logq, s)) - float(g["h_mc"])) < 1e-5
    assert abs(float(hp.mc_entropy(x, logq, None)) - float(g["h_mc_noprior"])) < 1e-5
    assert abs(float(hp.cov_entropy(x)) - float(g["h_cov"])) < 1e-5
    k = golden("kde1d_6d")
    xs, mats, edges, meas = t32(k["x"]), t32(k["matrices"]), t32(k["edges"]), t32(k["meas"])
    n = int(g["loss_n"])
    L, H, D = hp.mentflow_loss(xs[:n], t32(g["loss_logq"])[:n], list(mats), _screens1d(edges, len(mats)),
                               [[m] for m in meas], s, float(g["loss_penalty"]))
    assert abs(float(L) - float(g["loss_L"])) <= 1e-5 * abs(float(g["loss_L"]))
    assert abs(float(H) - float(g["loss_H"])) <= 1e-5 * abs(float(g["loss_H"]))
    assert torch.allclose(torch.stack(D), t32(g["loss_D"]), rtol=1e-5)


def test_ment_prob_sampling_and_update(golden):
    g = golden("ment_4d")
    mats, edges, meas = t32(g["matrices"]), t32(g["edges"]), t32(g["meas"])
    k = len(mats)
    scr = _screens1d(edges, k)
    tables = [[t] for t in t32(g["tables0"])]
    s = float(g["prior_scale"])
    prob = hp.ment_prob(t32(g["xq"]), list(mats), scr, tables, s)
    assert torch.allclose(prob, t32(g["prob_q"]), rtol=1e-5, atol=1e-30)
    res, xmax = int(g["grid_res"]), float(g["grid_xmax"])
    gedges = [torch.linspace(-xmax, xmax, res + 1) for _ in range(4)]
    pts = hp.grid_points([hp.centres(e) for e in gedges])
    pg = hp.ment_prob(pts, list(mats), scr, tables, s)
    assert torch.allclose(pg, t32(g["prob_grid"]), rtol=1e-5, atol=1e-30)
    # same global RNG stream as the reference's sampler (sample.py:27-57)
    torch.manual_seed(int(g["sample_seed"]))
    xs = hp.sample_grid(pg.reshape(4 * [res]), gedges, int(g["n_samples"]))
    assert torch.equal(xs[:512], t32(g["xs_head"]))
    # one simulate(0, 0) in sample mode: sample -> transform -> KDE -> normalise
    torch.manual_seed(int(g["sample_seed"]))
    xs = hp.sample_grid(pg.reshape(4 * [res]), gedges, int(g["n_samples"]))
    pred = hp.kde_profile_1d(hp.linear_map(xs, mats[0])[:, 0], edges, scr[0][0].sigma)
    pred = hp.normalize_projection(pred, edges[1] - edges[0])
    assert torch.allclose(pred, t32(g["pred0"]), rtol=1e-5, atol=1e-8)
    # a whole Gauss-Seidel sweep (ment.py:336-371): sequential over k, tables feed forward
    torch.manual_seed(int(g["gs_seed"]))
    cur = [t.clone() for t in t32(g["tables0"])]
    for i in range(k):
        pgi = hp.ment_prob(pts, list(mats), scr, [[t] for t in cur], s)
        xs = hp.sample_grid(pgi.reshape(4 * [res]), gedges, int(g["n_samples"]))
        pred = hp.kde_profile_1d(hp.linear_map(xs, mats[i])[:, 0], edges, scr[i][0].sigma)
        pred = hp.normalize_projection(pred, edges[1] - edges[0])
        cur[i] = hp.gauss_seidel_table(cur[i], meas[i], pred, float(g["lr"]), float(g["thresh"]))
    assert torch.allclose(torch.stack(cur), t32(g["tables1"]), rtol=1e-5, atol=1e-8)


def test_projection_order_reproduces_reference_bits(golden):
    """The CUDA kernels project with an ascending fused-multiply-add chain starting from zero
    (kde1d.cu project_row).  Restated here in numpy (float64 product + sum rounded once = fmaf): it reproduces the
    reference's `x @ M.T` bit for bit on its own stored projections, which is what makes fused exact histograms
    through general matrices bit-exact."""
    g = golden("kde1d_6d")
    x, mats, uproj = g["x"].astype(np.float32), g["matrices"].astype(np.float32), g["uproj"].astype(np.float32)
    for k in range(mats.shape[0]):
        acc = np.zeros(x.shape[0], np.float32)
        for i in range(x.shape[1]):
            acc = (x[:, i].astype(np.float64) * np.float64(mats[k, 0, i]) + acc.astype(np.float64)).astype(np.float32)
        assert np.array_equal(acc.view(np.uint32), uproj[k].view(np.uint32))


def test_ment_prob_with_2d_screens(golden):
    """N-D Lagrange tables (ment.py:20-52) on the reference's rec_nd_2d_ment set-up, plus a mixed 1-D / 2-D model
    and the integration-mode prediction of a 2-D screen (ment.py:267-317)."""
    g = golden("ment_2d_screens")
    mats, ex, ey = t32(g["matrices"]), t32(g["edges_x"]), t32(g["edges_y"])
    scr = [[hp.Screen2D(axis=(0, 2), edges_x=ex, edges_y=ey)] for _ in mats]
    tables = [[t] for t in t32(g["tables0"])]
    s = float(g["prior_scale"])
    xq = t32(g["xq"])
    assert torch.allclose(hp.ment_prob(xq, list(mats), scr, tables, s), t32(g["prob_q"]), rtol=1e-5, atol=1e-30)
    res, xmax = int(g["grid_res"]), float(g["grid_xmax"])
    pts = hp.grid_points([hp.centres(torch.linspace(-xmax, xmax, res + 1)) for _ in range(4)])
    assert torch.allclose(hp.ment_prob(pts, list(mats), scr, tables, s), t32(g["prob_grid"]), rtol=1e-5, atol=1e-30)
    mats1, e1 = t32(g["matrices1"]), t32(g["edges1"])
    scr_m = scr + _screens1d(e1, len(mats1))
    tab_m = tables + [[t] for t in t32(g["tables1d"])]
    got = hp.ment_prob(xq, list(mats) + list(mats1), scr_m, tab_m, s)
    assert torch.allclose(got, t32(g["prob_q_mixed"]), rtol=1e-5, atol=1e-30)
    # integration mode, screen 2: u = [pixel on axes (0, 2); grid on axes (1, 3)], x = M^-1 u, sum of rho
    lim, shape = g["int_limits"], [int(v) for v in g["int_shape"]]
    c1 = torch.linspace(float(lim[0][0]), float(lim[0][1]), shape[0])
    c3 = torch.linspace(float(lim[1][0]), float(lim[1][1]), shape[1])
    grid = hp.grid_points([c1, c3])
    cx, cy = hp.centres(ex), hp.centres(ey)
    minv = torch.linalg.inv(mats[2])
    pred = torch.zeros(cx.numel(), cy.numel())
    for i, px in enumerate(cx):
        for j, py in enumerate(cy):
            u = torch.zeros(grid.shape[0], 4)
            u[:, 0], u[:, 2], u[:, 1], u[:, 3] = px, py, grid[:, 0], grid[:, 1]
            pred[i, j] = hp.ment_prob(hp.linear_map(u, minv), list(mats), scr, tables, s).sum()
    pred = hp.normalize_projection(pred, (ex[1] - ex[0]) * (ey[1] - ey[0]))
    ref = t32(g["pred_2_0"])
    assert torch.allclose(pred, ref, rtol=1e-4, atol=1e-7 * float(ref.max()))


def test_ment_integrate_mode(golden):
    g = golden("ment_2d_integrate")
    mats, edges, meas = t32(g["matrices"]), t32(g["edges"]), t32(g["meas"])
    k, s = len(mats), float(g["prior_scale"])
    scr = _screens1d(edges, k)
    lo, hi = [float(v) for v in g["int_limits"]]
    ipts = torch.linspace(lo, hi, int(g["int_shape"]))
    c = hp.centres(edges)

    def integrate(i, tables):
        """ment.py:267-317 for a 1-D screen in 2-D: u = (pixel, t); x = M^-1 u; sum rho."""
        minv = torch.linalg.inv(mats[i])
        pred = torch.zeros(len(c))
        for b in range(len(c)):
            u = torch.stack([torch.full_like(ipts, float(c[b])), ipts], dim=1)
            pred[b] = hp.ment_prob(torch.matmul(u, minv.T), list(mats), scr, tables, s).sum()
        return hp.normalize_projection(pred, edges[1] - edges[0])

    tables = [[hp.initial_table(m)] for m in meas]
    assert torch.allclose(integrate(1, tables), t32(g["pred_1_0"]), rtol=1e-5, atol=1e-8)
    for _ in range(2):
        for i in range(k):
            tables[i][0] = hp.gauss_seidel_table(tables[i][0], meas[i], integrate(i, tables), 1.0, 1e-10)
    got = torch.stack([t[0] for t in tables])
    assert torch.allclose(got, t32(g["tables_after_2"]), rtol=1e-4, atol=1e-7)


def test_multipole_kick_and_projection_transform_match_reference(golden):
    """oracle.hotpath.multipole_kick / projection_transform and the package's transform classes against
    the reference's MultipoleTransform / ProjectionTransform outputs (simulate/transform.py:78-156)."""
    import mentflow_b200 as mf
    g = golden("multipole")
    for d, order, strength, skew in g["kick_cases"]:
        key = f"d{int(d)}_o{int(order)}_{'s' if skew else 'n'}"
        x, u = t32(g["kick_x_" + key]), t32(g["kick_u_" + key])
        assert torch.equal(hp.multipole_kick(x, int(order), float(strength), bool(skew)), u)
        assert torch.equal(mf.simulate.MultipoleTransform(int(order), float(strength), bool(skew))(x), u)
    for bad in (1, 2, 6):
        with pytest.raises(ValueError):
            hp.multipole_kick(t32(g["kick_x_d2_o3_n"]), bad, 1.0)
        with pytest.raises(ValueError):
            mf.simulate.MultipoleTransform(bad, 1.0)(t32(g["kick_x_d2_o3_n"]))
    x4, direction = t32(g["pt_x"]), t32(g["pt_direction"])
    assert torch.equal(hp.projection_transform(x4, direction), t32(g["pt_u"]))
    assert torch.equal(mf.simulate.ProjectionTransform(direction)(x4), t32(g["pt_u"]))
    # inverse of the kick: momentum reversal, kick, momentum reversal; the argument is not modified
    kick = mf.simulate.MultipoleTransform(3, 0.7)
    x = t32(g["kick_x_d4_o3_n"])
    keep = x.clone()
    back = kick.inverse(kick(x))
    assert torch.equal(x, keep)
    assert torch.allclose(back[:, [0, 2]], x[:, [0, 2]])


def test_multipole_terms_reproduce_the_composite_map(golden):
    """Host-side folding of linear -> kick -> linear into (w, wa, wb, a, b): the measured coordinate of the
    closed form equals row . CompositeTransform(x) of the reference semantics, all orders, skew or not."""
    import mentflow_b200 as mf
    from mentflow_b200.simulate.simulate import _multipole_chain, multipole_terms
    gen = torch.Generator().manual_seed(4)
    for d in (2, 4, 6):
        x = torch.randn(300, d, generator=gen, dtype=torch.float64) * 0.7
        for order in (3, 4, 5):
            for skew in (False, True):
                pre = torch.randn(d, d, generator=gen, dtype=torch.float64)
                post = torch.randn(d, d, generator=gen, dtype=torch.float64)
                chain = mf.simulate.CompositeTransform(mf.simulate.LinearTransform(pre),
                                                       mf.simulate.MultipoleTransform(order, 0.6, skew),
                                                       mf.simulate.LinearTransform(post))
                got = _multipole_chain(chain)
                assert got is not None
                for axis in range(d):
                    w, mp = multipole_terms(post[axis], got[0], got[1], d)
                    w, mp = w.double(), mp.double()
                    wa, wb, a, b, m = mp[:d], mp[d:2 * d], mp[2 * d], mp[2 * d + 1], int(mp[2 * d + 2]) - 1
                    z = torch.complex(x @ wa, x @ wb) ** m
                    u = x @ w + a * z.real + b * z.imag
                    ref = chain(x)[:, axis]
                    assert float((u - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    # a chain with two kicks, or an unsupported order, is not folded
    two = mf.simulate.CompositeTransform(mf.simulate.MultipoleTransform(3, 0.1), mf.simulate.MultipoleTransform(3, 0.1))
    assert _multipole_chain(two) is None
    assert _multipole_chain(mf.simulate.MultipoleTransform(2, 0.1)) is None


def test_measurement_noise_matches_reference(golden):
    """Histogram.forward's noise (diagnostics/diagnostics.py:50-68): the oracle's apply_noise and the package's
    Histogram.apply_noise applied to the reference's clean profile reproduce the reference's noisy one --
    same generator stream (re-seeded per call), gaussian and uniform, 1-D and 2-D screens, clamp at zero."""
    import mentflow_b200 as mf
    g = golden("noise")
    edges = t32(g["edges"])
    for kind in ("gaussian", "uniform"):
        for seed in (3, 11):
            for tag, clean in (("h1", t32(g["h1_clean"])), ("h2", t32(g["h2_clean"]))):
                ref = t32(g[f"{tag}_{kind}_{seed}"])
                assert float(ref.min()) >= 0.0 and not torch.equal(ref, clean)
                got = hp.apply_noise(clean, 0.4, kind, seed)
                assert torch.allclose(got, ref, rtol=1e-6, atol=1e-9)
                if tag == "h1":
                    diag = mf.diagnostics.Histogram1D(axis=0, edges=edges, bandwidth=0.5, noise=True, noise_scale=0.4,
                                                      noise_type=kind, seed=seed)
                else:
                    diag = mf.diagnostics.Histogram2D(axis=(0, 2), edges=(edges, edges), bandwidth=(0.5, 0.5),
                                                      noise=True, noise_scale=0.4, noise_type=kind, seed=seed)
                assert torch.allclose(diag.apply_noise(clean), ref, rtol=1e-6, atol=1e-9)
                assert torch.equal(diag.apply_noise(clean), diag.apply_noise(clean))     # re-seeded on every call
                diag.set_noise(False)
                assert torch.equal(diag.apply_noise(clean), clean)
