"""TEST INFRASTRUCTURE -- generate ``tests/golden/*.npz`` from the REAL reference package.

Run in the build container (where ``/root/reference`` exists):

    python -m oracle.make_goldens

Each fixture stores the seeded inputs and the outputs the reference's own code produced for
them (reference functions named in each block).  The fixtures pin ``oracle/hotpath.py``
(``tests/test_oracle_golden.py``, CPU) and the CUDA kernels (``tests/test_gpu_*.py``).
The zuko flow cannot run here (zuko absent) so the NSF fixture comes from the restatement in
``oracle/zuko_nsf.py`` evaluated in float64 -- it guards against drift, it does not pin zuko.
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from oracle.zuko_nsf import NSFOracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def isotropic_matrices(k, d, seed):
    """experiments/rec_nd_1d/setup.py:28-49 (values only)."""
    rng = torch.Generator().manual_seed(seed)
    dirs = torch.randn((k, d), generator=rng)
    dirs = dirs / torch.norm(dirs, dim=1)[:, None]
    mats = []
    for v in dirs:
        m = torch.eye(d)
        m[0, :] = v
        mats.append(m.float())
    return mats


def rotation_matrices(k):
    """experiments/rec_2d/linear/setup.py:29-43 + simulate/transform.py:12-15."""
    mats = []
    for a in np.linspace(0.0, np.pi, k, endpoint=False):
        c, s = np.cos(a), np.sin(a)
        mats.append(torch.tensor([[c, s], [-s, c]]).type(torch.float32))
    return mats


def corner_matrices(d):
    """experiments/rec_nd_2d/setup.py:38-53."""
    mats = []
    for i in range(d):
        for j in range(i):
            parts = []
            for k, l in zip((0, 2), (j, i)):
                m = torch.eye(d)
                m[k, k] = m[l, l] = 0.0
                m[k, l] = m[l, k] = 1.0
                parts.append(m.float())
            mats.append(torch.linalg.multi_dot(parts[::-1]))
    return mats


def main():
    mf = ref_import.load()
    os.makedirs(OUT, exist_ok=True)

    # ---------------------------------------------------------------- KDE-1D, 6D isotropic
    torch.manual_seed(11)
    d, k, nb, n = 6, 9, 64, 3000
    x = (torch.randn(n, d) * torch.tensor([1.0, 0.6, 1.4, 0.8, 1.1, 0.9])).float()
    x[:5] *= 4.0                      # a few particles outside the screen
    mats = isotropic_matrices(k, d, seed=0)
    edges = torch.linspace(-3.5, 3.5, nb + 1)
    diag = mf.diagnostics.Histogram1D(axis=0, edges=edges, bandwidth=0.5)
    tfs = [mf.simulate.LinearTransform(m) for m in mats]
    preds = mf.simulate.forward(x, tfs, [[diag] for _ in tfs])
    kde = torch.stack([p[0] for p in preds])
    diag.kde = False
    hard = torch.stack([p[0] for p in mf.simulate.forward(x, tfs, [[diag] for _ in tfs])])
    diag.kde = True
    uproj = torch.stack([diag.project(t(x)) for t in tfs])
    # measurement = hard histogram of a different sample, renormalised (experiments/setup.py:63-73)
    xm = torch.randn(20000, d).float() * 0.9
    diag.kde = False
    meas = torch.stack([p[0] for p in mf.simulate.forward(xm, tfs, [[diag] for _ in tfs])])
    diag.kde = True
    meas = meas / meas.sum(dim=1, keepdim=True) / (edges[1] - edges[0])
    kl = torch.stack([mf.loss.kl_divergence(p, m) for p, m in zip(kde, meas)])
    ma = torch.stack([mf.loss.mean_absolute_error(p, m) for p, m in zip(kde, meas)])
    ms = torch.stack([mf.loss.mean_square_error(p, m) for p, m in zip(kde, meas)])
    # gradient of mean KL w.r.t. particles through the reference's autograd
    xg = x.clone().requires_grad_(True)
    pg = mf.simulate.forward(xg, tfs, [[diag] for _ in tfs])
    loss = sum(mf.loss.kl_divergence(p[0], m) for p, m in zip(pg, meas)) / k
    loss.backward()
    np.savez_compressed(os.path.join(OUT, "kde1d_6d.npz"), x=npy(x), matrices=npy(torch.stack(mats)),
                        edges=npy(edges), bandwidth=0.5, kde=npy(kde), hard=npy(hard),
                        uproj=npy(uproj), meas=npy(meas), kl=npy(kl), mae=npy(ma), mse=npy(ms),
                        mean_kl=npy(loss), grad_x=npy(xg.grad))

    # ---------------------------------------------------------------- KDE-1D, 2D rotations
    torch.manual_seed(12)
    n, k, nb = 4000, 7, 85
    dist = mf.distributions.get_distribution("swissroll", seed=21)
    x2 = dist.sample(n).float()
    mats2 = rotation_matrices(k)
    edges2 = torch.linspace(-3.5, 3.5, nb + 1)
    diag2 = mf.diagnostics.Histogram1D(axis=0, edges=edges2, bandwidth=0.5)
    tfs2 = [mf.simulate.LinearTransform(m) for m in mats2]
    kde2 = torch.stack([p[0] for p in mf.simulate.forward(x2, tfs2, [[diag2] for _ in tfs2])])
    # direction-based projection with a wider kernel (diagnostics.py:104-106,120-121)
    direction = torch.tensor([0.3, -1.7])
    diag2d = mf.diagnostics.Histogram1D(axis=0, edges=edges2, bandwidth=1.25, direction=direction)
    kde2_dir = torch.stack([p[0] for p in mf.simulate.forward(x2, tfs2, [[diag2d] for _ in tfs2])])
    np.savez_compressed(os.path.join(OUT, "kde1d_2d.npz"), x=npy(x2), matrices=npy(torch.stack(mats2)),
                        edges=npy(edges2), kde=npy(kde2), direction=npy(direction),
                        bandwidth_dir=1.25, kde_dir=npy(kde2_dir))

    # ---------------------------------------------------------------- KDE-2D, 4D corner screens
    torch.manual_seed(13)
    d, n = 4, 2500
    x4 = (torch.randn(n, d) * torch.tensor([1.0, 0.7, 1.3, 0.9])).float()
    x4[:, 2] += 0.5 * x4[:, 0] ** 2 - 0.5
    mats4 = corner_matrices(d)
    ex = torch.linspace(-3.5, 3.5, 34)
    ey = torch.linspace(-4.0, 4.0, 30)
    diag4 = mf.diagnostics.Histogram2D(axis=(0, 2), edges=[ex, ey], bandwidth=(0.5, 0.75))
    tfs4 = [mf.simulate.LinearTransform(m) for m in mats4]
    kde4 = torch.stack([p[0] for p in mf.simulate.forward(x4, tfs4, [[diag4] for _ in tfs4])])
    diag4.kde = False
    hard4 = torch.stack([p[0] for p in mf.simulate.forward(x4, tfs4, [[diag4] for _ in tfs4])])
    diag4.kde = True
    meas4 = hard4 / hard4.sum(dim=(1, 2), keepdim=True) / ((ex[1] - ex[0]) * (ey[1] - ey[0]))
    kl4 = torch.stack([mf.loss.kl_divergence(p, m) for p, m in zip(kde4, torch.roll(meas4, 1, 0))])
    xg = x4.clone().requires_grad_(True)
    pg = mf.simulate.forward(xg, tfs4, [[diag4] for _ in tfs4])
    loss4 = sum(mf.loss.kl_divergence(p[0], m) for p, m in zip(pg, torch.roll(meas4, 1, 0))) / len(tfs4)
    loss4.backward()
    np.savez_compressed(os.path.join(OUT, "kde2d_4d.npz"), x=npy(x4), matrices=npy(torch.stack(mats4)),
                        edges_x=npy(ex), edges_y=npy(ey), bandwidth=np.array([0.5, 0.75]),
                        kde=npy(kde4), hard=npy(hard4), meas=npy(torch.roll(meas4, 1, 0)), kl=npy(kl4),
                        mean_kl=npy(loss4), grad_x=npy(xg.grad))

    # ---------------------------------------------------------------- prior / entropy / loss
    torch.manual_seed(14)
    x6 = torch.randn(2000, 6).float() * 1.2
    logq = (-0.5 * (x6 ** 2).sum(1) - 5.0 + 0.1 * torch.randn(2000)).float()
    prior = mf.prior.Gaussian(ndim=6, scale=3.0)
    lp = prior.log_prob(x6)
    h_mc = mf.entropy.MonteCarloEntropyEstimator(prior=prior)(x6, logq)
    h_mc0 = mf.entropy.MonteCarloEntropyEstimator(prior=None)(x6, logq)
    h_cov = mf.entropy.CovarianceEntropyEstimator()(x6, logq)

    class FixedGenerator(torch.nn.Module):
        def sample_and_log_prob(self, n):
            return x[:n], logq6[:n]

    logq6 = (-0.5 * (x ** 2).sum(1) - 5.5).float()
    model = mf.MENTFlow(transforms=tfs, diagnostics=[[diag] for _ in tfs],
                        measurements=[[m] for m in meas], generator=FixedGenerator(), prior=prior,
                        entropy_estimator=mf.entropy.MonteCarloEntropyEstimator(prior=prior),
                        discrepancy_function=mf.loss.kl_divergence, penalty_parameter=25.0)
    L, H, D = model.loss(2000)
    np.savez_compressed(os.path.join(OUT, "entropy_loss.npz"), x=npy(x6), logq=npy(logq), prior_scale=3.0,
                        prior_log_prob=npy(lp), h_mc=npy(h_mc), h_mc_noprior=npy(h_mc0), h_cov=npy(h_cov),
                        loss_n=2000, loss_logq=npy(logq6), loss_penalty=25.0, loss_L=npy(L), loss_H=npy(H),
                        loss_D=npy(torch.stack(D)))

    # ---------------------------------------------------------------- classical MENT (4D, 1D screens)
    torch.manual_seed(15)
    d, k, nb = 4, 6, 32
    matsm = isotropic_matrices(k, d, seed=3)
    edgesm = torch.linspace(-4.0, 4.0, nb + 1)
    diagm = mf.diagnostics.Histogram1D(axis=0, edges=edgesm, bandwidth=0.5)
    tfm = [mf.simulate.LinearTransform(m) for m in matsm]
    xt = torch.randn(50000, d).float()
    xt[:, 0] *= 1.5
    diagm.kde = False
    measm = [p[0] for p in mf.simulate.forward(xt, tfm, [[diagm] for _ in tfm])]
    diagm.kde = True
    measm = [m / m.sum() / (edgesm[1] - edgesm[0]) for m in measm]
    measm[2][:3] = 0.0            # exercise the g == 0 branches
    res = 12
    sampler = mf.sample.GridSampler(limits=d * [(-4.0, 4.0)], shape=tuple(d * [res]))
    priorm = mf.prior.Gaussian(ndim=d, scale=2.0)
    ment = mf.ment.MENT(ndim=d, transforms=tfm, diagnostics=[[diagm] for _ in tfm],
                        measurements=[[m] for m in measm], prior=priorm, mode="sample",
                        sampler=sampler, n_samples=30000)
    # randomise the tables so that interpolation is exercised
    for i in range(k):
        lf = ment.lagrange_functions[i][0]
        lf.set_values(lf.values * (0.5 + torch.rand(nb)))
    tables0 = torch.stack([ment.lagrange_functions[i][0].values.clone() for i in range(k)])
    xq = torch.randn(3000, d).float() * 1.6
    xq[:4] *= 5.0
    prob_q = ment.prob(xq)
    prob_grid = ment.prob(sampler.get_grid_points())
    torch.manual_seed(99)
    xs = ment.sample(30000)
    torch.manual_seed(99)
    pred0 = ment.simulate(0, 0)
    torch.manual_seed(123)
    ment.gauss_seidel_update(lr=0.9, thresh=1.0e-10)
    tables1 = torch.stack([ment.lagrange_functions[i][0].values.clone() for i in range(k)])
    np.savez_compressed(os.path.join(OUT, "ment_4d.npz"), matrices=npy(torch.stack(matsm)), edges=npy(edgesm),
                        meas=npy(torch.stack(measm)), prior_scale=2.0, grid_res=res, grid_xmax=4.0,
                        tables0=npy(tables0), xq=npy(xq), prob_q=npy(prob_q),
                        prob_grid=npy(prob_grid), sample_seed=99, n_samples=30000,
                        xs_head=npy(xs[:512]), xs_mean=npy(xs.double().mean(0)),
                        xs_cov=npy(torch.cov(xs.double().T)), pred0=npy(pred0), gs_seed=123, lr=0.9,
                        thresh=1.0e-10, tables1=npy(tables1))

    # integrate mode, 2D problem with 1D screens (ment.py:267-317), deterministic
    mats_i = rotation_matrices(5)
    edges_i = torch.linspace(-3.0, 3.0, 25)
    diag_i = mf.diagnostics.Histogram1D(axis=0, edges=edges_i, bandwidth=0.5)
    tf_i = [mf.simulate.LinearTransform(m) for m in mats_i]
    xi = mf.distributions.get_distribution("two-spirals", seed=1).sample(40000).float()
    diag_i.kde = False
    meas_i = [p[0] for p in mf.simulate.forward(xi, tf_i, [[diag_i] for _ in tf_i])]
    diag_i.kde = True
    meas_i = [m / m.sum() / (edges_i[1] - edges_i[0]) for m in meas_i]
    ment_i = mf.ment.MENT(ndim=2, transforms=tf_i, diagnostics=[[diag_i] for _ in tf_i],
                          measurements=[[m] for m in meas_i], prior=mf.prior.Gaussian(ndim=2, scale=5.0),
                          mode="integrate", integration_limits=[[[(-3.0, 3.0)]] for _ in tf_i],
                          integration_shape=[[(40,)] for _ in tf_i])
    pred_i0 = ment_i.simulate(1, 0)
    ment_i.gauss_seidel_update(lr=1.0, thresh=1.0e-10)
    ment_i.gauss_seidel_update(lr=1.0, thresh=1.0e-10)
    tables_i = torch.stack([ment_i.lagrange_functions[i][0].values.clone() for i in range(5)])
    np.savez_compressed(os.path.join(OUT, "ment_2d_integrate.npz"), matrices=npy(torch.stack(mats_i)),
                        edges=npy(edges_i), meas=npy(torch.stack(meas_i)), prior_scale=5.0,
                        int_limits=np.array([-3.0, 3.0]), int_shape=40, pred_1_0=npy(pred_i0),
                        tables_after_2=npy(tables_i))

    # ---------------------------------------------------------------- NSF (restatement, fp64)
    for dd in (2, 6):
        torch.manual_seed(100 + dd)
        flow = NSFOracle(dd).double()
        with torch.no_grad():
            for p in flow.parameters():
                p.mul_(2.5)           # "trained-like": particles spread over spline bins
                p.copy_(p.float().double())   # weights exactly fp32-representable
        z = torch.randn(512, dd, dtype=torch.float64)
        z[:3] *= 3.0                  # a few outside +-5 after a layer or two
        steps = flow.forward_steps(z)
        xx, lq = flow.forward_and_log_prob(z)
        state = {kk: (npy(v).astype(np.float32) if v.is_floating_point() else npy(v))
                 for kk, v in flow.state_dict().items()}
        np.savez_compressed(os.path.join(OUT, f"nsf_{dd}d.npz"), z=npy(z), x=npy(xx), logq=npy(lq),
                            steps=npy(torch.stack(steps)), log_prob_of_x=npy(flow.log_prob(xx)),
                            **{"sd:" + kk: v for kk, v in state.items()})
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f"  {f:28s} {os.path.getsize(os.path.join(OUT, f)) / 1024:8.1f} KB")


def make_multipole():
    """Non-linear transfer maps: MultipoleTransform (simulate/transform.py:78-146) alone for D = 2, 4, 6,
    and the rec_2d/nonlinear composition multipole -> rotation (experiments/rec_2d/nonlinear/setup.py:24-44)
    through the reference's simulate.forward with KDE screens, with the gradient of the mean KL."""
    mf = ref_import.load()
    out = {}
    torch.manual_seed(21)
    cases = []
    for d in (2, 4, 6):
        x = (torch.randn(500, d) * 0.8).float()
        for order in (3, 4, 5):
            for skew in (False, True):
                strength = 0.7 if order < 5 else -0.4
                u = mf.simulate.MultipoleTransform(order=order, strength=strength, skew=skew)(x)
                key = f"d{d}_o{order}_{'s' if skew else 'n'}"
                out[f"kick_x_{key}"] = npy(x)
                out[f"kick_u_{key}"] = npy(u)
                cases.append((d, order, strength, skew))
    out["kick_cases"] = np.array([(d, o, s, int(k)) for d, o, s, k in cases], dtype=np.float64)
    # rec_2d/nonlinear: order 3, strengths linspace(-s, s, num), constant rotation
    num, order, smax, angle = 6, 3, 1.0, math.radians(45.0)
    n, nb = 4000, 85
    x = (torch.randn(n, 2) * torch.tensor([1.0, 0.7])).float()
    edges = torch.linspace(-3.5, 3.5, nb + 1)
    strengths = np.linspace(-smax, smax, num)
    c, s_ = np.cos(angle), np.sin(angle)
    matrix = torch.tensor([[c, s_], [-s_, c]]).type(torch.float32)
    tfs = [mf.simulate.CompositeTransform(mf.simulate.MultipoleTransform(order=order, strength=float(st)),
                                          mf.simulate.LinearTransform(matrix)) for st in strengths]
    diag = mf.diagnostics.Histogram1D(axis=0, edges=edges, bandwidth=0.5)
    diags = [[diag] for _ in tfs]
    kde = torch.stack([p[0] for p in mf.simulate.forward(x, tfs, diags)])
    diag.kde = False
    hard = torch.stack([p[0] for p in mf.simulate.forward(x, tfs, diags)])
    xm = (torch.randn(20000, 2) * torch.tensor([0.8, 1.1])).float()
    meas = torch.stack([p[0] for p in mf.simulate.forward(xm, tfs, diags)])
    diag.kde = True
    meas = meas / meas.sum(dim=1, keepdim=True) / (edges[1] - edges[0])
    xg = x.clone().requires_grad_(True)
    pg = mf.simulate.forward(xg, tfs, diags)
    loss = sum(mf.loss.kl_divergence(p[0], m) for p, m in zip(pg, meas)) / num
    loss.backward()
    out.update(nl_x=npy(x), nl_edges=npy(edges), nl_strengths=strengths, nl_order=order, nl_matrix=npy(matrix),
               nl_kde=npy(kde), nl_hard=npy(hard), nl_meas=npy(meas), nl_mean_kl=npy(loss), nl_grad_x=npy(xg.grad))
    # ProjectionTransform (simulate/transform.py:149-156)
    direction = torch.tensor([0.3, -1.2, 0.5, 2.0])
    x4 = torch.randn(100, 4).float()
    out.update(pt_x=npy(x4), pt_direction=npy(direction), pt_u=npy(mf.simulate.ProjectionTransform(direction)(x4)))
    np.savez_compressed(os.path.join(OUT, "multipole.npz"), **out)
    print("wrote multipole.npz")


def make_ment_2d_screens():
    """Classical MENT with TWO-dimensional screens (ment.py:20-52 N-D LagrangeFunction, :184-199, :267-317;
    experiments/config/rec_nd_2d_ment.yaml: 4-D, corner optics): the density at explicit points and on the
    sampler grid with randomised 2-D tables, a mixed 1-D + 2-D model, and the integration-mode prediction of
    one 2-D screen followed by a Gauss-Seidel sweep (deterministic: no sampling involved)."""
    mf = ref_import.load()
    torch.manual_seed(41)
    d = 4
    mats = corner_matrices(d)                      # 6 axis pairs onto the measured axes (0, 2)
    ex = torch.linspace(-4.0, 4.0, 17)
    ey = torch.linspace(-3.5, 3.5, 15)
    diag = mf.diagnostics.Histogram2D(axis=(0, 2), edges=[ex, ey], bandwidth=(0.5, 0.5))
    tfs = [mf.simulate.LinearTransform(m) for m in mats]
    xt = torch.randn(60000, d).float() * torch.tensor([1.2, 0.8, 1.0, 0.9])
    xt[:, 2] += 0.4 * xt[:, 0]
    diag.kde = False
    meas = [p[0] for p in mf.simulate.forward(xt, tfs, [[diag] for _ in tfs])]
    diag.kde = True
    cell = (ex[1] - ex[0]) * (ey[1] - ey[0])
    meas = [m / m.sum() / cell for m in meas]
    meas[1][:2, :3] = 0.0                          # exercise the g == 0 branch
    res = 9
    sampler = mf.sample.GridSampler(limits=d * [(-4.0, 4.0)], shape=tuple(d * [res]))
    prior = mf.prior.Gaussian(ndim=d, scale=2.0)
    ment = mf.ment.MENT(ndim=d, transforms=tfs, diagnostics=[[diag] for _ in tfs], measurements=[[m] for m in meas],
                        prior=prior, mode="sample", sampler=sampler, n_samples=1000)
    for i in range(len(tfs)):
        lf = ment.lagrange_functions[i][0]
        lf.set_values(lf.values * (0.5 + torch.rand(lf.values.shape)))
    tables0 = torch.stack([ment.lagrange_functions[i][0].values.clone() for i in range(len(tfs))])
    xq = torch.randn(3000, d).float() * 1.5
    xq[:4] *= 5.0
    prob_q = ment.prob(xq)
    prob_grid = ment.prob(sampler.get_grid_points())
    out = dict(matrices=npy(torch.stack(mats)), edges_x=npy(ex), edges_y=npy(ey), meas=npy(torch.stack(meas)),
               prior_scale=2.0, grid_res=res, grid_xmax=4.0, tables0=npy(tables0), xq=npy(xq), prob_q=npy(prob_q),
               prob_grid=npy(prob_grid))
    # mixed model: the six 2-D screens plus three 1-D screens
    mats1 = isotropic_matrices(3, d, seed=5)
    e1 = torch.linspace(-4.0, 4.0, 25)
    diag1 = mf.diagnostics.Histogram1D(axis=0, edges=e1, bandwidth=0.5)
    tfs1 = [mf.simulate.LinearTransform(m) for m in mats1]
    diag1.kde = False
    meas1 = [p[0] for p in mf.simulate.forward(xt, tfs1, [[diag1] for _ in tfs1])]
    diag1.kde = True
    meas1 = [m / m.sum() / (e1[1] - e1[0]) for m in meas1]
    mixed = mf.ment.MENT(ndim=d, transforms=tfs + tfs1, diagnostics=[[diag] for _ in tfs] + [[diag1] for _ in tfs1],
                         measurements=[[m] for m in meas] + [[m] for m in meas1], prior=prior, mode="sample",
                         sampler=sampler, n_samples=1000)
    for i in range(len(tfs)):
        mixed.lagrange_functions[i][0].set_values(tables0[i].clone())
    t1 = []
    for i in range(len(tfs1)):
        lf = mixed.lagrange_functions[len(tfs) + i][0]
        lf.set_values(lf.values * (0.5 + torch.rand(lf.values.shape)))
        t1.append(lf.values.clone())
    out.update(matrices1=npy(torch.stack(mats1)), edges1=npy(e1), meas1=npy(torch.stack(meas1)),
               tables1d=npy(torch.stack(t1)), prob_q_mixed=npy(mixed.prob(xq)))
    # integration mode with 2-D screens: 4-D, the two unmeasured axes on a 13 x 11 grid
    ment_i = mf.ment.MENT(ndim=d, transforms=tfs, diagnostics=[[diag] for _ in tfs], measurements=[[m] for m in meas],
                          prior=prior, mode="integrate",
                          integration_limits=[[[(-4.0, 4.0), (-3.0, 3.0)]] for _ in tfs],
                          integration_shape=[[(13, 11)] for _ in tfs])
    for i in range(len(tfs)):
        ment_i.lagrange_functions[i][0].set_values(tables0[i].clone())
    # reference quirk: _simulate_integrate reshapes with `diagnostic.shape` (ment.py:309), which the reference's
    # Histogram2D never defines -- the attribute is supplied here so that the reference's own arithmetic runs
    diag.shape = (ex.numel() - 1, ey.numel() - 1)
    pred = ment_i.simulate(2, 0)
    ment_i.gauss_seidel_update(lr=0.8, thresh=1.0e-10)
    tables_gs = torch.stack([ment_i.lagrange_functions[i][0].values.clone() for i in range(len(tfs))])
    out.update(int_limits=np.array([[-4.0, 4.0], [-3.0, 3.0]]), int_shape=np.array([13, 11]), pred_2_0=npy(pred), lr=0.8,
               tables_after_gs=npy(tables_gs))
    np.savez_compressed(os.path.join(OUT, "ment_2d_screens.npz"), **out)
    print("wrote ment_2d_screens.npz")


def make_baseline_sized():
    """Reference outputs at the sizes of BASELINE.json's configs (the other fixtures are small on purpose):
    C3  rec_nd_1d: 6-D, K = 100 random projections, 64 bins, N = 20,000 (experiments/rec_nd_1d, run_gmm.sh)
    C4  rec_nd_2d: 6-D, 15 corner screens 85 x 85, N = 5,000
    C5  classical MENT: 6-D, K = 25, sampler grid 8^6: density on the grid (every 5th cell) and one sample-mode
        Gauss-Seidel sweep with 200,000 particles per update (torch CPU generator: statistical comparison only)."""
    mf = ref_import.load()
    # ---- C3
    torch.manual_seed(51)
    d, k, nb, n = 6, 100, 64, 20000
    x = (torch.randn(n, d) * torch.tensor([1.0, 0.6, 1.4, 0.8, 1.1, 0.9])).float()
    x[:, 1] += 0.3 * x[:, 0] ** 2 - 0.3
    mats = isotropic_matrices(k, d, seed=0)
    edges = torch.linspace(-3.5, 3.5, nb + 1)
    diag = mf.diagnostics.Histogram1D(axis=0, edges=edges, bandwidth=0.5)
    tfs = [mf.simulate.LinearTransform(m) for m in mats]
    kde = torch.stack([p[0] for p in mf.simulate.forward(x, tfs, [[diag] for _ in tfs])])
    diag.kde = False
    hard = torch.stack([p[0] for p in mf.simulate.forward(x, tfs, [[diag] for _ in tfs])])
    xm = torch.randn(50000, d).float() * 0.9
    meas = torch.stack([p[0] for p in mf.simulate.forward(xm, tfs, [[diag] for _ in tfs])])
    diag.kde = True
    meas = meas / meas.sum(dim=1, keepdim=True) / (edges[1] - edges[0])
    logq = (-0.5 * (x ** 2).sum(dim=1) - 3.0).float()
    prior = mf.prior.Gaussian(ndim=d, scale=3.0)

    class Fixed(torch.nn.Module):
        def sample_and_log_prob(self, n_):
            return x[:n_], logq[:n_]

    model = mf.MENTFlow(transforms=tfs, diagnostics=[[diag] for _ in tfs], measurements=[[m] for m in meas],
                        generator=Fixed(), prior=prior, entropy_estimator=mf.entropy.MonteCarloEntropyEstimator(prior=prior),
                        discrepancy_function=mf.loss.kl_divergence, penalty_parameter=25.0)
    L, H, D = model.loss(n)
    np.savez_compressed(os.path.join(OUT, "c3_k100.npz"), x=npy(x).astype(np.float32), logq=npy(logq), edges=npy(edges),
                        matrix_seed=0, kde=npy(kde), hard=npy(hard), meas=npy(meas), loss_L=npy(L), loss_H=npy(H),
                        loss_D=npy(torch.stack(D)), prior_scale=3.0, penalty=25.0)
    print("wrote c3_k100.npz")
    # ---- C4
    torch.manual_seed(52)
    n4 = 5000
    x4 = (torch.randn(n4, d) * torch.tensor([1.0, 0.7, 1.3, 0.9, 0.8, 1.1])).float()
    x4[:, 2] += 0.5 * x4[:, 0] ** 2 - 0.5
    mats4 = corner_matrices(d)
    e4 = torch.linspace(-3.5, 3.5, 86)
    diag4 = mf.diagnostics.Histogram2D(axis=(0, 2), edges=[e4, e4.clone()], bandwidth=(0.5, 0.5))
    tfs4 = [mf.simulate.LinearTransform(m) for m in mats4]
    kde4 = torch.stack([p[0] for p in mf.simulate.forward(x4, tfs4, [[diag4] for _ in tfs4])])
    np.savez_compressed(os.path.join(OUT, "c4_6d_85.npz"), x=npy(x4), edges=npy(e4), kde=npy(kde4).astype(np.float32))
    print("wrote c4_6d_85.npz")
    # ---- C5
    torch.manual_seed(53)
    k5, nb5, res = 25, 64, 8
    mats5 = isotropic_matrices(k5, d, seed=7)
    e5 = torch.linspace(-3.5, 3.5, nb5 + 1)
    diag5 = mf.diagnostics.Histogram1D(axis=0, edges=e5, bandwidth=0.5)
    tfs5 = [mf.simulate.LinearTransform(m) for m in mats5]
    xt = torch.randn(100000, d).float() * torch.tensor([1.0, 0.8, 1.2, 0.9, 1.1, 0.7])
    diag5.kde = False
    meas5 = [p[0] for p in mf.simulate.forward(xt, tfs5, [[diag5] for _ in tfs5])]
    diag5.kde = True
    meas5 = [m / m.sum() / (e5[1] - e5[0]) for m in meas5]
    sampler = mf.sample.GridSampler(limits=d * [(-3.5, 3.5)], shape=tuple(d * [res]))
    ment = mf.ment.MENT(ndim=d, transforms=tfs5, diagnostics=[[diag5] for _ in tfs5], measurements=[[m] for m in meas5],
                        prior=mf.prior.Gaussian(ndim=d, scale=2.0), mode="sample", sampler=sampler, n_samples=200000)
    for i in range(k5):
        lf = ment.lagrange_functions[i][0]
        lf.set_values(lf.values * (0.6 + 0.8 * torch.rand(nb5)))
    tables0 = torch.stack([ment.lagrange_functions[i][0].values.clone() for i in range(k5)])
    prob_grid = ment.prob(sampler.get_grid_points())
    torch.manual_seed(54)
    ment.gauss_seidel_update(lr=0.9, thresh=1.0e-10)
    tables1 = torch.stack([ment.lagrange_functions[i][0].values.clone() for i in range(k5)])
    np.savez_compressed(os.path.join(OUT, "c5_ment_6d.npz"), matrix_seed=7, edges=npy(e5), meas=npy(torch.stack(meas5)),
                        prior_scale=2.0, grid_res=res, grid_xmax=3.5, tables0=npy(tables0), prob_grid_stride=5,
                        prob_grid_every5=npy(prob_grid[::5]), prob_grid_sum=npy(prob_grid.double().sum()),
                        n_samples=200000, lr=0.9, tables1=npy(tables1))
    print("wrote c5_ment_6d.npz")


def make_noise():
    """Measurement noise of Histogram.forward (diagnostics/diagnostics.py:50-68): a generator re-seeded on
    every call, multiplicative gaussian / uniform noise, clamped at zero; 1-D and 2-D screens."""
    mf = ref_import.load()
    torch.manual_seed(31)
    x = torch.randn(2000, 4).float()
    edges = torch.linspace(-3.5, 3.5, 33)
    out = {"x": npy(x), "edges": npy(edges)}
    for kind in ("gaussian", "uniform"):
        for seed in (3, 11):
            d1 = mf.diagnostics.Histogram1D(axis=0, edges=edges, bandwidth=0.5, noise=True, noise_scale=0.4,
                                            noise_type=kind, seed=seed)
            d2 = mf.diagnostics.Histogram2D(axis=(0, 2), edges=(edges, edges), bandwidth=(0.5, 0.5), noise=True,
                                            noise_scale=0.4, noise_type=kind, seed=seed)
            out[f"h1_{kind}_{seed}"] = npy(d1(x))
            out[f"h2_{kind}_{seed}"] = npy(d2(x))
    d1 = mf.diagnostics.Histogram1D(axis=0, edges=edges, bandwidth=0.5)
    d2 = mf.diagnostics.Histogram2D(axis=(0, 2), edges=(edges, edges), bandwidth=(0.5, 0.5))
    out["h1_clean"] = npy(d1(x))
    out["h2_clean"] = npy(d2(x))
    np.savez_compressed(os.path.join(OUT, "noise.npz"), **out)
    print("wrote noise.npz")


if __name__ == "__main__":
    if "--only-multipole" in sys.argv:
        make_multipole()
    elif "--only-noise" in sys.argv:
        make_noise()
    elif "--only-ment2d" in sys.argv:
        make_ment_2d_screens()
    elif "--only-baseline-sized" in sys.argv:
        make_baseline_sized()
    else:
        main()
        make_multipole()
        make_noise()
        make_ment_2d_screens()
        make_baseline_sized()
