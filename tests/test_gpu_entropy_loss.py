"""GPU parity: entropy estimators, prior, and the assembled MENTFlow loss vs reference goldens."""
import pytest
import torch

import mentflow_b200 as mf
from mfb_testutil import cuda, t32

pytestmark = pytest.mark.gpu
TOL = 1.0e-4


def test_entropy_estimators_match_reference(golden):
    g = golden("entropy_loss")
    x, logq = cuda(g["x"]), cuda(g["logq"])
    prior = mf.prior.Gaussian(ndim=6, scale=float(g["prior_scale"]))
    assert torch.allclose(prior.log_prob(x).cpu(), t32(g["prior_log_prob"]), rtol=1e-5, atol=1e-5)
    h = mf.entropy.MonteCarloEntropyEstimator(prior=prior)(x, logq)
    assert abs(float(h) - float(g["h_mc"])) <= TOL * abs(float(g["h_mc"]))
    h0 = mf.entropy.MonteCarloEntropyEstimator(prior=None)(x, logq)
    assert abs(float(h0) - float(g["h_mc_noprior"])) <= TOL * abs(float(g["h_mc_noprior"]))
    hc = mf.entropy.CovarianceEntropyEstimator()(x, logq)
    assert abs(float(hc) - float(g["h_cov"])) <= TOL * abs(float(g["h_cov"]))


def test_entropy_gradients():
    torch.manual_seed(0)
    x = torch.randn(5000, 6, device="cuda", requires_grad=True)
    lq = torch.randn(5000, device="cuda", requires_grad=True)
    prior = mf.prior.Gaussian(ndim=6, scale=3.0)
    h = mf.entropy.MonteCarloEntropyEstimator(prior=prior)(x, lq)
    h.backward()
    xr = x.detach().clone().requires_grad_(True)
    lr = lq.detach().clone().requires_grad_(True)
    href = torch.mean(lr) - torch.mean(prior.log_prob(xr))
    href.backward()
    assert abs(float(h) - float(href)) < 1e-5
    assert torch.allclose(x.grad, xr.grad, rtol=1e-5, atol=1e-9)
    assert torch.allclose(lq.grad, lr.grad, rtol=1e-6)


def test_covariance_entropy_gradient_matches_reference_expression():
    """entropy.py:27-38 is differentiable in the reference (torch.cov / torch.det); so is the kernel path."""
    import numpy as np
    torch.manual_seed(1)
    a = torch.randn(6, 6, device="cuda") * 0.7
    x = (torch.randn(20000, 6, device="cuda") @ a).requires_grad_(True)
    est = mf.entropy.CovarianceEntropyEstimator()
    h = est(x, None)
    (3.0 * h).backward()
    xr = x.detach().double().requires_grad_(True)
    cov = torch.cov(xr.T)
    href = -3.0 * np.log(2.0 * np.pi * np.e) - torch.log(torch.sqrt(torch.det(cov)) + est.pad)
    (3.0 * href).backward()
    assert abs(float(h) - float(href)) <= 1e-5 * abs(float(href))
    scale = float(xr.grad.abs().max())
    assert float((x.grad.double() - xr.grad).abs().max()) <= 1e-4 * scale


def test_mentflow_loss_matches_reference(golden):
    g, k = golden("entropy_loss"), golden("kde1d_6d")
    mats, edges, meas = t32(k["matrices"]), t32(k["edges"]), cuda(k["meas"])
    n = int(g["loss_n"])
    x, logq = cuda(k["x"])[:n].contiguous(), cuda(g["loss_logq"])[:n].contiguous()

    class Fixed(torch.nn.Module):
        def sample_and_log_prob(self, size):
            return x[:size], logq[:size]

    prior = mf.prior.Gaussian(ndim=6, scale=float(g["prior_scale"]))
    tfs = [mf.simulate.LinearTransform(m.cuda()) for m in mats]
    diag = mf.diagnostics.Histogram1D(axis=0, edges=edges, bandwidth=0.5).to("cuda")
    model = mf.MENTFlow(transforms=tfs, diagnostics=[[diag] for _ in tfs], measurements=[[m] for m in meas],
                        generator=Fixed(), prior=prior,
                        entropy_estimator=mf.entropy.MonteCarloEntropyEstimator(prior=prior),
                        discrepancy_function=mf.loss.kl_divergence, penalty_parameter=float(g["loss_penalty"]))
    L, H, D = model.loss(n)
    assert abs(float(L) - float(g["loss_L"])) <= TOL * abs(float(g["loss_L"]))
    assert abs(float(H) - float(g["loss_H"])) <= TOL * abs(float(g["loss_H"]))
    assert torch.allclose(torch.stack(D).cpu(), t32(g["loss_D"]), rtol=TOL, atol=1e-8)


def test_graphed_loss_matches_eager_and_tracks_weight_updates():
    """CUDA-graph replay of the forward pass (mentflow_b200.graphs.GraphedLoss) gives the eager
    result bit for bit, accepts pinned host noise, and re-captures after the weights change."""
    import torch
    import mentflow_b200 as mf
    from mentflow_b200 import workloads
    from mentflow_b200.graphs import GraphedLoss
    torch.manual_seed(3)
    dev = torch.device("cuda")
    wl = workloads.isotropic_1d(ndim=6, num=12, bins=64, xmax=3.5, seed=0)
    gen = mf.generate.NSFGenerator(6).to(dev)
    tfs = [mf.simulate.LinearTransform(m.to(dev)) for m in wl["matrices"]]
    diag = mf.diagnostics.Histogram1D(axis=0, edges=wl["edges"], bandwidth=0.5).to(dev)
    diags = [[diag] for _ in tfs]
    truth = workloads.gaussian_mixture(20_000, ndim=6, seed=1, device=dev)
    diag.kde = False
    meas = mf.simulate.forward(truth, tfs, diags)
    diag.kde = True
    width = float(wl["edges"][1] - wl["edges"][0])
    meas = [[m[0] / m[0].sum() / width] for m in meas]
    prior = mf.prior.Gaussian(ndim=6, scale=3.0)
    model = mf.MENTFlow(transforms=tfs, diagnostics=diags, measurements=meas, generator=gen, prior=prior,
                        entropy_estimator=mf.entropy.MonteCarloEntropyEstimator(prior=prior),
                        discrepancy_function=mf.loss.kl_divergence, penalty_parameter=10.0)
    n = 20_000
    g = GraphedLoss(model, n)
    z = torch.randn(n, 6)
    zp = z.pin_memory()

    def eager(zz):
        with torch.no_grad():
            x, lq = gen.forward_and_log_prob(zz.to(dev))
            return model.loss_from_particles(x, lq)

    L, H, D = g(zp)
    Le, He, De = eager(z)
    assert torch.equal(L, Le) and torch.equal(H, He) and torch.equal(torch.stack(D), torch.stack(De))
    with torch.no_grad():
        gen.w_out.mul_(1.5)        # "optimiser step": the next call must see the new weights
    L2 = g(zp)[0].clone()
    assert torch.equal(L2, eager(z)[0]) and not torch.equal(L2, Le)
    L3 = g(None)[0]                # noise drawn on the device
    assert torch.isfinite(L3)


def test_graphed_train_step_reduces_the_loss():
    """zero_grad + loss + backward + AdamW step as one CUDA-graph replay: the loss of a small 2-D
    reconstruction goes down over 30 replays and the parameters stay finite."""
    import torch
    import mentflow_b200 as mf
    from mentflow_b200 import workloads
    from mentflow_b200.graphs import GraphedTrainStep
    torch.manual_seed(0)
    dev = torch.device("cuda")
    wl = workloads.rotations_2d(7, 64, 3.5)
    gen = mf.generate.NSFGenerator(2).to(dev)
    tfs = [mf.simulate.LinearTransform(m.to(dev)) for m in wl["matrices"]]
    diag = mf.diagnostics.Histogram1D(axis=0, edges=wl["edges"], bandwidth=0.5).to(dev)
    diags = [[diag] for _ in tfs]
    truth = workloads.gaussian_mixture(50_000, ndim=2, seed=1, device=dev)
    with torch.no_grad():
        meas = [[p[0]] for p in mf.simulate.forward(truth, tfs, diags)]
    prior = mf.prior.Gaussian(ndim=2, scale=3.0)
    model = mf.MENTFlow(transforms=tfs, diagnostics=diags, measurements=meas, generator=gen, prior=prior,
                        entropy_estimator=mf.entropy.MonteCarloEntropyEstimator(prior=prior),
                        discrepancy_function=mf.loss.kl_divergence, penalty_parameter=50.0)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3, weight_decay=0.0, capturable=True)
    step = GraphedTrainStep(model, opt, 20_000)
    first = float(step()[0])
    for _ in range(30):
        L, H, D = step()
    last = float(L)
    assert bool(step.finite) and last < first, (first, last)
    assert all(torch.isfinite(p).all() for p in model.parameters())


def test_graphed_train_step_equals_the_eager_loop_and_honours_lr_changes():
    """The reference's training loop body (train/train.py:164-169, :205-207) as graph replays: after N steps the
    parameters equal those of the eager loop on the same seed (the capture's warm-up passes leave parameters,
    optimiser state and the random stream untouched), and a scheduler changing a float learning rate is honoured
    (the step is re-captured)."""
    import copy
    import torch
    import mentflow_b200 as mf
    from mentflow_b200 import workloads
    from mentflow_b200.graphs import GraphedTrainStep
    dev = torch.device("cuda")
    wl = workloads.isotropic_1d(ndim=4, num=8, bins=48, xmax=3.5, seed=2)
    tfs = [mf.simulate.LinearTransform(m.to(dev)) for m in wl["matrices"]]
    diag = mf.diagnostics.Histogram1D(axis=0, edges=wl["edges"], bandwidth=0.5).to(dev)
    diags = [[diag] for _ in tfs]
    truth = workloads.gaussian_mixture(50_000, ndim=4, seed=1, device=dev)
    with torch.no_grad():
        meas = [[p[0]] for p in mf.simulate.forward(truth, tfs, diags)]
    prior = mf.prior.Gaussian(ndim=4, scale=3.0)

    def make():
        torch.manual_seed(5)
        gen = mf.generate.NSFGenerator(4).to(dev)
        model = mf.MENTFlow(transforms=tfs, diagnostics=diags, measurements=meas, generator=gen, prior=prior,
                            entropy_estimator=mf.entropy.MonteCarloEntropyEstimator(prior=prior),
                            discrepancy_function=mf.loss.kl_divergence, penalty_parameter=20.0)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.0, capturable=True)
        return model, opt

    n, steps = 10_000, 6
    lrs = [1e-3, 1e-3, 1e-3, 4e-4, 4e-4, 4e-4]            # ReduceLROnPlateau-style drop after three steps
    # eager loop
    model_e, opt_e = make()
    torch.manual_seed(123)
    losses_e = []
    for i in range(steps):
        for g_ in opt_e.param_groups:
            g_["lr"] = lrs[i]
        opt_e.zero_grad(set_to_none=True)
        L, H, D = model_e.loss(n)
        L.backward()
        opt_e.step()
        losses_e.append(float(L))
    # graph replays
    model_g, opt_g = make()
    before = [p.detach().clone() for p in model_g.parameters()]
    step = GraphedTrainStep(model_g, opt_g, n)
    torch.manual_seed(123)
    step._capture()                                          # explicit capture: must be side-effect free
    assert all(torch.equal(a, b) for a, b in zip(before, model_g.parameters()))
    assert torch.cuda.default_generators[dev.index or 0].get_offset() == 0
    losses_g = []
    for i in range(steps):
        for g_ in opt_g.param_groups:
            g_["lr"] = lrs[i]
        losses_g.append(float(step()[0]))
    assert losses_g == losses_e, (losses_g, losses_e)
    for a, b in zip(model_g.parameters(), model_e.parameters()):
        assert torch.allclose(a, b, rtol=0, atol=1e-7), float((a - b).abs().max())
    st_g = opt_g.state[next(iter(model_g.parameters()))]["step"]
    assert int(st_g) == steps                                # every counted step is one replay


def test_scalar_tail_kernels_match_torch():
    """mfb_mc_entropy and mfb_loss_tail (one launch each) against the expressions they replace, values and gradients
    (entropy.py:58-62, core.py:111-113)."""
    from mentflow_b200 import ops
    torch.manual_seed(3)
    m = torch.tensor([-1234.5678, 9876.54321], dtype=torch.float64, device="cuda")
    a, b, c = 1.0 / 1000.0, 0.5 / 9.0 / 1000.0, -3.25
    h = ops.mc_entropy(m, a, b, c)
    want = (m[0] * a + m[1] * b - c).float()
    assert h.dtype == torch.float32 and abs(float(h) - float(want)) <= 1e-6 * abs(float(want))
    for k in (1, 7, 100, 333):
        d = torch.rand(k, device="cuda", requires_grad=True)
        hh = torch.tensor(0.7, device="cuda", requires_grad=True)
        L = ops.LossTail.apply(hh, d, 12.5)
        ref = hh.detach() + 12.5 * d.detach().mean()
        assert abs(float(L.detach()) - float(ref)) <= 2e-6 * abs(float(ref))
        (L * 3.0).backward()
        assert abs(float(hh.grad) - 3.0) < 1e-6
        assert torch.allclose(d.grad, torch.full_like(d, 3.0 * 12.5 / k), rtol=1e-6)
        L0 = ops.LossTail.apply(None, d.detach(), 2.0)
        assert abs(float(L0) - 2.0 * float(d.detach().mean())) <= 2e-6
