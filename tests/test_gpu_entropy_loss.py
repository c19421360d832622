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
