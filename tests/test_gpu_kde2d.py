"""GPU parity: fused projection + 2-D KDE / exact 2-D histogram vs reference goldens."""
import pytest
import torch

import mentflow_b200 as mf
from mentflow_b200 import ops
from mfb_testutil import cuda, geom_rows, profile_err, t32
from oracle import hotpath as hp

pytestmark = pytest.mark.gpu
TOL = 1.0e-4


def _setup(g, kde=True):
    mats = t32(g["matrices"])
    ex, ey = t32(g["edges_x"]), t32(g["edges_y"])
    bw = tuple(float(b) for b in g["bandwidth"])
    tfs = [mf.simulate.LinearTransform(m.cuda()) for m in mats]
    diag = mf.diagnostics.Histogram2D(axis=(0, 2), edges=[ex, ey], bandwidth=bw, kde=kde).to("cuda")
    return tfs, [[diag] for _ in tfs]


def test_kde2d_matches_reference_golden(golden):
    g = golden("kde2d_4d")
    tfs, diags = _setup(g)
    x = cuda(g["x"])
    got = torch.stack([o[0] for o in mf.simulate.forward(x, tfs, diags)])
    assert profile_err(got, t32(g["kde"])) < TOL
    again = torch.stack([o[0] for o in mf.simulate.forward(x, tfs, diags)])
    assert torch.equal(got, again)           # fixed-point accumulation: bit-reproducible
    meas = cuda(g["meas"])
    kl = torch.stack([mf.loss.kl_divergence(p, m) for p, m in zip(got, meas)])
    assert torch.allclose(kl.cpu(), t32(g["kl"]), rtol=TOL, atol=1e-7)


def test_hist2d_exact_on_permutation_screens(golden):
    g = golden("kde2d_4d")
    tfs, diags = _setup(g, kde=False)
    x = cuda(g["x"])
    got = torch.stack([o[0] for o in mf.simulate.forward(x, tfs, diags)]).cpu()
    # corner matrices are permutations: projections are exact, so counts are bit-exact
    ref = t32(g["hard"])
    assert torch.allclose(got, ref, rtol=2e-6, atol=0)
    ex, ey = t32(g["edges_x"]), t32(g["edges_y"])
    mats = t32(g["matrices"])
    for k in range(len(mats)):
        uv = hp.linear_map(t32(g["x"]), mats[k])[:, [0, 2]]
        proj = mats[k][[0, 2]].cuda()[None]
        counts = ops.project_hist2d(x, proj, ex.cuda()[None], ey.cuda()[None])[0].cpu()
        assert torch.equal(counts, hp.hist_counts_2d(uv, ex, ey))


def test_kde2d_gradient_matches_reference_autograd(golden):
    g = golden("kde2d_4d")
    tfs, diags = _setup(g)
    meas = cuda(g["meas"])
    x = cuda(g["x"]).requires_grad_(True)
    out = mf.simulate.forward(x, tfs, diags)
    loss = sum(mf.loss.kl_divergence(o[0], m) for o, m in zip(out, meas)) / len(tfs)
    loss.backward()
    assert abs(float(loss) - float(g["mean_kl"])) <= TOL * abs(float(g["mean_kl"]))
    ref = t32(g["grad_x"])
    assert (x.grad.cpu() - ref).abs().max() <= TOL * ref.abs().max()


@pytest.mark.parametrize("n,d,k,bx,by", [(1, 4, 1, 8, 8), (999, 6, 15, 85, 85), (200000, 4, 6, 85, 85), (5000, 3, 2, 20, 31)])
def test_kde2d_vs_oracle_shapes(n, d, k, bx, by):
    gen = torch.Generator().manual_seed(n + k)
    x = torch.randn(n, d, generator=gen)
    w = torch.randn(k, 2, d, generator=gen)
    w = w / w.norm(dim=2, keepdim=True)
    ex, ey = torch.linspace(-3.5, 3.5, bx + 1), torch.linspace(-3.0, 3.0, by + 1)
    (gx, sx), (gy, sy) = geom_rows(ex, 0.5, k), geom_rows(ey, 0.5, k)
    geom = torch.stack([gx, gy], dim=1)
    prof = ops.project_kde2d(x.cuda(), w.cuda(), geom.cuda(), 0.5, bx, by).cpu()
    ref = torch.stack([hp.kde_profile_2d(x @ w[i, 0], x @ w[i, 1], ex, ey, sx, sy, chunk=20000)
                       for i in range(k)])
    assert profile_err(prof, ref) < TOL
