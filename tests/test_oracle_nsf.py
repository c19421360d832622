"""Property tests defending oracle/zuko_nsf.py (parity vs. real zuko is unpinned).
CPU only."""
import math

import numpy as np
import pytest
import torch

from oracle.zuko_nsf import NSFOracle, RQSpline, layer_order, masked_mlp_masks


def make_flow(d, seed=0, scale=2.5, dtype=torch.float64, **kw):
    torch.manual_seed(seed)
    flow = NSFOracle(d, **kw).to(dtype)
    with torch.no_grad():
        for p in flow.parameters():
            p.mul_(scale)
    return flow


def test_mask_counts_match_survey():
    for layer in (0, 1):
        masks = masked_mlp_masks(layer_order(6, layer), 59, [64] * 3)
        assert [tuple(m.shape) for m in masks] == [(64, 6), (64, 64), (64, 64), (354, 64)]
        assert [int(m.sum()) for m in masks] == [190, 2458, 2458, 11446]
    assert sum(p.numel() for p in NSFOracle(6).parameters()) == 158890


def test_golden_fixture_is_reproduced(golden):
    for d in (2, 6):
        g = golden(f"nsf_{d}d")
        flow = NSFOracle(d).double()
        sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd:")}
        flow.load_state_dict({k: (v.double() if v.dtype == torch.float32 else v) for k, v in sd.items()})
        z = torch.from_numpy(g["z"])
        x, lq = flow.forward_and_log_prob(z)
        assert torch.allclose(x, torch.from_numpy(g["x"]), atol=1e-12)
        assert torch.allclose(lq, torch.from_numpy(g["logq"]), atol=1e-11)
        steps = torch.stack(flow.forward_steps(z))
        assert torch.allclose(steps, torch.from_numpy(g["steps"]), atol=1e-12)


@pytest.mark.parametrize("d", [2, 4, 6])
def test_autoregressive_triangularity_and_ladj(d):
    flow = make_flow(d, seed=d)
    z = torch.randn(6, d, dtype=torch.float64)
    for li, layer in enumerate(flow.layers):
        order = layer_order(d, li)
        for row in z:
            J = torch.autograd.functional.jacobian(lambda a: layer(a[None])[0], row)
            # output i depends on input j only if order[i] >= order[j]
            for i in range(d):
                for j in range(d):
                    if order[i] < order[j]:
                        assert J[i, j] == 0.0
            _, ladj = layer.call_and_ladj(row[None])
            assert abs(torch.log(torch.abs(torch.det(J))) - ladj[0]) < 1e-9
    x, lq = flow.forward_and_log_prob(z)
    for i in range(3):
        J = torch.autograd.functional.jacobian(lambda a: flow(a[None])[0], z[i])
        want = flow.base_log_prob(z[i:i + 1])[0] - torch.logdet(J)
        assert abs(want - lq[i]) < 1e-9


@pytest.mark.parametrize("d", [2, 6])
def test_inverse_round_trip_and_log_prob(d):
    flow = make_flow(d, seed=10 + d, scale=1.5)
    z = torch.randn(400, d, dtype=torch.float64)
    x, lq = flow.forward_and_log_prob(z)
    assert (flow.inverse(x) - z).abs().max() < 1e-6
    assert (flow.log_prob(x) - lq).abs().max() < 1e-6


def test_spline_is_monotone_identity_outside_and_unit_end_slopes():
    torch.manual_seed(3)
    w, h = torch.randn(1, 20, dtype=torch.float64) * 3, torch.randn(1, 20, dtype=torch.float64) * 3
    dd = torch.randn(1, 19, dtype=torch.float64) * 3
    sp = RQSpline(w, h, dd)
    xs = torch.linspace(-7, 7, 4001, dtype=torch.float64)[:, None]
    y, ladj = sp.call_and_ladj(xs)
    assert (torch.diff(y[:, 0]) > 0).all()
    out = xs[:, 0].abs() > 5
    assert torch.equal(y[out], xs[out]) and (ladj[out] == 0).all()
    assert sp.derivatives[0, 0] == 1.0 and sp.derivatives[0, -1] == 1.0
    assert abs(float(sp.horizontal[0, 0]) + 5) < 1e-12 and abs(float(sp.horizontal[0, -1]) - 5) < 1e-12
    # ladj equals log of the numerical derivative
    xin = xs[~out][:, None] if False else xs[(xs[:, 0].abs() < 4.9)]
    xin = xin.clone().requires_grad_(True)
    yy, ll = sp.call_and_ladj(xin)
    (gy,) = torch.autograd.grad(yy.sum(), xin)
    assert (gy.log() - ll).abs().max() < 1e-9
    # inverse
    assert (sp.inverse(y) - xs).abs().max() < 1e-9


def test_first_feature_of_each_layer_is_bias_only():
    flow = make_flow(6, seed=5)
    z1 = torch.randn(5, 6, dtype=torch.float64)
    z2 = z1.clone()
    z2[:, 1:] = torch.randn(5, 5, dtype=torch.float64)
    # layer 0 (natural order): feature 0's spline is unconditional
    a, b = flow.layers[0](z1), flow.layers[0](z2)
    assert torch.equal(a[:, 0], b[:, 0])


def test_zuko_key_layout_if_available():
    """the day zuko is importable, (a) the restatement must reproduce it numerically and
    (b) a reference-built flow's state_dict keys must be one of the two spellings mentflow_b200 reads
    (tests/golden/zuko_expected_keys.txt).  Skipped in this image (zuko 1.3.1 is not installable offline), which is
    why the flow oracle's parity is reported as unpinned."""
    import os
    zuko = pytest.importorskip("zuko")
    from oracle.zuko_nsf import cross_check_against_zuko
    assert cross_check_against_zuko(features=6, seed=0, n=2048) < 1e-10
    flow = zuko.flows.NSF(6, transforms=5, hidden_features=[64] * 3, bins=20)
    inv = zuko.flows.Flow(flow.transform.inv, flow.base)
    keys = {"_flow." + k for k in inv.state_dict().keys()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "zuko_expected_keys.txt")
    text = open(path).read()
    a = {l for l in text.split("# spelling B")[0].splitlines() if l and not l.startswith("#")}
    b = {l for l in text.split("# spelling B")[1].splitlines() if l and not l.startswith("#")}
    weights = lambda s: {k for k in s if k.endswith(("weight", "bias"))}
    assert weights(keys) in (weights(a), weights(b)), sorted(keys)[:6]
