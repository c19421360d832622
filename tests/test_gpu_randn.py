"""GPU parity of the library's Philox base-noise kernel with torch.randn on the same generator state
(generate/flows/zuko.py:15-16: the reference draws z with torch through zuko's DiagNormal)."""
import pytest
import torch

import mentflow_b200 as mf
from mentflow_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 2), (7, 3), (1000, 6), (100_003, 6), (1_000_000, 6), (2_500_000, 4)])
def test_sample_base_is_torch_randn_bitwise(shape):
    n, d = shape
    gen = mf.generate.NSFGenerator(d).to("cuda")
    torch.manual_seed(1234)
    warm = torch.randn(5, device="cuda")                 # a non-zero offset
    want = torch.randn(n, d, device="cuda")
    after = torch.randn(11, device="cuda")
    torch.manual_seed(1234)
    warm2 = torch.randn(5, device="cuda")
    got = gen.sample_base(n)
    after2 = torch.randn(11, device="cuda")              # torch's stream continues where it would have
    assert torch.equal(warm, warm2)
    assert got.shape == (n, d) and torch.equal(got, want)
    assert torch.equal(after, after2)


def test_raw_entry_point_and_increment():
    from mentflow_b200 import _lib
    lib = _lib.load()
    g = torch.cuda.default_generators[torch.cuda.current_device()]
    for numel in (1, 255, 256, 257, 303_104 * 4, 303_104 * 4 + 1, 7_000_001):
        torch.manual_seed(7)
        off0 = g.get_offset()
        want = torch.randn(numel, device="cuda")
        inc = g.get_offset() - off0
        assert lib.mfb_randn_offset_increment(numel) == inc
        out = torch.empty(numel, device="cuda")
        ops.PhiloxStream("cuda").__class__  # noqa: B018  (import check)
        torch.manual_seed(7)
        ops.PhiloxStream(torch.device("cuda", torch.cuda.current_device())).normal_(out)
        assert torch.equal(out, want)


def test_graph_replay_draws_fresh_noise_and_follows_the_seed(golden):
    from mentflow_b200 import workloads
    from mentflow_b200.graphs import GraphedLoss
    dev = torch.device("cuda", torch.cuda.current_device())
    wl = workloads.isotropic_1d(ndim=4, num=6, bins=32, xmax=3.5, seed=0)
    torch.manual_seed(0)
    gen = mf.generate.NSFGenerator(4).to(dev)
    tfs = [mf.simulate.LinearTransform(m.to(dev)) for m in wl["matrices"]]
    diag = mf.diagnostics.Histogram1D(axis=0, edges=wl["edges"], bandwidth=0.5).to(dev)
    diags = [[diag] for _ in tfs]
    truth = torch.randn(20000, 4, device=dev)
    with torch.no_grad():
        meas = [[p[0]] for p in mf.simulate.forward(truth, tfs, diags)]
    prior = mf.prior.Gaussian(ndim=4, scale=3.0)
    model = mf.MENTFlow(transforms=tfs, diagnostics=diags, measurements=meas, generator=gen, prior=prior,
                        entropy_estimator=mf.entropy.MonteCarloEntropyEstimator(prior=prior),
                        discrepancy_function=mf.loss.kl_divergence, penalty_parameter=10.0)
    n = 30_000
    graphed = GraphedLoss(model, n)
    torch.manual_seed(99)
    a = [float(graphed()[0]) for _ in range(3)]
    assert len(set(a)) == 3                                  # fresh noise at every replay
    torch.manual_seed(99)
    b = [float(graphed()[0]) for _ in range(3)]
    assert a == b                                            # and it follows torch's seed
    # eager model.loss(n) on the same seed sees the same particles: identical numbers
    torch.manual_seed(99)
    with torch.no_grad():
        c = [float(model.loss(n)[0]) for _ in range(3)]
    assert a == c
    # ... and they are torch.randn's particles
    torch.manual_seed(99)
    with torch.no_grad():
        z = torch.randn(n, 4, device=dev)
        x, lq = gen.forward_and_log_prob(z)
        assert float(model.loss_from_particles(x, lq)[0]) == a[0]
