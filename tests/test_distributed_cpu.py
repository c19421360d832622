"""world_size-2 gloo tests (CPU) of the particle-sharding logic: shard bookkeeping, the reducer
that all-reduces unnormalised profile sums before the non-linear normalisation, and the gradient
all-reduce.  The payloads are produced by the CPU oracle (the CUDA kernels need a GPU); what is
tested here is that sharded + reduced == unsharded, which is what makes the N>1 path exact."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mentflow_b200 import distributed as mfd
from oracle import hotpath as hp


def test_shard_sizes_cover_everything():
    for total, world in [(10, 3), (1_000_003, 8), (5, 8), (0, 2)]:
        sizes = mfd.shard_sizes(total, world)
        assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
        covered = []
        for r in range(world):
            sl = mfd.shard_slice(total, r, world)
            covered += list(range(total))[sl]
        assert covered == list(range(total))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, equal, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        n = 4000 if equal else 4001
        x = torch.randn(n, 3)
        logq = torch.randn(n)
        edges = torch.linspace(-3, 3, 33)
        w = torch.tensor([0.6, -0.8, 0.0])
        sl = mfd.shard_slice(n, rank, world)
        reducer = mfd.ShardReducer(equal_shards=equal)
        # forward exchange: unnormalised KDE sums (fp32), exact histogram counts (int64), entropy sums (fp64)
        sums = hp.kde_sums_1d(x[sl] @ w, edges, 0.5 * float(edges[1] - edges[0])).float()[None]
        n_total = reducer(sums, float(x[sl].shape[0]))
        counts = hp.hist_counts_1d(x[sl] @ w, edges)[None].clone()
        reducer(counts, 0.0)
        mom = torch.stack([logq[sl].double().sum(), (x[sl].double() ** 2).sum()])
        reducer(mom, float(x[sl].shape[0]))
        # packed exchange (one all-reduce per forward step): the entropy sums ride as (hi, lo) float pairs at
        # the tail of the float32 buffer of the profile sums.  The split / join kernels are CUDA; their
        # arithmetic is restated here so that the host logic runs under gloo
        from mentflow_b200 import ops as _ops

        def _split(values, out):
            hi = values.float()
            out[: values.numel()] = hi
            out[values.numel():] = (values - hi.double()).float()

        _ops.f64_split = _split
        _ops.f64_join = lambda pairs: pairs[: pairs.numel() // 2].double() + pairs[pairs.numel() // 2:].double()
        packed_ok = True
        if equal:
            mom2 = torch.stack([logq[sl].double().sum(), (x[sl].double() ** 2).sum()])
            reducer.stash(mom2)
            assert reducer.tail_floats() == 4
            flat = torch.empty(32 + 4)
            sums2 = flat[:32].view(1, 32)
            sums2.copy_(hp.kde_sums_1d(x[sl] @ w, edges, 0.5 * float(edges[1] - edges[0])).float()[None])
            calls0 = reducer.calls
            n2 = reducer(sums2, float(x[sl].shape[0]), flat=flat)
            got = reducer.pop_result()
            packed_ok = (reducer.calls == calls0 + 1 and n2 == n and reducer.tail_floats() == 0
                         and torch.allclose(sums2, sums, rtol=1e-6)
                         and torch.allclose(got, mom, rtol=1e-6) and reducer.pop_result() is None)
            reducer.calls = calls0
        # backward exchange: flattened gradient all-reduce
        p1, p2 = torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 3))
        p1.grad = torch.full((5,), float(rank + 1))
        p2.grad = torch.arange(6.0).reshape(2, 3) * (rank + 1)
        mfd.allreduce_gradients([p1, p2])
        if rank == 0:
            full_sums = hp.kde_sums_1d(x @ w, edges, 0.5 * float(edges[1] - edges[0])).float()[None]
            ok = (n_total == n
                  and torch.allclose(sums, full_sums, rtol=1e-6)
                  and torch.equal(counts[0], hp.hist_counts_1d(x @ w, edges))
                  and torch.allclose(mom, torch.stack([logq.double().sum(), (x.double() ** 2).sum()]), rtol=1e-12)
                  and torch.equal(p1.grad, torch.full((5,), 3.0))
                  and torch.equal(p2.grad, torch.arange(6.0).reshape(2, 3) * 3)
                  and reducer.calls == 3 and packed_ok)
            out.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("equal", [True, False])
def test_sharded_reduction_equals_unsharded(equal):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, equal, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.get(timeout=10) is True
