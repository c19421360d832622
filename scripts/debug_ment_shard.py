"""world_size-2 debug: sharded MENT draw / profile against the single-GPU one (run under torchrun)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch, torch.distributed as dist
import mentflow_b200 as mf
from mentflow_b200 import distributed as mfd, workloads
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
d, K, res, n = 6, 5, 12, 2_000_000
wl = workloads.isotropic_1d(d, K, 64, 3.5)
tfs = [mf.simulate.LinearTransform(m.to(dev)) for m in wl["matrices"]]
diag = mf.diagnostics.Histogram1D(axis=0, edges=wl["edges"], bandwidth=0.5).to(dev)
diags = [[diag] for _ in tfs]
truth = workloads.gaussian_mixture(200_000, ndim=d, seed=1, device=dev)
with torch.no_grad():
    meas = [[p[0]] for p in mf.simulate.forward(truth, tfs, diags)]
def make(shard):
    s = mf.sample.GridSampler(limits=d * [(-3.5, 3.5)], shape=tuple(d * [res]), device=dev)
    m = mf.ment.MENT(ndim=d, transforms=tfs, diagnostics=diags, measurements=meas, prior=mf.prior.Gaussian(ndim=d, scale=3.0),
                     mode="sample", sampler=s, n_samples=n, device=dev)
    if shard: mfd.shard_model(m)
    return m
a, b = make(True), make(False)
torch.manual_seed(5); xa = a.sample(n)
torch.manual_seed(5); xb = b.sample(n)
sl = mfd.shard_slice(n, rank, world)
print(rank, "shard", a.shard, "local", tuple(xa.shape), "slice equal:", bool(torch.equal(xa, xb[sl])), flush=True)
torch.manual_seed(6); pa = a.simulate(1, 0)
torch.manual_seed(6); pb = b.simulate(1, 0)
print(rank, "profile rel diff", float((pa - pb).abs().max() / pb.abs().max()), "reducer calls", a.reducer.calls, flush=True)
torch.manual_seed(7); a.gauss_seidel_update(lr=0.9)
torch.manual_seed(7); b.gauss_seidel_update(lr=0.9)
ta = torch.stack([lf[0].values for lf in a.lagrange_functions]); tb = torch.stack([lf[0].values for lf in b.lagrange_functions])
r = ((ta - tb).abs() / tb.abs().clamp_min(1e-6)).flatten()
print(rank, "tables rel diff median", float(r.median()), "max", float(r.max()), flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
