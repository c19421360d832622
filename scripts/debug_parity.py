"""Diagnostic (not a test): print parity errors of the CUDA kernels vs fp64 oracle next to the
error of a plain torch-fp32 evaluation of the same thing."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mentflow_b200 as mf
from mentflow_b200 import ops
from mfb_testutil import *
from oracle import hotpath as hp

def nsf_case(name, gen, z):
    r64 = oracle_from_generator(gen); r32 = oracle_from_generator(gen, torch.float32)
    with torch.no_grad():
        x64, l64 = r64.forward_and_log_prob(z.double()); x32, l32 = r32.forward_and_log_prob(z.float())
        x, l = gen.to("cuda").forward_and_log_prob(z.float().cuda())
    def stats(a, b):
        e = ((a.double().cpu() - b).abs() / b.abs().clamp_min(1)).flatten()
        return f"max {e.max():.2e} p99.99 {torch.quantile(e[:4000000], 0.9999):.2e} median {e.median():.2e} n>1e-4 {(e>1e-4).sum()}/{e.numel()}"
    print(name, "\n  cuda  x:", stats(x, x64), "\n  torch x:", stats(x32, x64), "\n  cuda  l:", stats(l, l64), "\n  torch l:", stats(l32, l64))

for d in (2, 6):
    g = dict(np.load(os.path.join(ROOT, f"tests/golden/nsf_{d}d.npz")))
    nsf_case(f"golden {d}d", generator_from_golden(g), torch.from_numpy(g["z"]))
for d, n, scale in [(3, 257, 3.0), (6, 100003, 3.0), (6, 100003, 1.0)]:
    torch.manual_seed(d * 10 + 1)
    gen = mf.generate.NSFGenerator(d)
    with torch.no_grad():
        for p in gen.parameters(): p.mul_(scale)
    z = torch.randn(n, d); z[: max(1, n // 100)] *= 4.0
    nsf_case(f"d={d} n={n} scale={scale}", gen, z)

# kde1d ragged case
n, d, k, nb = 2000, 8, 4, 200
gen = torch.Generator().manual_seed(n + d + k)
x = torch.randn(n, d, generator=gen); w = torch.randn(k, d, generator=gen); w = w / w.norm(dim=1, keepdim=True)
edges = torch.linspace(-3.5, 3.5, nb + 1); delta = float(edges[1] - edges[0])
geom, _ = geom_rows(edges, 0.5, k)
sums = ops.kde1d_sums(x.cuda(), w.cuda(), geom.cuda(), 0.5, nb).cpu().double()
ref = torch.stack([hp.kde_sums_1d(x @ w[i], edges, 0.5 * delta) for i in range(k)])
ref32 = torch.stack([hp._kernel_matrix(x @ w[i], hp.centres(edges), 0.5 * delta).sum(0) for i in range(k)]).double()
print("kde1d 2000x8 k4 b200: cuda err", float((sums - ref).abs().max() / ref.abs().max()), "torch32 err", float((ref32 - ref).abs().max() / ref.abs().max()))

# kde2d gradient
g = dict(np.load(os.path.join(ROOT, "tests/golden/kde2d_4d.npz")))
mats = t32(g["matrices"]); ex, ey = t32(g["edges_x"]), t32(g["edges_y"]); bw = tuple(float(b) for b in g["bandwidth"])
tfs = [mf.simulate.LinearTransform(m.cuda()) for m in mats]
diag = mf.diagnostics.Histogram2D(axis=(0, 2), edges=[ex, ey], bandwidth=bw).to("cuda")
meas = cuda(g["meas"]); xx = cuda(g["x"]).requires_grad_(True)
out = mf.simulate.forward(xx, tfs, [[diag] for _ in tfs])
loss = sum(mf.loss.kl_divergence(o[0], m) for o, m in zip(out, meas)) / len(tfs); loss.backward()
ref = t32(g["grad_x"])
# fp64 oracle gradient
x64 = t32(g["x"]).double().requires_grad_(True)
scr = [[hp.Screen2D(axis=(0, 2), edges_x=ex.double(), edges_y=ey.double(), bandwidth=bw)] for _ in mats]
o64 = hp.simulate(x64, [m.double() for m in mats], scr)
l64 = sum(hp.kl_div(o[0], m.double()) for o, m in zip(o64, t32(g["meas"]))) / len(mats); l64.backward()
print("kde2d loss", float(loss), float(g["mean_kl"]), float(l64))
gm = x64.grad.abs().max()
print("kde2d grad: cuda vs fp64", float((xx.grad.cpu().double() - x64.grad).abs().max() / gm), "ref32 vs fp64", float((ref.double() - x64.grad).abs().max() / gm),
      "cuda vs ref32", float((xx.grad.cpu() - ref).abs().max() / ref.abs().max()))
prof = torch.stack([o[0] for o in out]).detach().cpu().double(); p64 = torch.stack([o[0] for o in o64]).detach()
print("kde2d prof err vs fp64", float(((prof - p64).abs().amax(dim=(1, 2)) / p64.amax(dim=(1, 2))).max()), "min-bin ratio check", float((prof[p64 > 1e-14] / p64[p64 > 1e-14]).min()), float((prof[p64 > 1e-14] / p64[p64 > 1e-14]).max()))

# ---- NSF backward diagnostics
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_nsf import _grads_case
for args in [(6, 3000, 1.0, 3, 5, 20), (2, 1000, 1.5, 3, 5, 20), (3, 513, 1.0, 1, 2, 12)]:
    gr = _grads_case(*args, seed=1)
    print("nsf bwd", args)
    for name, (got, want) in gr.items():
        want = want.double(); e = (got.double().cpu() - want).abs()
        if name.endswith("0") or name == "z" or name.endswith(".0"):
            print(f"   {name:10s} max|g| {float(want.abs().max()):.3e} err/max {float(e.max() / want.abs().max()):.2e}")
