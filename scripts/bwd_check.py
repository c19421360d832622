"""Backward of the flow with the tcgen05 data-gradient chain vs the CUDA-core chain; timing."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
import mentflow_b200 as mf
from mentflow_b200 import _lib, ops
lib = _lib.load()

def grads(gen, z, a, b, flag, scale=1.0):
    ops.NSF_BWD_USE_TENSOR_CORES = bool(flag)
    for p in gen.parameters(): p.grad = None
    zc = z.clone().requires_grad_(True)
    x, lq = gen.forward_and_log_prob(zc)
    (((x * a).sum() + (lq * b).sum()) * scale).backward()
    return [zc.grad.clone()] + [p.grad.clone() for p in gen.parameters()]

for d, n, scale in [(6, 3000, 1.0), (6, 40000, 1e-6), (2, 1000, 1.0), (4, 777, 1e-3)]:
    torch.manual_seed(d + n)
    gen = mf.generate.NSFGenerator(d)
    with torch.no_grad():
        for p in gen.parameters(): p.mul_(1.5)
    gen = gen.to("cuda")
    z = torch.randn(n, d, device="cuda"); a = torch.randn(n, d, device="cuda"); b = torch.randn(n, device="cuda")
    g1 = grads(gen, z, a, b, True, scale); g0 = grads(gen, z, a, b, False, scale)
    errs = [float(((x - y).abs().max() / y.abs().max().clamp_min(1e-30))) for x, y in zip(g1, g0)]
    print(f"D={d} n={n} loss scale {scale}: max rel diff tc vs cuda-core per tensor: {['%.1e' % e for e in errs]}")

torch.manual_seed(0)
gen = mf.generate.NSFGenerator(6)
with torch.no_grad():
    for p in gen.parameters(): p.mul_(3.0)
gen = gen.to("cuda")
n = 1_000_000
z = torch.randn(n, 6, device="cuda"); a = torch.randn(n, 6, device="cuda"); b = torch.randn(n, device="cuda")
for flag in (True, False):
    for _ in range(2): grads(gen, z, a, b, flag)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): grads(gen, z, a, b, flag)
    e1.record(); e1.synchronize()
    print(f"tensor-core dgrad={flag}: forward+backward {e0.elapsed_time(e1) / 3:.2f} ms per 1e6 particles")
