"""Tensor-core NSF forward vs the CUDA-core kernel and the float64 oracle; quick timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import mentflow_b200 as mf
from mentflow_b200 import ops
from mfb_testutil import oracle_from_generator


def stats(name, a, b):
    e = ((a.double().cpu() - b.double().cpu()).abs() / b.double().cpu().abs().clamp_min(1.0)).flatten()
    print(f"  {name}: median {float(e.median()):.2e}  p99.9 {float(e.kthvalue(max(1, int(e.numel()*0.999))).values):.2e}  max {float(e.max()):.2e}  >1e-4: {int((e > 1e-4).sum())}/{e.numel()}")


for d, n, scale in [(6, 4096, 1.0), (6, 100003, 3.0), (2, 1000, 2.0), (4, 5000, 2.0), (3, 257, 1.0), (5, 333, 2.0)]:
    torch.manual_seed(d)
    gen = mf.generate.NSFGenerator(d)
    with torch.no_grad():
        for p in gen.parameters():
            p.mul_(scale)
    ref = oracle_from_generator(gen)
    gen = gen.to("cuda")
    z = torch.randn(n, d)
    z[: max(1, n // 100)] *= 4.0
    zc = z.cuda()
    with torch.no_grad():
        ops.NSF_USE_TENSOR_CORES = True
        x, lq = gen.forward_and_log_prob(zc)
        torch.cuda.synchronize()
        ops.NSF_USE_TENSOR_CORES = False
        x0, lq0 = gen.forward_and_log_prob(zc)
        xr, lr = ref.forward_and_log_prob(z.double())
    print(f"D={d} n={n} scale={scale}")
    stats("tc   x   vs f64", x, xr); stats("tc   logq vs f64", lq, lr)
    stats("cuda x   vs f64", x0, xr); stats("cuda logq vs f64", lq0, lr)

# timing, 6D, 1e6 particles
torch.manual_seed(0)
gen = mf.generate.NSFGenerator(6)
with torch.no_grad():
    for p in gen.parameters():
        p.mul_(3.0)
gen = gen.to("cuda")
z = torch.randn(1_000_000, 6, device="cuda")
for flag in (True, False):
    ops.NSF_USE_TENSOR_CORES = flag
    with torch.no_grad():
        for _ in range(3):
            gen.forward_and_log_prob(z)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gen.forward_and_log_prob(z)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"tensor_cores={flag}: {ms:.3f} ms per 1e6 particles -> {1e6 / ms * 1e3:.3e} particles/s")
