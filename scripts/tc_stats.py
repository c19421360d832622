"""Error statistics vs the float64 oracle: tcgen05 kernel, CUDA-core kernel, torch-fp32 (the reference's own
arithmetic), on the golden weights and on the benchmark's weights (default init x3, seed 0)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
import mentflow_b200 as mf
from mentflow_b200 import ops
from mfb_testutil import generator_from_golden, oracle_from_generator

def err(a, b):
    return ((a.double().cpu() - b.double().cpu()).abs() / b.double().cpu().abs().clamp_min(1.0)).flatten()

def stats(name, a, b, base=None):
    e = err(a, b)
    q = lambda f: float(e.kthvalue(max(1, int(e.numel() * f))).values)
    bad = int((e > 1e-4).sum())
    rel = "" if base is None else f" ({bad / max(base, 1):.2f}x torch32)"
    print(f"  {name:18s} median {float(e.median()):.2e} p99 {q(0.99):.2e} p99.9 {q(0.999):.2e} max {float(e.max()):.2e} >1e-4: {bad}{rel}", flush=True)
    return bad

def run(tag, gen, d):
    ref64, ref32 = oracle_from_generator(gen), oracle_from_generator(gen, torch.float32)
    torch.manual_seed(5)
    z = torch.randn(100_000, d)
    with torch.no_grad():
        xr, lr = ref64.forward_and_log_prob(z.double())
        x32, l32 = ref32.forward_and_log_prob(z)
        print(f"{tag} D={d}")
        bx = stats("torch32 x", x32, xr); bl = stats("torch32 logq", l32, lr)
        for flag, nm in ((True, "tcgen05"), (False, "cuda-core")):
            ops.NSF_USE_TENSOR_CORES = flag
            x, lq = gen.forward_and_log_prob(z.cuda())
            stats(nm + " x", x, xr, bx); stats(nm + " logq", lq, lr, bl)

for d in (2, 6):
    g = dict(np.load(os.path.join(R, "tests", "golden", f"nsf_{d}d.npz")))
    run("golden", generator_from_golden(g, "cuda"), d)
for d in (2, 4, 6):
    torch.manual_seed(0)
    gen = mf.generate.NSFGenerator(d)
    with torch.no_grad():
        for p in gen.parameters():
            p.mul_(3.0)
    run("bench x3", gen.to("cuda"), d)
