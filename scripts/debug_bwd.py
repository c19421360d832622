import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_nsf import _grads_case
for args, seed in [((6, 3000, 1.0, 3, 5, 20), 3006), ((2, 1000, 1.5, 3, 5, 20), 1002), ((2, 1000, 1.5, 3, 5, 20), 1)]:
    gr = _grads_case(*args, seed=seed)
    print("nsf bwd", args, seed)
    for name, (got, want) in gr.items():
        want = want.double(); e = (got.double().cpu() - want).abs() / want.abs().max()
        print(f"   {name:10s} max|g| {float(want.abs().max()):.3e} max {float(e.max()):.2e} median {float(e.flatten().median()):.2e} n>2e-3 {int((e>2e-3).sum())}/{e.numel()}")
