#!/bin/bash
# round 2, session AN: tensor-core inverse (density direction)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nsf.py -m gpu -x -q -k "inverse" 2>&1 | tail -15 | tee gpurun_out/r2an_tests.txt
cat > /tmp/inv_time.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, mentflow_b200 as mf
from mentflow_b200 import ops
torch.manual_seed(0)
gen = mf.generate.NSFGenerator(6)
with torch.no_grad():
    for p in gen.parameters(): p.mul_(1.5)
gen = gen.to("cuda")
n = 1_000_000
with torch.no_grad():
    x, lq = gen.sample_and_log_prob(n)
    for flag in (False, True):
        ops.NSF_INV_USE_TENSOR_CORES = flag
        for _ in range(2): lp = gen.log_prob(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): lp = gen.log_prob(x)
        b.record(); b.synchronize()
        ms = a.elapsed_time(b) / 3
        err = (lp - lq).abs()
        print("tensor cores" if flag else "cuda cores  ", "log_prob(x) of 1e6 particles: %.3f ms = %.3g particles/s; |log_prob(x) - log q| median %.2e, >1e-3: %d, max %.2e" % (ms, n / ms * 1e3, float(err.median()), int((err > 1e-3).sum()), float(err.max())))
PY
timeout 300 python /tmp/inv_time.py 2>&1 | tail -4 | tee gpurun_out/r2an_time.txt
