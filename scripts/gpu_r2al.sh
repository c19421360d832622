#!/bin/bash
# round 2, session AL: KDE-2D backward with four particles per thread per table load
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kde2d.py tests/test_gpu_baseline_sized.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2al_tests.txt
cat > /tmp/k2b.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch
from mentflow_b200 import ops
from mfb_testutil import geom_rows
gen = torch.Generator().manual_seed(9)
n, d, k, bx, by = 1_000_000, 6, 15, 85, 85
x = (torch.randn(n, d, generator=gen) * 0.8).cuda().requires_grad_(True)
w = torch.randn(k, 2, d, generator=gen); w = (w / w.norm(dim=2, keepdim=True)).cuda()
ex, ey = torch.linspace(-3.5, 3.5, bx + 1), torch.linspace(-3.5, 3.5, by + 1)
gx_, sx = geom_rows(ex, 0.5, k); gy_, sy = geom_rows(ey, 0.5, k)
geom = torch.stack([gx_, gy_], dim=1).cuda()
prof = ops.ProjectKDE2D.apply(x, w, geom, 0.5, bx, by, None) if hasattr(ops, "ProjectKDE2D") else None
gp = torch.randn_like(prof)
for _ in range(2): (prof * gp).sum().backward(retain_graph=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): (prof * gp).sum().backward(retain_graph=True)
b.record(); b.synchronize()
print("kde2d backward (15 screens 85x85, 1e6 particles): %.3f ms" % (a.elapsed_time(b) / 5))
PY
timeout 200 python /tmp/k2b.py 2>&1 | tail -3 | tee gpurun_out/r2al_time.txt
