"""One KDE-1D deposit at the bench size (for ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mentflow_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from mfb_testutil import geom_rows

n, d, k, nb = 1_000_000, 6, 100, 64
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, d, generator=gen, device="cuda")
w = torch.randn(k, d, generator=gen, device="cuda")
w = w / w.norm(dim=1, keepdim=True)
geom = geom_rows(torch.linspace(-3.5, 3.5, nb + 1), 0.5, k)[0].cuda()
for _ in range(3):
    ops.kde1d_sums(x, w, geom, 0.5, nb)
torch.cuda.synchronize()
