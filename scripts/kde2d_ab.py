"""A/B of the two KDE-2D forward paths (tcgen05 GEMM vs windowed fixed-point deposits): agreement by
magnitude decade, against a float64 dense evaluation on a subsample, and time at the C4 size."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))
import torch

from mentflow_b200 import _lib, ops
from mfb_testutil import geom_rows

lib = _lib.load()
gen = torch.Generator().manual_seed(9)
n, d, k, bx, by = 60_000, 6, 7, 85, 85
x = (torch.randn(n, d, generator=gen) * 0.8).cuda()
w = torch.randn(k, 2, d, generator=gen)
w = (w / w.norm(dim=2, keepdim=True)).cuda()
ex, ey = torch.linspace(-3.5, 3.5, bx + 1), torch.linspace(-3.5, 3.5, by + 1)
gx, sx = geom_rows(ex, 0.5, k)
gy, sy = geom_rows(ey, 0.5, k)
geom = torch.stack([gx, gy], dim=1).cuda()
ops.KDE2D_USE_TENSOR_CORES = True
tc = ops.kde2d_sums(x, w, geom, 0.5, bx, by)[0].double()
ops.KDE2D_USE_TENSOR_CORES = False
fx = ops.kde2d_sums(x, w, geom, 0.5, bx, by)[0].double()
# float64 dense reference on the GPU
cx = (0.5 * (ex[1:] + ex[:-1])).double().cuda()
cy = (0.5 * (ey[1:] + ey[:-1])).double().cuda()
xd = x.double()
ref = torch.zeros(k, bx, by, dtype=torch.float64, device="cuda")
for i in range(k):
    ux, uy = xd @ w[i, 0].double(), xd @ w[i, 1].double()
    kx = torch.exp(-0.5 * ((ux[:, None] - cx[None]) / float(sx)) ** 2)
    ky = torch.exp(-0.5 * ((uy[:, None] - cy[None]) / float(sy)) ** 2)
    ref[i] = kx.T @ ky
peak = float(ref.max())
for lo, hi in [(1e-3, 10), (1e-6, 1e-3), (1e-9, 1e-6), (1e-12, 1e-9), (1e-16, 1e-12)]:
    m = (ref > lo * peak) & (ref <= hi * peak)
    if int(m.sum()) == 0:
        continue
    print(f"bins in ({lo:g}, {hi:g}] x peak: {int(m.sum()):6d}   max rel err  tensor-core {float(((tc - ref).abs() / ref)[m].max()):.2e}"
          f"   fixed-point {float(((fx - ref).abs() / ref)[m].max()):.2e}")

n = 1_000_000
k = 15
x = torch.randn(n, d, device="cuda")
w = torch.randn(k, 2, d, device="cuda")
w = w / w.norm(dim=2, keepdim=True)
geom = torch.stack([geom_rows(ex, 0.5, k)[0], geom_rows(ey, 0.5, k)[0]], dim=1).cuda()
for flag in (1, 0):
    ops.KDE2D_USE_TENSOR_CORES = bool(flag)
    for _ in range(2):
        ops.kde2d_sums(x, w, geom, 0.5, bx, by)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ops.kde2d_sums(x, w, geom, 0.5, bx, by)
    b.record()
    b.synchronize()
    print("tensor cores" if flag else "fixed point ", f"{a.elapsed_time(b) / 5:.3f} ms per 1e6 particles x 15 screens 85x85")
ops.KDE2D_USE_TENSOR_CORES = True
