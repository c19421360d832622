import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mentflow_b200.graphs import GraphedLoss
class A: ndim=6; num_proj=100; bins=64
dev = torch.device("cuda")
model, wl = bench.build_model(A, dev)
n = 1_000_000
z_host = torch.randn(n, 6).pin_memory()
for chunks in (1, 2, 3, 4, 8):
    g = GraphedLoss(model, n, host_chunks=chunks)
    ncap = [0]
    orig = g._capture_host
    def cap(z, orig=orig): ncap[0] += 1; orig(z)
    g._capture_host = cap
    for _ in range(3): g(z_host)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        L = g(z_host)[0]; v = float(L.item())
    dt = (time.perf_counter() - t0) / 10
    print(f"chunks={chunks}: {dt*1e3:.3f} ms per e2e step, captures={ncap[0]}")
# plain copy bandwidth
zd = torch.empty(n, 6, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): zd.copy_(z_host, non_blocking=True)
torch.cuda.synchronize(); print("H2D 24MB:", (time.perf_counter()-t0)/10*1e3, "ms")
