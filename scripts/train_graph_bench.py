import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, time
import bench
from mentflow_b200.graphs import GraphedTrainStep
class A: ndim=6; num_proj=100; bins=64
dev = torch.device("cuda")
model, wl = bench.build_model(A, dev)
for n in (25_000, 100_000):
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.0, capturable=True)
    def eager():
        opt.zero_grad(set_to_none=True)
        L, H, D = model.loss(n); L.backward(); opt.step()
    for _ in range(3): eager()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): eager()
    torch.cuda.synchronize(); te = (time.perf_counter() - t0) / 20
    g = GraphedTrainStep(model, opt, n)
    for _ in range(3): g()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): g()
    torch.cuda.synchronize(); tg = (time.perf_counter() - t0) / 20
    print(f"n={n}: eager train step {te*1e3:.3f} ms, graph replay {tg*1e3:.3f} ms  ({n/tg:.3e} particles/s)")
