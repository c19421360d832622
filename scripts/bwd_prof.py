import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mentflow_b200 as mf
torch.manual_seed(0)
gen = mf.generate.NSFGenerator(6)
with torch.no_grad():
    for p in gen.parameters(): p.mul_(3.0)
gen = gen.to("cuda")
n = 1_000_000
z = torch.randn(n, 6, device="cuda"); a = torch.randn(n, 6, device="cuda"); b = torch.randn(n, device="cuda")
for _ in range(2):
    for p in gen.parameters(): p.grad = None
    x, lq = gen.forward_and_log_prob(z.clone().requires_grad_(True))
    ((x * a).sum() + (lq * b).sum()).backward()
torch.cuda.synchronize()
