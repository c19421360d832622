"""Small end-to-end case for compute-sanitizer: tensor-core flow (D=6 and D=2), KDE-1D, loss."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mentflow_b200 as mf
from mentflow_b200 import workloads
dev = torch.device("cuda")
for d, n in ((6, 3001), (2, 777)):
    torch.manual_seed(d)
    gen = mf.generate.NSFGenerator(d).to(dev)
    z = torch.randn(n, d, device=dev)
    with torch.no_grad():
        x, lq = gen.forward_and_log_prob(z)
    wl = workloads.isotropic_1d(6, 10, 64, 3.5) if d == 6 else workloads.rotations_2d(7, 85, 3.5)
    tfs = [mf.simulate.LinearTransform(m.to(dev)) for m in wl["matrices"]]
    diag = mf.diagnostics.Histogram1D(axis=0, edges=wl["edges"], bandwidth=0.5).to(dev)
    with torch.no_grad():
        prof = mf.simulate.forward(x, tfs, [[diag] for _ in tfs])
    torch.cuda.synchronize()
    print(d, float(x.abs().mean()), float(lq.mean()), float(prof[0][0].sum()))
print("done")
