import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import mentflow_b200 as mf
from mentflow_b200 import ops
from mfb_testutil import oracle_from_generator

def run(d, passes, n, tails, bwd_tc=True, scale=2.0):
    torch.manual_seed(70 + d + (passes or 0))
    gen = mf.generate.build_generator("nsf", input_features=d, output_features=d, hidden_layers=3, hidden_units=64,
                                      transforms=5, bins=20, passes=passes)
    with torch.no_grad():
        for p in gen.parameters():
            p.mul_(scale)
    ref, ref32 = oracle_from_generator(gen), oracle_from_generator(gen, torch.float32)
    gen = gen.to("cuda")
    z = torch.randn(n, d)
    if tails:
        z[: max(1, n // 100)] *= 4.0
    a, b = torch.randn(n, d), torch.randn(n)
    ops.NSF_BWD_USE_TENSOR_CORES = bwd_tc
    zc = z.clone().cuda().requires_grad_(True)
    xg, lg = gen.forward_and_log_prob(zc)
    ((xg * a.cuda()).sum() + (lg * b.cuda()).sum()).backward()
    def og(flow, dtype):
        zz = z.to(dtype).clone().requires_grad_(True)
        xo, lo = flow.forward_and_log_prob(zz)
        ((xo * a.to(dtype)).sum() + (lo * b.to(dtype)).sum()).backward()
        return {"b_in": torch.stack([flow.layers[t].hyper[0].bias.grad for t in range(5)]).double(),
                "w_out": torch.stack([flow.layers[t].hyper[6].weight.grad * flow.layers[t].hyper[6].mask for t in range(5)]).double(),
                "b_out": torch.stack([flow.layers[t].hyper[6].bias.grad for t in range(5)]).double()}
    want, t32 = og(ref, torch.float64), og(ref32, torch.float32)
    got = {"b_in": gen.b_in.grad, "w_out": gen.w_out.grad, "b_out": gen.b_out.grad}
    out = []
    for name in got:
        sc = float(want[name].abs().max())
        e = (got[name].cpu().double() - want[name]).abs() / sc
        per_layer = e.reshape(5, -1).max(dim=1).values
        out.append(f"{name}: mine {float(e.max()):.1e} t32 {float((t32[name]-want[name]).abs().max())/sc:.1e} per-layer {[f'{float(v):.0e}' for v in per_layer]}")
    print(f"d={d} passes={passes} n={n} tails={tails} bwd_tc={bwd_tc}: " + " | ".join(out), flush=True)

run(6, 2, 20000, True)
run(6, 2, 20000, False)
run(6, None, 20000, True, bwd_tc=False)
run(6, None, 20000, True, bwd_tc=True)
run(6, 2, 20000, True, scale=1.0)
