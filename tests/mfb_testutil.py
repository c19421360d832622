"""Shared helpers for the test-suite (imported as a top-level module; tests/ has no __init__)."""
import numpy as np
import torch


def t32(a):
    return torch.from_numpy(np.asarray(a)).to(torch.float32)


def cuda(a):
    return t32(a).cuda()


def oracle_from_generator(gen, dtype=torch.float64):
    """NSFOracle carrying the weights of a mentflow_b200 NSFGenerator (via zuko-style names)."""
    from oracle.zuko_nsf import NSFOracle
    flow = NSFOracle(gen.features, gen.hidden_units, gen.hidden_layers, gen.transforms, gen.bins,
                     passes=getattr(gen, "passes", None))
    sd = gen.state_dict()
    new = {}
    for k, v in flow.state_dict().items():
        # layers.{t}.hyper.{i}.{weight,bias,mask} / layers.{t}.order
        parts = k.split(".")
        src = "_flow.transform.transforms." + ".".join(parts[1:])
        new[k] = sd[src].detach().cpu().to(v.dtype)
    flow.load_state_dict(new)
    return flow.to(dtype)


def generator_from_golden(g, device="cpu"):
    """NSFGenerator loaded with the weights stored in tests/golden/nsf_*.npz."""
    import mentflow_b200 as mf
    d = g["z"].shape[1]
    gen = mf.generate.NSFGenerator(d)
    sd = {}
    for k, v in g.items():
        if k.startswith("sd:") and (k.endswith("weight") or k.endswith("bias")):
            sd["_flow.transform.transforms." + k[len("sd:layers."):]] = torch.from_numpy(v).float()
    gen.load_state_dict(sd)
    return gen.to(device)


def rel_err(a, b):
    """max |a-b| / max(1, |b|) elementwise (SURVEY 8c metric for x and log q)."""
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() / b.abs().clamp_min(1.0)).max())


def profile_err(a, b):
    """max over profiles of max|a-b| / max|b| (SURVEY 8c metric for KDE profiles)."""
    a, b = a.double().cpu(), b.double().cpu()
    a, b = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
    return float(((a - b).abs().max(dim=1).values / b.abs().max(dim=1).values).max())


def geom_rows(edges, bandwidth, k):
    """k identical C-ABI geometry records [c0, spacing, sigma, delta, 0...] for fp32 edges."""
    e = edges.float()
    c = 0.5 * (e[:-1] + e[1:])
    delta = float(c[1] - c[0])
    spacing = float((c[-1].double() - c[0].double()) / (c.numel() - 1))
    sigma = float(bandwidth * (e[1] - e[0]))
    return torch.tensor([[float(c[0]), spacing, sigma, delta, 0, 0, 0, 0]] * k, dtype=torch.float32), sigma


def err_stats(a, b):
    """elementwise |a-b| / max(1,|b|): (max, 99.9th percentile)."""
    e = ((a.double().cpu() - b.double().cpu()).abs() / b.double().cpu().abs().clamp_min(1.0)).flatten()
    k = max(1, int(e.numel() * 0.999))
    return float(e.max()), float(e.kthvalue(k).values)
