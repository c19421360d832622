"""CPU study of where the flow's fp32 error tail comes from (no GPU needed).

Evaluates the 5-layer flow three ways on the bench's x3 weights and compares with float64:
  torch32   the oracle restatement in fp32 (what the reference itself computes)
  direct    the CUDA kernels' spline formulation (bin width/height straight from the softmax terms,
            un-normalised running sums) in fp32 with IEEE exp/divide
  variants  of `direct` with single ingredients switched (see VARIANTS)
The conditioner is torch's fp32 F.linear in every fp32 arm, so differences are the spline's.
"""
import math
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch

from oracle.zuko_nsf import NSFOracle

B = 5.0
LOG2E = 1.4426950408889634
CW = 2.0 / abs(math.log(1e-3))
CD = 1.0 / abs(math.log(1e-3))


def noisy(x, ulps, gen):
    """x perturbed by a uniform relative error of +-ulps * 2^-24 (models an approximate MUFU result)."""
    if ulps == 0:
        return x
    return x * (1.0 + (torch.rand(x.shape, generator=gen, dtype=x.dtype) * 2 - 1) * ulps * 2.0 ** -24)


def spline_direct(phi, v, opt, gen):
    """phi (n, 59) raw conditioner output, v (n,). Returns y, log jac (natural)."""
    nb = 20
    dt = phi.dtype
    a = phi * LOG2E if opt.get("log2", True) else phi
    cw = CW / LOG2E if opt.get("log2", True) else CW
    cd = CD / LOG2E if opt.get("log2", True) else CD
    ex = (lambda t: torch.exp2(t)) if opt.get("log2", True) else (lambda t: torch.exp(t))

    def clip_exp(t, c):
        d = 1.0 + c * t.abs()
        if opt.get("quad_rcp", False):
            # one reciprocal per four parameters: 1/d_i = r * prod of the other three (extra roundings)
            n4 = (t.shape[-1] + 3) // 4 * 4
            dd = torch.nn.functional.pad(d, (0, n4 - t.shape[-1]), value=1.0).unflatten(-1, (-1, 4))
            p01, p23 = dd[..., 0] * dd[..., 1], dd[..., 2] * dd[..., 3]
            r = noisy(1.0 / (p01 * p23), opt.get("rcp_ulps", 1), gen)
            r01, r23 = r * p23, r * p01
            inv = torch.stack([r01 * dd[..., 1], r01 * dd[..., 0], r23 * dd[..., 3], r23 * dd[..., 2]], -1)
            inv = inv.flatten(-2)[..., : t.shape[-1]]
            arg = t * inv
        else:
            arg = t / d
        return noisy(ex(arg), opt.get("exp_ulps", 0), gen)

    ew = clip_exp(a[:, :nb], cw)
    eh = clip_exp(a[:, nb:2 * nb], cw)
    dpar = a[:, 2 * nb:]
    if opt.get("maxsub", False):
        pass
    # running sums in groups of four
    def prefix(e):
        g = e.unflatten(-1, (5, 4))
        gs = (g[..., 0] + g[..., 1]) + (g[..., 2] + g[..., 3])
        pre = torch.zeros(e.shape[0], 6, dtype=dt)
        for i in range(5):
            pre[:, i + 1] = pre[:, i] + gs[:, i]
        # knots: pre[g] + e0, + e1, + e2
        kn = torch.zeros(e.shape[0], 21, dtype=dt)
        for i in range(5):
            kn[:, 4 * i] = pre[:, i]
            c0 = pre[:, i] + g[:, i, 0]
            c1 = c0 + g[:, i, 1]
            c2 = c1 + g[:, i, 2]
            kn[:, 4 * i + 1], kn[:, 4 * i + 2], kn[:, 4 * i + 3] = c0, c1, c2
        kn[:, 20] = pre[:, 5]
        return kn, pre[:, 5]

    if opt.get("seqsum", False):
        def prefix(e):  # noqa: F811  plain sequential running sum
            kn = torch.zeros(e.shape[0], 21, dtype=dt)
            for j in range(20):
                kn[:, j + 1] = kn[:, j] + e[:, j]
            return kn, kn[:, 20]

    knx, sumw = prefix(ew)
    kny, sumh = prefix(eh)
    target = (v + B) * (0.5 / B) * sumw
    k = (knx[:, 1:20] < target[:, None]).sum(-1)
    idx = k[:, None]
    x0c = knx.gather(1, idx).squeeze(1)
    ek = ew.gather(1, idx).squeeze(1)
    y0c = kny.gather(1, idx).squeeze(1)
    hk = eh.gather(1, idx).squeeze(1)
    dpad = torch.nn.functional.pad(dpar, (1, 1), value=0.0)
    tl = dpad.gather(1, idx).squeeze(1)
    tr = dpad.gather(1, idx + 1).squeeze(1)
    d0 = noisy(ex(tl / (1.0 + cd * tl.abs())), opt.get("exp_ulps", 0), gen)
    d1 = noisy(ex(tr / (1.0 + cd * tr.abs())), opt.get("exp_ulps", 0), gen)
    if opt.get("knot_diff", False):
        # zuko style: normalised knots, width/height by differencing
        x0 = B * (2 * (x0c / sumw) - 1)
        x1c = knx.gather(1, idx + 1).squeeze(1)
        x1 = B * (2 * (x1c / sumw) - 1)
        y0 = B * (2 * (y0c / sumh) - 1)
        y1 = B * (2 * (kny.gather(1, idx + 1).squeeze(1) / sumh) - 1)
        s = (y1 - y0) / (x1 - x0)
        t = (v - x0) / (x1 - x0)
        dy = y1 - y0
    else:
        mode = opt.get("numer", "target")
        if mode == "target":
            t = (target - x0c) / ek
        elif mode == "fma":      # v*0.1*sum + (0.5 sum - x0c) with one rounding at the magnitude of the result
            c = 0.5 * sumw - x0c
            t = ((v * (0.5 / B)).double() * sumw.double() + c.double()).to(dt) / ek
        elif mode == "fma2":     # distance from the nearer end of the knot array
            suf = sumw - x0c if not opt.get("suffix") else None
            c = 0.5 * sumw - x0c
            t = ((v * (0.5 / B)).double() * sumw.double() + c.double()).to(dt) / ek
        elif mode == "norm":     # normalised left knot, v exact
            x0 = B * (2.0 * (x0c / sumw) - 1.0)
            t = (v - x0) / (2.0 * B * (ek / sumw))
        hn = hk / sumh
        s = hn * sumw / ek
        y0 = 2 * B * (y0c / sumh) - B
        dy = 2 * B * hn
    t = t.clamp(0.0, 1.0)
    omt = 1.0 - t
    tomt = t * omt
    den = (d0 + d1 - 2.0 * s) * tomt + s
    y = y0 + dy * (s * t * t + d0 * tomt) / den
    jac = s * s * (2.0 * s * tomt + d0 * omt * omt + d1 * t * t) / (den * den)
    inside = (v > -B) & (v <= B)
    return torch.where(inside, y, v), torch.where(inside, jac, torch.ones_like(jac))


def flow_direct(ref, z, opt, seed=0):
    gen = torch.Generator().manual_seed(seed)
    v = z
    dt = z.dtype
    n, D = z.shape
    total = torch.zeros(n, dtype=dt)
    for layer in ref.layers:
        phi = layer.hyper(v).unflatten(-1, (D, 59))
        y = torch.empty_like(v)
        jac = torch.ones(n, dtype=dt)
        lad = torch.zeros(n, dtype=dt)
        for f in range(D):
            yf, jf = spline_direct(phi[:, f], v[:, f], opt, gen)
            y[:, f] = yf
            jac = jac * jf
            lad = lad + jf.log()
        total = total + (jac.log() if opt.get("jacprod", True) else lad)
        v = y
    return v, ref.base_log_prob(z) - total


def stats(name, a, b):
    e = ((a.double() - b).abs() / b.abs().clamp_min(1.0)).flatten()
    q = lambda f: float(e.kthvalue(max(1, int(e.numel() * f))).values)
    print(f"  {name:26s} median {float(e.median()):.2e} p99 {q(0.99):.2e} p99.9 {q(0.999):.2e} "
          f"max {float(e.max()):.2e} >1e-4: {float((e > 1e-4).float().mean()) * 100:.3f}%", flush=True)


VARIANTS = {
    "direct exact": {},
    "direct natural-exp": {"log2": False},
    "direct seq-sum": {"seqsum": True},
    "direct knot-diff": {"knot_diff": True},
    "direct sum-of-logs": {"jacprod": False},
    "direct exp 2ulp": {"exp_ulps": 2},
    "direct quad-rcp": {"quad_rcp": True},
    "direct quad-rcp+exp 2ulp": {"quad_rcp": True, "exp_ulps": 2},
    "fma-numer": {"numer": "fma"},
    "norm-numer": {"numer": "norm"},
}

if __name__ == "__main__":
    D = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
    torch.manual_seed(0)
    ref32 = NSFOracle(D)
    with torch.no_grad():
        for p in ref32.parameters():
            p.mul_(scale)
    import copy
    ref64 = copy.deepcopy(ref32).double()
    torch.manual_seed(5)
    z = torch.randn(n, D)
    with torch.no_grad():
        xr, lr = ref64.forward_and_log_prob(z.double())
        x32, l32 = ref32.forward_and_log_prob(z)
        print(f"D={D} scale={scale} n={n}")
        stats("torch32 x", x32, xr)
        stats("torch32 logq", l32, lr)
        # the direct formulation in float64: the formulation itself must agree with the oracle
        x, lq = flow_direct(ref64, z.double(), {})
        stats("direct f64 x", x, xr)
        stats("direct f64 logq", lq, lr)
        for name, opt in VARIANTS.items():
            x, lq = flow_direct(ref32, z, opt)
            stats(name + " x", x, xr)
            stats(name + " logq", lq, lr)
