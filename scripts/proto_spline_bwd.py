"""Prototype (float64, scalar loops) of the spline backward formulas used by the CUDA kernel,
checked against autograd of the oracle's RQSpline."""
import math, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.zuko_nsf import RQSpline
B = 5.0
LS = abs(math.log(1e-3)); CW = 2.0 / LS; CD = 1.0 / LS

def spline_fwd_bwd(raw, v, gy, gl, nb):
    """raw: list of 3nb-1 floats. returns y, ladj, graw(list), gv  for L = gy*y + gl*ladj"""
    w = [r / (1 + CW * abs(r)) for r in raw[:nb]]
    h = [r / (1 + CW * abs(r)) for r in raw[nb:2 * nb]]
    mw, mh = max(w), max(h)
    ew = [math.exp(a - mw) for a in w]; eh = [math.exp(a - mh) for a in h]
    sw, sh = sum(ew), sum(eh)
    W = [e / sw for e in ew]; H = [e / sh for e in eh]
    # search
    cum = 0.0; xl = -B; k = -1
    for j in range(nb):
        cum += W[j]; xr = 2 * B * cum - B
        if k < 0 and xl < v <= xr:
            k = j; x0 = xl
        xl = xr
    graw = [0.0] * (3 * nb - 1)
    if k < 0:
        return v, 0.0, graw, gy
    y0 = 2 * B * sum(H[:k]) - B
    wk, hk = W[k], H[k]
    def dval(idx):
        r = raw[2 * nb + idx]; c = r / (1 + CD * abs(r)); return math.exp(c), 1.0 / (1 + CD * abs(r)) ** 2
    d0, dd0 = (1.0, 0.0) if k == 0 else dval(k - 1)
    d1, dd1 = (1.0, 0.0) if k == nb - 1 else dval(k)
    dx, dy = 2 * B * wk, 2 * B * hk
    s = hk / wk
    t = (v - x0) / dx
    omt = 1 - t; q = t * omt
    A = d0 + d1 - 2 * s
    den = s + A * q
    N1 = s * t * t + d0 * q
    N2 = 2 * s * q + d0 * omt * omt + d1 * t * t
    y = y0 + dy * N1 / den
    ladj = 2 * math.log(s) + math.log(N2) - 2 * math.log(den)
    # ---- partials w.r.t. (t, s, d0, d1, y0, dy)
    dq_dt = 1 - 2 * t
    dden_dt = A * dq_dt; dden_ds = 1 - 2 * q; dden_dd = q
    dN1_dt = 2 * s * t + d0 * dq_dt; dN1_ds = t * t; dN1_dd0 = q
    dN2_dt = 2 * s * dq_dt - 2 * d0 * omt + 2 * d1 * t; dN2_ds = 2 * q; dN2_dd0 = omt * omt; dN2_dd1 = t * t
    # y = y0 + dy*N1/den
    r = N1 / den
    dy_dt = dy * (dN1_dt - r * dden_dt) / den
    dy_ds = dy * (dN1_ds - r * dden_ds) / den
    dy_dd0 = dy * (dN1_dd0 - r * dden_dd) / den
    dy_dd1 = dy * (-r * dden_dd) / den
    dy_ddy = r
    # ladj
    dl_dt = dN2_dt / N2 - 2 * dden_dt / den
    dl_ds = 2 / s + dN2_ds / N2 - 2 * dden_ds / den
    dl_dd0 = dN2_dd0 / N2 - 2 * dden_dd / den
    dl_dd1 = dN2_dd1 / N2 - 2 * dden_dd / den
    g_t = gy * dy_dt + gl * dl_dt
    g_s = gy * dy_ds + gl * dl_ds
    g_d0 = gy * dy_dd0 + gl * dl_dd0
    g_d1 = gy * dy_dd1 + gl * dl_dd1
    g_y0 = gy
    g_dy = gy * dy_ddy
    # t = (v - x0)/dx ; s = hk/wk ; dx = 2B wk ; dy = 2B hk ; x0 = 2B sum_{j<k} W_j - B ; y0 likewise
    gv = g_t / dx
    g_x0 = -g_t / dx
    g_dx = -g_t * t / dx
    gW = [0.0] * nb; gH = [0.0] * nb
    for j in range(k):
        gW[j] = 2 * B * g_x0
        gH[j] = 2 * B * g_y0
    gW[k] = 2 * B * g_dx - g_s * s / wk
    gH[k] = 2 * B * g_dy + g_s / wk
    # softmax backward + soft clip
    dotW = sum(gW[j] * W[j] for j in range(nb)); dotH = sum(gH[j] * H[j] for j in range(nb))
    for j in range(nb):
        graw[j] = W[j] * (gW[j] - dotW) / (1 + CW * abs(raw[j])) ** 2
        graw[nb + j] = H[j] * (gH[j] - dotH) / (1 + CW * abs(raw[nb + j])) ** 2
    if k > 0: graw[2 * nb + k - 1] = g_d0 * d0 * dd0
    if k < nb - 1: graw[2 * nb + k] = g_d1 * d1 * dd1
    return y, ladj, graw, gv

torch.manual_seed(0)
nb = 20
worst = 0
for trial in range(300):
    raw = (torch.randn(3 * nb - 1, dtype=torch.float64) * 3).requires_grad_(True)
    v = (torch.randn((), dtype=torch.float64) * 3).requires_grad_(True)
    if trial % 50 == 0: v = (v.detach() * 3).requires_grad_(True)
    gy, gl = float(torch.randn(())), float(torch.randn(()))
    sp = RQSpline(raw[None, :nb], raw[None, nb:2 * nb], raw[None, 2 * nb:])
    y, l = sp.call_and_ladj(v[None])
    (gy * y + gl * l).sum().backward()
    y2, l2, graw, gv = spline_fwd_bwd(raw.detach().tolist(), float(v), gy, gl, nb)
    e = max(abs(y2 - float(y)), abs(l2 - float(l)), abs(gv - float(v.grad)) / (1 + abs(float(v.grad))),
            float((torch.tensor(graw) - raw.grad).abs().max() / (1 + raw.grad.abs().max())))
    worst = max(worst, e)
print("worst mismatch vs autograd:", worst)

# locate the worst component
torch.manual_seed(0)
for trial in range(300):
    raw = (torch.randn(3 * nb - 1, dtype=torch.float64) * 3).requires_grad_(True)
    v = (torch.randn((), dtype=torch.float64) * 3).requires_grad_(True)
    if trial % 50 == 0: v = (v.detach() * 3).requires_grad_(True)
    gy, gl = float(torch.randn(())), float(torch.randn(()))
    sp = RQSpline(raw[None, :nb], raw[None, nb:2 * nb], raw[None, 2 * nb:])
    y, l = sp.call_and_ladj(v[None])
    (gy * y + gl * l).sum().backward()
    y2, l2, graw, gv = spline_fwd_bwd(raw.detach().tolist(), float(v.detach()), gy, gl, nb)
    errs = [abs(y2 - float(y.detach())), abs(l2 - float(l.detach())), abs(gv - float(v.grad)) / (1 + abs(float(v.grad))),
            float((torch.tensor(graw) - raw.grad).abs().max() / (1 + raw.grad.abs().max()))]
    if max(errs) > 1e-10:
        print(trial, errs, float(v.detach()))
