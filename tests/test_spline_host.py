"""The tcgen05 kernels' register spline (mentflow_b200/csrc/nsf_spline_regs.cuh), compiled for the HOST
(tests/csrc/spline_host.cu) and checked against the float64 oracle -- no GPU needed.

This is the arithmetic the GPU epilogue runs (summation order, centred differences, fused
multiply-adds, two-level bin search); only ex2.approx / rcp.approx are replaced by libm.  The
flow-level test states the parity bar of the flow: on the benchmark's x3 weights the number of log q
entries beyond rel 1e-4 of float64 is at most 1.25x what the reference's own fp32 evaluation
(torch-fp32 restatement of zuko) shows on the same input.
"""
import copy
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle.zuko_nsf import NSFOracle, RQSpline

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "spline_host.cu")
HDR = os.path.join(os.path.dirname(HERE), "mentflow_b200", "csrc", "nsf_spline_regs.cuh")
TOL = 1.0e-4


@pytest.fixture(scope="module")
def lib():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = os.path.join(HERE, "_build", "libspline_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Xcompiler",
                        "-ffp-contract=off", "-diag-suppress", "549", SRC, "-o", out], check=True)
    L = ctypes.CDLL(out)
    fp = ctypes.POINTER(ctypes.c_float)
    L.spline_host_fwd.argtypes = [fp, fp, ctypes.c_int64, fp, fp]
    L.spline_host_bwd.argtypes = [fp, fp, fp, fp, ctypes.c_int64, fp, fp]
    L.spline_host_inv.argtypes = [fp, fp, ctypes.c_int64, fp, fp]
    return L


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def host_spline(lib, phi, v):
    """phi (n, 59) float32 tensor, v (n,) -> y, jac (float32 tensors)."""
    n = phi.shape[0]
    a = np.zeros((n, 64), dtype=np.float32)
    a[:, :59] = phi.numpy()
    vv = np.ascontiguousarray(v.numpy(), dtype=np.float32)
    y = np.empty(n, dtype=np.float32)
    jac = np.empty(n, dtype=np.float32)
    lib.spline_host_fwd(_ptr(a), _ptr(vv), n, _ptr(y), _ptr(jac))
    return torch.from_numpy(y), torch.from_numpy(jac)


def host_flow(lib, ref32, z):
    """The flow with torch's fp32 conditioner and the kernel's spline arithmetic."""
    v = z
    n, d = z.shape
    total = torch.zeros(n)
    for layer in ref32.layers:
        phi = layer.hyper(v).unflatten(-1, (d, 59))
        y = torch.empty_like(v)
        jac = torch.ones(n)
        for f in range(d):
            yf, jf = host_spline(lib, phi[:, f].contiguous(), v[:, f].contiguous())
            y[:, f] = yf
            jac = jac * jf
        total = total + jac.log()
        v = y
    return v, ref32.base_log_prob(z) - total


def rel(a, b):
    return ((a.double() - b.double()).abs() / b.double().abs().clamp_min(1.0)).flatten()


@pytest.mark.parametrize("scale", [1.0, 3.0, 6.0])
def test_single_spline_matches_float64(lib, scale):
    torch.manual_seed(int(scale))
    n = 20000
    phi = torch.randn(n, 59) * scale
    v = torch.randn(n) * 2.5
    v[:200] = torch.tensor([-5.0, 5.0, -5.0000005, 5.0000005] * 50)      # the edges of the box
    v[200:300] *= 3.0                                                      # identity tails
    y, jac = host_spline(lib, phi, v)
    w, h, d = phi.double().split([20, 20, 19], dim=-1)
    yr, ladj = RQSpline(w, h, d).call_and_ladj(v.double())
    inside = (v > -5.0) & (v <= 5.0)
    assert torch.equal(y[~inside], v[~inside]) and bool((jac[~inside] == 1.0).all())
    ey = rel(y, yr)
    el = rel(jac.double().log(), ladj)
    # one spline, random parameters: bins as narrow as 1e-5 of the box make t ill-conditioned for any fp32
    # evaluation; the bulk must sit at rounding level and the tail must be thin
    assert float(ey.median()) < 2e-7 and float(el.median()) < 2e-6
    assert float((ey > TOL).float().mean()) < 2e-3 and float((el > TOL).float().mean()) < 2e-2


@pytest.mark.parametrize("d,scale", [(6, 3.0), (4, 3.0), (2, 3.0), (6, 1.0)])
def test_flow_tail_no_worse_than_reference_fp32(lib, d, scale):
    torch.manual_seed(0)
    ref32 = NSFOracle(d)
    with torch.no_grad():
        for p in ref32.parameters():
            p.mul_(scale)
    ref64 = copy.deepcopy(ref32).double()
    torch.manual_seed(5)
    z = torch.randn(40000, d)
    with torch.no_grad():
        xr, lr = ref64.forward_and_log_prob(z.double())
        x32, l32 = ref32.forward_and_log_prob(z)
        x, lq = host_flow(lib, ref32, z)
    for got, t32, truth in ((x, x32, xr), (lq, l32, lr)):
        e, e32 = rel(got, truth), rel(t32, truth)
        bad, bad32 = int((e > TOL).sum()), int((e32 > TOL).sum())
        assert bad <= 1.25 * bad32 + 2, f"{bad} entries beyond {TOL}, torch-fp32 {bad32}"
        assert float(e.median()) <= 1.25 * float(e32.median()) + 1e-7
        assert float(e.max()) <= max(TOL, 2.5 * float(e32.max()))


def test_spline_backward_matches_autograd(lib):
    torch.manual_seed(3)
    n = 4000
    phi = (torch.randn(n, 59) * 2.0)
    v = torch.randn(n) * 2.0
    v[:50] *= 4.0
    gy = torch.randn(n)
    gl = torch.randn(n)
    a = np.zeros((n, 64), dtype=np.float32)
    a[:, :59] = phi.numpy()
    gphi = np.empty((n, 64), dtype=np.float32)
    gv = np.empty(n, dtype=np.float32)
    lib.spline_host_bwd(_ptr(a), _ptr(v.numpy()), _ptr(gy.numpy()), _ptr(gl.numpy()), n, _ptr(gphi), _ptr(gv))
    p64 = phi.double().requires_grad_(True)
    v64 = v.double().requires_grad_(True)
    w, h, d = p64.split([20, 20, 19], dim=-1)
    y, ladj = RQSpline(w, h, d).call_and_ladj(v64)
    (y * gy.double() + ladj * gl.double()).sum().backward()
    gp = torch.from_numpy(gphi[:, :59]).double()
    scale = p64.grad.abs().max(dim=1, keepdim=True).values.clamp_min(1e-3)
    e = ((gp - p64.grad).abs() / scale).flatten()
    ev = (torch.from_numpy(gv).double() - v64.grad).abs() / v64.grad.abs().clamp_min(1.0)
    assert float(e.median()) < 1e-6 and float((e > 1e-3).float().mean()) < 2e-3
    assert float(ev.median()) < 1e-6 and float((ev > 1e-3).float().mean()) < 5e-3


@pytest.mark.parametrize("scale", [1.0, 3.0])
def test_register_spline_inverse_matches_float64(lib, scale):
    """rq_spline_regs_inv (the density direction of the tensor-core kernels) against zuko's MonotonicRQSTransform in
    float64: v = RQS^-1(y) and the forward Jacobian at v; also the round trip through the forward spline."""
    torch.manual_seed(7)
    n = 20000
    phi = torch.randn(n, 59) * scale
    y = (torch.rand(n) - 0.5) * 12.0                       # inside and outside the box
    a = np.zeros((n, 64), dtype=np.float32)
    a[:, :59] = phi.numpy()
    yy = np.ascontiguousarray(y.numpy(), dtype=np.float32)
    v = np.empty(n, dtype=np.float32)
    jac = np.empty(n, dtype=np.float32)
    lib.spline_host_inv(_ptr(a), _ptr(yy), n, _ptr(v), _ptr(jac))
    v, jac = torch.from_numpy(v), torch.from_numpy(jac)
    p64 = phi.double()
    sp = RQSpline(p64[:, :20], p64[:, 20:40], p64[:, 40:])
    v64 = sp.inverse(y.double())
    _, ladj64 = sp.call_and_ladj(v64)
    ev = rel(v, v64)
    el = rel(jac.double().log(), ladj64)
    assert float(ev.median()) < 1e-6 and float(el.median()) < 1e-6
    # isolated ill-conditioned bins (tiny widths next to large heights) leave a tail, as in the forward direction
    assert int((ev > TOL).sum()) <= n // 500 and int((el > TOL).sum()) <= n // 200, (int((ev > TOL).sum()), int((el > TOL).sum()))
    back, _ = host_spline(lib, phi, v)
    eb = rel(back, y)
    assert float(eb.median()) < 1e-6 and int((eb > TOL).sum()) <= n // 500
