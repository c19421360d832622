"""Thin torch layer over the C ABI: device pointers in, device pointers out.

Each function validates its tensors (CUDA, fp32, contiguous), allocates outputs/workspaces
with torch, and launches on torch's *current* stream (autograd runs backward on a worker
thread, so nothing here caches streams or keeps global state).  Differentiable ops are
``torch.autograd.Function`` s whose backward is again a hand-written kernel.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib

GEOM_STRIDE = 8
FLAG_NO_TENSOR_CORES = 1            # MFB_FLAG_NO_TENSOR_CORES of the C ABI (per call; the library keeps no switches)
# Python-side choices between two implementations of the same result, read at call time (A/B tests)
KDE2D_USE_TENSOR_CORES = True
NSF_BWD_USE_TENSOR_CORES = True


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_f32(name: str, t: torch.Tensor) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: mentflow_b200 kernels need CUDA tensors (got {t.device}); "
                           "there is no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if not t.is_contiguous() or t.data_ptr() % 16:
        t = t.contiguous()
        if t.data_ptr() % 16:
            t = t.clone()
    return t


# --------------------------------------------------------------------------------------
# projection + KDE, 1-D screens
# --------------------------------------------------------------------------------------
def _check_mp(mp: torch.Tensor, k: int, d: int) -> torch.Tensor:
    mp = _check_f32("mp", mp)
    if mp.shape != (k, 2 * d + 4):
        raise ValueError(f"multipole terms must have shape ({k}, {2 * d + 4}), got {tuple(mp.shape)}")
    return mp


def kde1d_sums(x: torch.Tensor, proj: torch.Tensor, geom: torch.Tensor, ratio: float, nbins: int,
               mp: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """S[k, b] = sum_n exp(-0.5 ((u_kn - c_b) / sigma_k)^2)  (unnormalised), u_kn = proj_k . x_n, plus the
    multipole terms of ``mp`` (K, 2D+4) when given (see ``simulate.multipole_terms``)."""
    lib = _lib.load()
    x, proj, geom = _check_f32("x", x), _check_f32("proj", proj), _check_f32("geom", geom)
    n, d = x.shape
    k = proj.shape[0]
    sums = out if out is not None else torch.empty((k, nbins), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        wbytes = lib.mfb_kde1d_workspace_bytes(n, d, k, nbins)
        work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=x.device)
        if mp is None:
            _lib.check(lib.mfb_project_kde1d_fwd(_ptr(x), n, d, _ptr(proj), _ptr(geom), k, nbins, float(ratio),
                                                 _ptr(sums), _ptr(work), wbytes, _stream()), "project_kde1d_fwd")
        else:
            mp = _check_mp(mp, k, d)
            _lib.check(lib.mfb_project_kde1d_mp_fwd(_ptr(x), n, d, _ptr(proj), _ptr(mp), _ptr(geom), k, nbins,
                                                    float(ratio), _ptr(sums), _ptr(work), wbytes, _stream()),
                       "project_kde1d_mp_fwd")
    return sums


def kde1d_normalize(sums: torch.Tensor, n_total: float, geom: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    k, b = sums.shape
    prof = torch.empty_like(sums)
    with torch.cuda.device(sums.device):
        _lib.check(lib.mfb_kde1d_normalize(_ptr(sums), float(n_total), _ptr(geom), k, b, _ptr(prof), _stream()),
                   "kde1d_normalize")
    return prof


def kde1d_normalize_bwd(sums, n_total, geom, gprof):
    lib = _lib.load()
    k, b = sums.shape
    gprof = _check_f32("gprof", gprof)
    gsums = torch.empty_like(sums)
    with torch.cuda.device(sums.device):
        _lib.check(lib.mfb_kde1d_normalize_bwd(_ptr(sums), float(n_total), _ptr(geom), k, b, _ptr(gprof),
                                               _ptr(gsums), _stream()), "kde1d_normalize_bwd")
    return gsums


def kde1d_grad_x(x, proj, geom, ratio, gsums, out: Optional[torch.Tensor] = None, mp: Optional[torch.Tensor] = None):
    lib = _lib.load()
    n, d = x.shape
    k, b = gsums.shape
    acc = 1 if out is not None else 0
    gx = out if out is not None else torch.empty_like(x)
    with torch.cuda.device(x.device):
        if mp is None:
            _lib.check(lib.mfb_project_kde1d_bwd(_ptr(x), n, d, _ptr(proj), _ptr(geom), k, b, float(ratio),
                                                 _ptr(gsums), _ptr(gx), acc, _stream()), "project_kde1d_bwd")
        else:
            mp = _check_mp(mp, k, d)
            _lib.check(lib.mfb_project_kde1d_mp_bwd(_ptr(x), n, d, _ptr(proj), _ptr(mp), _ptr(geom), k, b, float(ratio),
                                                    _ptr(gsums), _ptr(gx), acc, _stream()), "project_kde1d_mp_bwd")
    return gx


KL_PAD = 1.0e-12   # loss.py:15-17


def kde1d_loss_forward(x, proj, geom, ratio, nbins, n_total, meas):
    """Deposit + merge + normalise (+ KL against ``meas``) in two launches: (sums, profiles, kl)."""
    lib = _lib.load()
    x, proj, geom = _check_f32("x", x), _check_f32("proj", proj), _check_f32("geom", geom)
    n, d = x.shape
    k = proj.shape[0]
    sums = torch.empty((k, nbins), dtype=torch.float32, device=x.device)
    prof = torch.empty_like(sums)
    kl = torch.empty(k, dtype=torch.float32, device=x.device) if meas is not None else None
    with torch.cuda.device(x.device):
        wbytes = lib.mfb_kde1d_workspace_bytes(n, d, k, nbins)
        work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=x.device)
        _lib.check(lib.mfb_project_kde1d_loss_fwd(_ptr(x), n, d, _ptr(proj), _ptr(geom), k, nbins, float(ratio),
                                                  float(n_total), _ptr(meas), KL_PAD, _ptr(sums), _ptr(prof),
                                                  _ptr(kl), _ptr(work), wbytes, _stream()), "project_kde1d_loss_fwd")
    return sums, prof, kl


def kde1d_finish(sums, n_total, geom, meas):
    """Normalise merged sums (+ KL against ``meas``): (profiles, kl)."""
    lib = _lib.load()
    k, b = sums.shape
    prof = torch.empty_like(sums)
    kl = torch.empty(k, dtype=torch.float32, device=sums.device) if meas is not None else None
    with torch.cuda.device(sums.device):
        _lib.check(lib.mfb_kde1d_finish(_ptr(sums), float(n_total), _ptr(geom), k, b, _ptr(meas), KL_PAD,
                                        _ptr(prof), _ptr(kl), _stream()), "kde1d_finish")
    return prof, kl


def kde1d_finish_bwd(sums, n_total, geom, meas, gprof, gkl):
    lib = _lib.load()
    k, b = sums.shape
    gprof = _check_f32("gprof", gprof) if gprof is not None else None
    gkl = _check_f32("gkl", gkl) if gkl is not None else None
    gsums = torch.empty_like(sums)
    with torch.cuda.device(sums.device):
        _lib.check(lib.mfb_kde1d_finish_bwd(_ptr(sums), float(n_total), _ptr(geom), k, b, _ptr(meas), KL_PAD,
                                            _ptr(gprof), _ptr(gkl), _ptr(gsums), _stream()), "kde1d_finish_bwd")
    return gsums


def kde1d_loss_forward_p2p(peer, x, proj, geom, ratio, nbins, tail, n_total, meas):
    """Sharded forward tail in two launches: deposit, then merge + cross-rank sum over NVLink peer memory +
    normalisation (+ KL).  Returns (sums, profiles, kl, tail_sum)."""
    lib = _lib.load()
    x, proj, geom = _check_f32("x", x), _check_f32("proj", proj), _check_f32("geom", geom)
    n, d = x.shape
    k = proj.shape[0]
    nt = 0 if tail is None else int(tail.numel())
    block, handle, ptrs, state = peer.block_for(k, nbins, nt, x.device)
    sums = torch.empty((k, nbins), dtype=torch.float32, device=x.device)
    prof = torch.empty_like(sums)
    kl = torch.empty(k, dtype=torch.float32, device=x.device) if meas is not None else None
    tail_out = torch.empty(nt, dtype=torch.float64, device=x.device) if nt else None
    arr = (ctypes.c_uint64 * len(ptrs))(*ptrs)
    with torch.cuda.device(x.device):
        wbytes = lib.mfb_kde1d_workspace_bytes(n, d, k, nbins)
        work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=x.device)
        _lib.check(lib.mfb_project_kde1d_loss_fwd_p2p(_ptr(x), n, d, _ptr(proj), _ptr(geom), k, nbins, float(ratio),
                                                      float(n_total), _ptr(meas), KL_PAD, ctypes.cast(arr, ctypes.c_void_p),
                                                      peer.rank, peer.world, _ptr(state), _ptr(tail), nt, _ptr(sums),
                                                      _ptr(prof), _ptr(kl), _ptr(tail_out), _ptr(work), wbytes, _stream()),
                   "project_kde1d_loss_fwd_p2p")
    return sums, prof, kl, tail_out


def kde1d_finish_p2p(peer, sums_local, tail, n_total, geom, meas):
    """Cross-rank sum over NVLink peer memory + normalisation (+ KL) in one kernel: (sums, profiles, kl, tail_sum).
    ``peer``: ``distributed.PeerExchange``; ``tail``: float64 vector reduced alongside (or None)."""
    lib = _lib.load()
    k, b = sums_local.shape
    nt = 0 if tail is None else int(tail.numel())
    block, handle, ptrs, state = peer.block_for(k, b, nt, sums_local.device)
    sums = torch.empty_like(sums_local)
    prof = torch.empty_like(sums_local)
    kl = torch.empty(k, dtype=torch.float32, device=sums_local.device) if meas is not None else None
    tail_out = torch.empty(nt, dtype=torch.float64, device=sums_local.device) if nt else None
    arr = (ctypes.c_uint64 * len(ptrs))(*ptrs)
    with torch.cuda.device(sums_local.device):
        _lib.check(lib.mfb_kde1d_finish_p2p(ctypes.cast(arr, ctypes.c_void_p), peer.rank, peer.world, _ptr(state),
                                            _ptr(sums_local), _ptr(tail), nt, float(n_total), _ptr(geom), k, b, _ptr(meas),
                                            KL_PAD, _ptr(sums), _ptr(prof), _ptr(kl), _ptr(tail_out), _stream()),
                   "kde1d_finish_p2p")
    return sums, prof, kl, tail_out


_FINISH_MAX_BINS = 6000   # shared-memory bound of the fused tail kernel


class ProjectKDE1D(torch.autograd.Function):
    """profiles[k, b] of K one-dimensional screens, differentiable w.r.t. the particles.

    ``reducer`` (optional) all-reduces the unnormalised sums across ranks *before* the
    non-linear normalisation and returns the global particle count.
    ``meas`` (optional, (K, B)): also return kl[k] = KL(meas_k || profile_k) of loss.py:15-17,
    evaluated by the same kernel that normalises.
    """

    @staticmethod
    def forward(ctx, x, proj, geom, ratio, nbins, reducer, meas, mp=None):
        ctx.set_materialize_grads(False)
        x = _check_f32("x", x)
        n_total = float(x.shape[0])
        if meas is not None:
            meas = _check_f32("meas", meas)
        if reducer is None and mp is None and x.shape[0] > 0 and nbins <= _FINISH_MAX_BINS:
            sums, prof, kl = kde1d_loss_forward(x, proj, geom, ratio, nbins, n_total, meas)
        else:
            # sharded: the unnormalised sums are all-reduced before the non-linear tail; float64 partial sums
            # another op stashed on the reducer (the entropy moments) ride at the end of the same buffer
            peer = getattr(reducer, "peer", None)
            if (peer is not None and mp is None and x.shape[0] > 0 and proj.shape[0] <= 1024 and nbins <= 4096
                    and (proj.shape[0] * nbins) % 2 == 0):
                # the cross-rank sum happens inside the tail kernel, over NVLink peer memory
                n_total = reducer.global_count(n_total)
                stash = reducer.take_stash()
                sums, prof, kl, tail_sum = kde1d_loss_forward_p2p(peer, x, proj, geom, ratio, nbins, stash, n_total, meas)
                if stash is not None:
                    reducer.set_stash_result(tail_sum)
                reducer.calls += 1
                ctx.save_for_backward(x, proj, geom, sums, meas, None, mp)
                ctx.ratio, ctx.n_total = ratio, n_total
                return prof if meas is None else (prof, kl)
            tail = reducer.tail_floats() if hasattr(reducer, "tail_floats") else 0
            if tail:
                flat = torch.empty(proj.shape[0] * nbins + tail, dtype=torch.float32, device=x.device)
                sums = kde1d_sums(x, proj, geom, ratio, nbins, mp, out=flat[: proj.shape[0] * nbins].view(-1, nbins))
                n_total = reducer(sums, n_total, flat=flat)
            else:
                sums = kde1d_sums(x, proj, geom, ratio, nbins, mp)
                if reducer is not None:
                    n_total = reducer(sums, n_total)
            if nbins <= _FINISH_MAX_BINS:
                prof, kl = kde1d_finish(sums, n_total, geom, meas)
            else:
                prof = kde1d_normalize(sums, n_total, geom)
                kl = None
                if meas is not None:
                    kl = (torch.xlogy(meas, meas) - meas * torch.log(prof + KL_PAD)).sum(dim=1) / nbins
        ctx.save_for_backward(x, proj, geom, sums, meas, prof if nbins > _FINISH_MAX_BINS else None, mp)
        ctx.ratio, ctx.n_total = ratio, n_total
        if meas is None:
            return prof
        return prof, kl

    @staticmethod
    def backward(ctx, gprof, gkl=None):
        x, proj, geom, sums, meas, prof, mp = ctx.saved_tensors
        if gprof is None and gkl is None:
            return None, None, None, None, None, None, None, None
        if prof is not None:       # wide screens: torch expression for the KL part
            if gkl is not None:
                extra = -(gkl / sums.shape[1])[:, None] * meas / (prof + KL_PAD)
                gprof = extra if gprof is None else gprof + extra
            gsums = kde1d_normalize_bwd(sums, ctx.n_total, geom, gprof)
        else:
            gsums = kde1d_finish_bwd(sums, ctx.n_total, geom, meas, gprof, gkl if meas is not None else None)
        gx = kde1d_grad_x(x, proj, geom, ctx.ratio, gsums, mp=mp)
        return gx, None, None, None, None, None, None, None


def project_kde1d(x, proj, geom, ratio, nbins, reducer=None, meas=None, mp=None):
    """(K, B) profiles; with ``meas`` a pair (profiles, kl[K]); ``mp``: multipole terms (K, 2D+4)."""
    return ProjectKDE1D.apply(x, proj, geom, ratio, nbins, reducer, meas, mp)


# --------------------------------------------------------------------------------------
# projection + exact histogram, 1-D screens
# --------------------------------------------------------------------------------------
def project_hist1d(x: torch.Tensor, proj: torch.Tensor, edges: torch.Tensor,
                   counts: Optional[torch.Tensor] = None, mp: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int64 counts[k, b]; edges is (K, B+1).  Adds into ``counts`` when given.  ``mp``: multipole
    terms (K, 2D+4) of the transfer maps, see ``simulate.multipole_terms``."""
    lib = _lib.load()
    x, proj, edges = _check_f32("x", x), _check_f32("proj", proj), _check_f32("edges", edges)
    n, d = x.shape
    k, b = proj.shape[0], edges.shape[1] - 1
    if counts is None:
        counts = torch.zeros((k, b), dtype=torch.int64, device=x.device)
    if n == 0:              # nothing to count
        return counts
    with torch.cuda.device(x.device):
        if mp is None:
            _lib.check(lib.mfb_project_hist1d(_ptr(x), n, d, _ptr(proj), _ptr(edges), k, b, _ptr(counts), _stream()),
                       "project_hist1d")
        else:
            mp = _check_mp(mp, k, d)
            _lib.check(lib.mfb_project_hist1d_mp(_ptr(x), n, d, _ptr(proj), _ptr(mp), _ptr(edges), k, b, _ptr(counts),
                                                 _stream()), "project_hist1d_mp")
    return counts


# --------------------------------------------------------------------------------------
# 2-D screens
# --------------------------------------------------------------------------------------
def kde2d_sums(x, proj, geom, ratio, bx, by):
    """(sums float32 [K,bx,by], acc int64 [2,K,bx,by]): acc holds the exact fixed-point
    accumulators (plane 0: value * 2^22, plane 1: remainder * 2^44) that ranks all-reduce."""
    lib = _lib.load()
    x, proj, geom = _check_f32("x", x), _check_f32("proj", proj), _check_f32("geom", geom)
    n, d = x.shape
    k = proj.shape[0]
    sums = torch.empty((k, bx, by), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        # workspace = the two int64 planes, then the per-CTA partial screens of the tensor-core path
        wbytes = int(lib.mfb_kde2d_workspace_bytes(n, d, k, bx, by))
        work = torch.empty((wbytes + 7) // 8, dtype=torch.int64, device=x.device)
        _lib.check(lib.mfb_project_kde2d_fwd(_ptr(x), n, d, _ptr(proj), _ptr(geom), k, bx, by, float(ratio),
                                             _ptr(sums), _ptr(work), wbytes,
                                             0 if KDE2D_USE_TENSOR_CORES else FLAG_NO_TENSOR_CORES, _stream()),
                   "project_kde2d_fwd")
    acc = work[: 2 * k * bx * by].view(2, k, bx, by)
    return sums, acc


def kde2d_normalize(sums, geom):
    lib = _lib.load()
    k, bx, by = sums.shape
    prof = torch.empty_like(sums)
    with torch.cuda.device(sums.device):
        _lib.check(lib.mfb_kde2d_normalize(_ptr(sums), _ptr(geom), k, bx, by, _ptr(prof), _stream()),
                   "kde2d_normalize")
    return prof


class ProjectKDE2D(torch.autograd.Function):
    """profiles[k, a, b] of K two-dimensional screens, differentiable w.r.t. the particles."""

    @staticmethod
    def forward(ctx, x, proj, geom, ratio, bx, by, reducer):
        x = _check_f32("x", x)
        sums, acc = kde2d_sums(x, proj, geom, ratio, bx, by)
        if reducer is not None:
            reducer(acc, float(x.shape[0]))          # exact integer all-reduce
            sums = (acc[0].to(torch.float64) * 2.0 ** -22 + acc[1].to(torch.float64) * 2.0 ** -44).to(torch.float32)
        prof = kde2d_normalize(sums, geom)
        ctx.save_for_backward(x, proj, geom, sums)
        ctx.ratio = ratio
        return prof

    @staticmethod
    def backward(ctx, gprof):
        lib = _lib.load()
        x, proj, geom, sums = ctx.saved_tensors
        k, bx, by = sums.shape
        gprof = _check_f32("gprof", gprof)
        gsums = torch.empty_like(sums)
        gx = torch.empty_like(x)
        n, d = x.shape
        with torch.cuda.device(x.device):
            _lib.check(lib.mfb_kde2d_normalize_bwd(_ptr(sums), _ptr(geom), k, bx, by, _ptr(gprof), _ptr(gsums),
                                                   _stream()), "kde2d_normalize_bwd")
            _lib.check(lib.mfb_project_kde2d_bwd(_ptr(x), n, d, _ptr(proj), _ptr(geom), k, bx, by,
                                                 float(ctx.ratio), _ptr(gsums), _ptr(gx), 0, _stream()),
                       "project_kde2d_bwd")
        return gx, None, None, None, None, None, None


def project_kde2d(x, proj, geom, ratio, bx, by, reducer=None):
    return ProjectKDE2D.apply(x, proj, geom, ratio, bx, by, reducer)


def project_hist2d(x, proj, edges_x, edges_y, counts=None):
    """int64 counts[k, a, b]; edges_x (K, bx+1), edges_y (K, by+1)."""
    lib = _lib.load()
    x, proj = _check_f32("x", x), _check_f32("proj", proj)
    edges_x, edges_y = _check_f32("edges_x", edges_x), _check_f32("edges_y", edges_y)
    n, d = x.shape
    k, bx, by = proj.shape[0], edges_x.shape[1] - 1, edges_y.shape[1] - 1
    if counts is None:
        counts = torch.zeros((k, bx, by), dtype=torch.int64, device=x.device)
    if n == 0:              # nothing to count
        return counts
    with torch.cuda.device(x.device):
        _lib.check(lib.mfb_project_hist2d(_ptr(x), n, d, _ptr(proj), _ptr(edges_x), _ptr(edges_y), k, bx, by,
                                          _ptr(counts), _stream()), "project_hist2d")
    return counts


# --------------------------------------------------------------------------------------
# neural spline flow
# --------------------------------------------------------------------------------------
def nsf_layer_param_floats(d: int, hidden_units: int, hidden_layers: int, bins: int) -> int:
    return int(_lib.load().mfb_nsf_layer_param_floats(d, hidden_units, hidden_layers, bins))


NSF_USE_TENSOR_CORES = True   # tcgen05 conditioner where the configuration is compiled; False = fp32 CUDA-core kernel


def orders_autoregressive(orders) -> bool:
    """every layer's order is a permutation (strict autoregressive ordering; coupling layers repeat values)"""
    return all(sorted(int(v) for v in o) == list(range(len(o))) for o in orders)


def nsf_tc_supported(d: int, hidden_units: int, hidden_layers: int, bins: int, orders=None) -> bool:
    """the tcgen05 layer kernels are compiled for this shape (and, when ``orders`` is given, these orderings)"""
    if orders is not None and not orders_autoregressive(orders):
        return False
    return bool(NSF_USE_TENSOR_CORES and _lib.load().mfb_nsf_tc_supported(d, hidden_units, hidden_layers, bins))


def nsf_tc_images(packed, orders, hidden_units, hidden_layers, bins):
    """fp16 (hi, lo) SWIZZLE_128B operand images of every layer's masked weights, (T, bytes) uint8;
    built on the device from the packed fp32 parameters (one small launch per forward call)."""
    lib = _lib.load()
    packed = _check_f32("packed", packed)
    t_layers, d = len(orders), len(orders[0])
    nbytes = int(lib.mfb_nsf_tc_image_bytes(d, hidden_layers))
    images = torch.empty((t_layers, nbytes), dtype=torch.uint8, device=packed.device)
    order_arr = (ctypes.c_int32 * (t_layers * d))(*[int(o) for order in orders for o in order])
    with torch.cuda.device(packed.device):
        _lib.check(lib.mfb_nsf_tc_prepare(_ptr(packed), packed.stride(0), t_layers, d, hidden_units, hidden_layers,
                                          bins, ctypes.cast(order_arr, ctypes.c_void_p), _ptr(images), None, 0,
                                          _stream()), "nsf_tc_prepare")
    return images


def nsf_layer_forward(v, params, order, hidden_units, hidden_layers, bins, logq_in, first_layer,
                      want_logq=True, image=None):
    """One autoregressive spline layer: returns (y, logq_out).  With ``image`` (one row of
    ``nsf_tc_images``) the tensor-core kernel runs, otherwise the fp32 CUDA-core kernel."""
    lib = _lib.load()
    v = _check_f32("v", v)
    n, d = v.shape
    y = torch.empty_like(v)
    logq_out = torch.empty(n, dtype=torch.float32, device=v.device) if want_logq else None
    if n == 0:          # an empty batch has no storage to point at: nothing to launch
        return y, logq_out
    order_arr = (ctypes.c_int32 * d)(*[int(o) for o in order])
    with torch.cuda.device(v.device):
        if image is not None:
            _lib.check(lib.mfb_nsf_tc_layer_fwd(_ptr(v), n, d, hidden_units, hidden_layers, bins, _ptr(image),
                                                ctypes.cast(order_arr, ctypes.c_void_p), _ptr(logq_in),
                                                1 if first_layer else 0, _ptr(y), _ptr(logq_out), _stream()),
                       "nsf_tc_layer_fwd")
        else:
            params = _check_f32("params", params)
            _lib.check(lib.mfb_nsf_layer_fwd(_ptr(v), n, d, hidden_units, hidden_layers, bins, _ptr(params),
                                             ctypes.cast(order_arr, ctypes.c_void_p), _ptr(logq_in),
                                             1 if first_layer else 0, _ptr(y), _ptr(logq_out), _stream()),
                       "nsf_layer_fwd")
    return y, logq_out


def _nsf_run_layers(z, packed, orders, hidden_units, hidden_layers, bins, want_logq, images=None):
    """All layers in sampling order; returns ([z, y_1, ..., x], log q).  ``images``: operand images
    of the tensor-core kernel when the caller has them cached for these weights."""
    d = z.shape[1]
    if not nsf_tc_supported(d, hidden_units, hidden_layers, bins, orders):
        images = None
    elif images is None:
        images = nsf_tc_images(packed, orders, hidden_units, hidden_layers, bins)
    steps, logq = [z], None
    for t, order in enumerate(orders):
        y, logq = nsf_layer_forward(steps[-1], packed[t], order, hidden_units, hidden_layers, bins, logq,
                                    first_layer=(t == 0), want_logq=want_logq,
                                    image=None if images is None else images[t])
        steps.append(y)
    return steps, logq


# --------------------------------------------------------------------------------------
# moments
# --------------------------------------------------------------------------------------
def moments(x: torch.Tensor, logq: Optional[torch.Tensor], with_cov: bool = False) -> torch.Tensor:
    """float64 vector: [sum logq, sum |x|^2] (+ [sum x_i, sum x_i x_j] with ``with_cov``)."""
    lib = _lib.load()
    x = _check_f32("x", x)
    if logq is not None:
        logq = _check_f32("logq", logq)
    n, d = x.shape
    m = 2 + d + d * d if with_cov else 2      # every entry is written by the kernel
    out = torch.empty(m, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        wbytes = lib.mfb_moments_workspace_bytes(n, d)
        work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=x.device)
        _lib.check(lib.mfb_moments(_ptr(x), _ptr(logq), n, d, 1 if with_cov else 0, _ptr(out), _ptr(work), wbytes,
                                   _stream()), "moments")
    return out


def mc_entropy(sums: torch.Tensor, a: float, b: float, c: float) -> torch.Tensor:
    """float32 scalar a * sums[0] + b * sums[1] - c, evaluated in double by one launch."""
    lib = _lib.load()
    assert sums.dtype == torch.float64 and sums.numel() >= 2 and sums.is_contiguous()
    h = torch.empty((), dtype=torch.float32, device=sums.device)
    with torch.cuda.device(sums.device):
        _lib.check(lib.mfb_mc_entropy(_ptr(sums), float(a), float(b), float(c), _ptr(h), _stream()), "mc_entropy")
    return h


class LossTail(torch.autograd.Function):
    """L = H + mu * mean(D) as one launch (core.py:111-113); dL/dH = 1, dL/dD_k = mu / K."""

    @staticmethod
    def forward(ctx, h, d, mu):
        lib = _lib.load()
        dd = _check_f32("d", d.detach())
        hh = None if h is None else _check_f32("h", h.detach())
        out = torch.empty(2, dtype=torch.float32, device=dd.device)
        with torch.cuda.device(dd.device):
            _lib.check(lib.mfb_loss_tail(_ptr(dd), dd.numel(), _ptr(hh), float(mu), _ptr(out), _stream()), "loss_tail")
        ctx.k, ctx.mu, ctx.shape, ctx.has_h = dd.numel(), float(mu), d.shape, h is not None
        return out[0]

    @staticmethod
    def backward(ctx, g):
        gd = (g * (ctx.mu / ctx.k)).expand(ctx.shape)
        return (g if ctx.has_h else None), gd, None


def f64_split(values: torch.Tensor, out: torch.Tensor) -> None:
    """float64 (n,) -> float32 (2n,) [hi | lo] written into ``out`` (a slice of an all-reduce buffer)."""
    lib = _lib.load()
    n = values.numel()
    assert values.dtype == torch.float64 and out.dtype == torch.float32 and out.numel() == 2 * n and out.is_contiguous()
    with torch.cuda.device(values.device):
        _lib.check(lib.mfb_f64_split(_ptr(values), n, _ptr(out), _stream()), "f64_split")


def f64_join(pairs: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    n = pairs.numel() // 2
    out = torch.empty(n, dtype=torch.float64, device=pairs.device)
    with torch.cuda.device(pairs.device):
        _lib.check(lib.mfb_f64_join(_ptr(pairs), n, _ptr(out), _stream()), "f64_join")
    return out


# --------------------------------------------------------------------------------------
# zuko-layout parameters <-> packed blocks
# --------------------------------------------------------------------------------------
class PackParameters(torch.autograd.Function):
    """(packed (T, floats_per_layer), packed_om (T, om_floats)) from the zuko-layout tensors of the whole flow in one
    launch; backward = one launch the other way (masked entries get exactly zero gradient).  ``packed_om`` is the
    same numbers in the backward kernels' out-major layout and carries no gradient of its own."""

    @staticmethod
    def forward(ctx, w_in, b_in, w_hid, b_hid, w_out, b_out, m_in, m_hid, m_out, dims):
        lib = _lib.load()
        T, d, hidden_units, hidden_layers, bins = dims
        ts = [_check_f32(n, t) for n, t in zip(("w_in", "b_in", "w_hid", "b_hid", "w_out", "b_out", "m_in", "m_hid", "m_out"),
                                                (w_in, b_in, w_hid, b_hid, w_out, b_out, m_in, m_hid, m_out))]
        dev = w_in.device
        with torch.cuda.device(dev):
            np_ = int(lib.mfb_nsf_layer_param_floats(d, hidden_units, hidden_layers, bins))
            nom = int(lib.mfb_nsf_layer_param_om_floats(d, hidden_units, hidden_layers))
            packed = torch.empty((T, np_), dtype=torch.float32, device=dev)
            packed_om = torch.empty((T, nom), dtype=torch.float32, device=dev)
            _lib.check(lib.mfb_nsf_pack_params(*[_ptr(t) if t.numel() else None for t in ts], T, d, hidden_units,
                                               hidden_layers, bins, _ptr(packed), _ptr(packed_om), _stream()),
                       "nsf_pack_params")
        ctx.save_for_backward(ts[6], ts[7], ts[8])
        ctx.dims = dims
        ctx.shapes = [t.shape for t in ts[:6]]
        ctx.mark_non_differentiable(packed_om)
        return packed, packed_om

    @staticmethod
    def backward(ctx, gpacked, _gom):
        lib = _lib.load()
        m_in, m_hid, m_out = ctx.saved_tensors
        T, d, hidden_units, hidden_layers, bins = ctx.dims
        gpacked = _check_f32("gpacked", gpacked)
        grads = [torch.empty(sh, dtype=torch.float32, device=gpacked.device) for sh in ctx.shapes]
        with torch.cuda.device(gpacked.device):
            _lib.check(lib.mfb_nsf_unpack_grads(_ptr(gpacked), _ptr(m_in), _ptr(m_hid) if m_hid.numel() else None,
                                                _ptr(m_out), T, d, hidden_units, hidden_layers, bins,
                                                *[_ptr(g) if g.numel() else None for g in grads], _stream()),
                       "nsf_unpack_grads")
        return (*grads, None, None, None, None)


# --------------------------------------------------------------------------------------
# whole flow (all layers), differentiable
# --------------------------------------------------------------------------------------
class NSFForward(torch.autograd.Function):
    """x, log q = flow(z): one kernel launch per autoregressive layer.  Only the layer inputs are
    kept (24 B / particle / layer); backward recomputes the conditioner activations instead of
    storing them (10.9 KB / particle in the reference's autograd graph)."""

    @staticmethod
    def forward(ctx, z, packed, packed_om, meta):
        orders, hidden_units, hidden_layers, bins, want_logq, images = meta
        z = _check_f32("z", z)
        packed = _check_f32("packed", packed)
        if images is None and NSF_USE_TENSOR_CORES and nsf_tc_supported(z.shape[1], hidden_units, hidden_layers, bins, orders):
            images = nsf_tc_images(packed, orders, hidden_units, hidden_layers, bins)
        steps, logq = _nsf_run_layers(z, packed, orders, hidden_units, hidden_layers, bins, want_logq, images)
        ctx.meta = meta
        ctx.images = images        # the backward kernels read the same operand images (~1 MB, not rebuilt)
        ctx.save_for_backward(packed, packed_om, *steps[:-1])
        if logq is None:
            logq = z.new_empty(0)
        return steps[-1], logq

    @staticmethod
    def backward(ctx, gx, glogq):
        orders, hidden_units, hidden_layers, bins, want_logq, _ = ctx.meta
        images = ctx.images
        packed, packed_om, *inputs = ctx.saved_tensors
        n, d = inputs[0].shape
        gx = _check_f32("gx", gx) if gx is not None else torch.zeros_like(inputs[0])
        gl = _check_f32("glogq", glogq) if (want_logq and glogq is not None and glogq.numel() == n) else None
        gz, gpacked = nsf_backward(inputs, packed, packed_om, orders, hidden_units, hidden_layers, bins, gx, gl,
                                   images=images)
        return gz, gpacked, None, None


NSF_BWD_CHUNK = 1 << 20   # particles per backward pass: bounds the recompute workspace (~2.9 KB/particle)


def nsf_backward(inputs, packed, packed_om, orders, hidden_units, hidden_layers, bins, gx, glogq, images=None):
    """(dL/dz, dL/dpacked) given the per-layer inputs saved by the forward pass.  ``images``: the tcgen05
    operand images of the same packed weights (``nsf_tc_images``), reused instead of being rebuilt."""
    lib = _lib.load()
    n, d = inputs[0].shape
    dev = inputs[0].device
    gpacked = torch.zeros_like(packed)
    gz = torch.empty_like(inputs[0])
    if n == 0:
        return gz, gpacked
    bwd_flags = 0 if (NSF_BWD_USE_TENSOR_CORES and orders_autoregressive(orders)) else FLAG_NO_TENSOR_CORES
    chunk = min(n, NSF_BWD_CHUNK)
    with torch.cuda.device(dev):
        wbytes = lib.mfb_nsf_layer_bwd_workspace_bytes(chunk, d, hidden_layers)
        work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        for start in range(0, n, chunk):
            m = min(chunk, n - start)
            g = gx[start:start + m]
            gl = None if glogq is None else glogq[start:start + m]
            for t in range(len(orders) - 1, -1, -1):
                order_arr = (ctypes.c_int32 * d)(*[int(o) for o in orders[t]])
                out = gz[start:start + m] if t == 0 else torch.empty((m, d), dtype=torch.float32, device=dev)
                _lib.check(lib.mfb_nsf_layer_bwd_img(_ptr(inputs[t][start:start + m]), _ptr(g), _ptr(gl), m, d,
                                                     hidden_units, hidden_layers, bins, _ptr(packed[t]),
                                                     _ptr(packed_om[t]), ctypes.cast(order_arr, ctypes.c_void_p),
                                                     1 if t == 0 else 0, _ptr(images[t]) if images is not None else None,
                                                     _ptr(out), _ptr(gpacked[t]), 1 if start > 0 else 0, _ptr(work),
                                                     wbytes, bwd_flags,
                                                     _stream()), "nsf_layer_bwd")
                g = out
    return gz, gpacked


def nsf_forward(z, packed, packed_om, orders, hidden_units, hidden_layers, bins, want_logq=True, want_steps=False,
                images=None):
    """Returns (x, logq or None, steps or None)."""
    if want_steps or not (torch.is_grad_enabled() and (z.requires_grad or packed.requires_grad)):
        # nothing to differentiate: skip the autograd.Function bookkeeping
        z = _check_f32("z", z.detach())
        steps, logq = _nsf_run_layers(z, _check_f32("packed", packed.detach()), orders, hidden_units, hidden_layers,
                                      bins, want_logq, images)
        return steps[-1], logq, (steps if want_steps else None)
    meta = (tuple(tuple(o) for o in orders), hidden_units, hidden_layers, bins, bool(want_logq), images)
    x, logq = NSFForward.apply(z, packed, packed_om, meta)
    return x, (logq if want_logq else None), None


NSF_INV_USE_TENSOR_CORES = True   # tcgen05 inverse where the forward's operand images exist; False = fp32 CUDA-core kernel


def nsf_inverse(x, packed, orders, hidden_units, hidden_layers, bins, want_logq=True, want_steps=False, images=None):
    """Density direction: returns (z, log q(x) or None, steps or None).  Not differentiable (no
    experiment of the reference trains through log_prob(x)).  ``images``: the forward kernel's operand
    images (``nsf_tc_images``): with them each layer is one tensor-core launch (S x 4 GEMM round trips per tile)
    instead of D full sweeps of the CUDA-core kernel."""
    lib = _lib.load()
    x = _check_f32("x", x.detach())
    packed = _check_f32("packed", packed.detach())
    n, d = x.shape
    steps = [x]
    acc = None
    with torch.cuda.device(x.device):
        for t in range(len(orders) - 1, -1, -1):
            v = torch.empty_like(x)
            out = torch.empty(n, dtype=torch.float32, device=x.device) if want_logq else None
            order_arr = (ctypes.c_int32 * d)(*[int(o) for o in orders[t]])
            if n == 0:          # empty batch: nothing to launch
                acc = out
                steps.append(v)
                continue
            if images is not None and NSF_INV_USE_TENSOR_CORES:
                _lib.check(lib.mfb_nsf_tc_layer_inv(_ptr(steps[-1]), n, d, hidden_units, hidden_layers, bins,
                                                    _ptr(images[t]), ctypes.cast(order_arr, ctypes.c_void_p), _ptr(acc),
                                                    1 if t == 0 else 0, _ptr(v), _ptr(out), _stream()), "nsf_tc_layer_inv")
            else:
                _lib.check(lib.mfb_nsf_layer_inv(_ptr(steps[-1]), n, d, hidden_units, hidden_layers, bins, _ptr(packed[t]),
                                                 ctypes.cast(order_arr, ctypes.c_void_p), _ptr(acc), 1 if t == 0 else 0,
                                                 _ptr(v), _ptr(out), _stream()), "nsf_layer_inv")
            acc = out
            steps.append(v)
    return steps[-1], acc, (steps if want_steps else None)


# --------------------------------------------------------------------------------------
# classical MENT
# --------------------------------------------------------------------------------------
def _host_i32(values):
    arr = (ctypes.c_int32 * len(values))(*[int(v) for v in values])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


def _host_f32(values):
    arr = (ctypes.c_float * len(values))(*[float(v) for v in values])
    return arr, ctypes.cast(arr, ctypes.c_void_p)


def ment_prob(x, proj, coords, tables, neg_half_inv_s2: float, log_norm: float) -> torch.Tensor:
    """rho at explicit points x (G, D); proj (K, D), coords / tables (K, B)."""
    lib = _lib.load()
    x, proj = _check_f32("x", x), _check_f32("proj", proj)
    coords, tables = _check_f32("coords", coords), _check_f32("tables", tables)
    g, d = x.shape
    k, b = tables.shape
    out = torch.empty(g, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.mfb_ment_prob(_ptr(x), g, d, _ptr(proj), _ptr(coords), _ptr(tables), k, b,
                                     float(neg_half_inv_s2), float(log_norm), _ptr(out), _stream()), "ment_prob")
    return out


def ment_prob_grid(shape, first_centre, step, proj, coords, tables, neg_half_inv_s2, log_norm) -> torch.Tensor:
    """rho on the centres of a regular grid (flattened, 'ij' order) without materialising the points."""
    lib = _lib.load()
    proj, coords, tables = _check_f32("proj", proj), _check_f32("coords", coords), _check_f32("tables", tables)
    d = len(shape)
    k, b = tables.shape
    g = 1
    for s in shape:
        g *= int(s)
    out = torch.empty(g, dtype=torch.float32, device=proj.device)
    a1, p1 = _host_i32(shape)
    a2, p2 = _host_f32(first_centre)
    a3, p3 = _host_f32(step)
    with torch.cuda.device(proj.device):
        _lib.check(lib.mfb_ment_prob_grid(d, p1, p2, p3, _ptr(proj), _ptr(coords), _ptr(tables), k, b,
                                          float(neg_half_inv_s2), float(log_norm), _ptr(out), _stream()),
                   "ment_prob_grid")
    return out


def ment_integrate(d, meas_coords, meas_axis, int_shape, int_first, int_step, minv, proj, coords, tables,
                   neg_half_inv_s2, log_norm) -> torch.Tensor:
    lib = _lib.load()
    meas_coords, minv = _check_f32("meas_coords", meas_coords), _check_f32("minv", minv)
    proj, coords, tables = _check_f32("proj", proj), _check_f32("coords", coords), _check_f32("tables", tables)
    k, b = tables.shape
    nb = meas_coords.shape[0]
    pred = torch.empty(nb, dtype=torch.float32, device=proj.device)
    a1, p1 = _host_i32(int_shape)
    a2, p2 = _host_f32(int_first)
    a3, p3 = _host_f32(int_step)
    with torch.cuda.device(proj.device):
        _lib.check(lib.mfb_ment_integrate(d, _ptr(meas_coords), nb, int(meas_axis), len(int_shape), p1, p2, p3,
                                          _ptr(minv), _ptr(proj), _ptr(coords), _ptr(tables), k, b,
                                          float(neg_half_inv_s2), float(log_norm), _ptr(pred), _stream()),
                   "ment_integrate")
    return pred


def _ment_groups(g1, g2, device):
    """C-ABI arguments of the two table families: g1 = (proj (K,D), coords (K,B), tables (K,B)) or None,
    g2 = (proj2 (K2,2,D), cx (K2,Bx), cy (K2,By), tables2 (K2,Bx,By)) or None."""
    keep = []
    if g1 is not None:
        proj, coords, tables = (_check_f32(n, t) for n, t in zip(("proj", "coords", "tables"), g1))
        keep += [proj, coords, tables]
        a1 = [_ptr(proj), _ptr(coords), _ptr(tables), tables.shape[0], tables.shape[1]]
    else:
        a1 = [None, None, None, 0, 2]
    if g2 is not None:
        proj2, cx, cy, tab2 = (_check_f32(n, t) for n, t in zip(("proj2", "cx", "cy", "tables2"), g2))
        if proj2.ndim != 3 or proj2.shape[1] != 2 or tab2.shape != (proj2.shape[0], cx.shape[1], cy.shape[1]):
            raise ValueError("2-D Lagrange tables: proj2 (K,2,D), cx (K,Bx), cy (K,By), tables2 (K,Bx,By) expected")
        keep += [proj2, cx, cy, tab2]
        a2 = [_ptr(proj2), _ptr(cx), _ptr(cy), _ptr(tab2), tab2.shape[0], tab2.shape[1], tab2.shape[2]]
    else:
        a2 = [None, None, None, None, 0, 0, 0]
    return a1, a2, keep


def ment_prob_nd(x, g1, g2, neg_half_inv_s2: float, log_norm: float) -> torch.Tensor:
    """rho at explicit points with one- and two-dimensional Lagrange tables (see ``_ment_groups``)."""
    lib = _lib.load()
    x = _check_f32("x", x)
    g, d = x.shape
    a1, a2, keep = _ment_groups(g1, g2, x.device)
    out = torch.empty(g, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.mfb_ment_prob_nd(_ptr(x), g, d, *a1, *a2, float(neg_half_inv_s2), float(log_norm), _ptr(out),
                                        _stream()), "ment_prob_nd")
    return out


def ment_prob_grid_nd(shape, first_centre, step, g1, g2, neg_half_inv_s2, log_norm, device) -> torch.Tensor:
    lib = _lib.load()
    d = len(shape)
    g = 1
    for s_ in shape:
        g *= int(s_)
    a1, a2, keep = _ment_groups(g1, g2, device)
    out = torch.empty(g, dtype=torch.float32, device=device)
    h1, p1 = _host_i32(shape)
    h2, p2 = _host_f32(first_centre)
    h3, p3 = _host_f32(step)
    with torch.cuda.device(device):
        _lib.check(lib.mfb_ment_prob_grid_nd(d, p1, p2, p3, *a1, *a2, float(neg_half_inv_s2), float(log_norm), _ptr(out),
                                             _stream()), "ment_prob_grid_nd")
    return out


def ment_integrate_nd(d, meas_coords, meas_axis, meas_coords2, meas_axis2, int_shape, int_first, int_step, minv, g1, g2,
                      neg_half_inv_s2, log_norm) -> torch.Tensor:
    """Integration-mode prediction of one screen: 1-D (meas_axis2 = -1) -> (B,), 2-D -> (Bx, By)."""
    lib = _lib.load()
    meas_coords, minv = _check_f32("meas_coords", meas_coords), _check_f32("minv", minv)
    nb = meas_coords.shape[0]
    nb2 = 0
    if meas_axis2 >= 0:
        meas_coords2 = _check_f32("meas_coords2", meas_coords2)
        nb2 = meas_coords2.shape[0]
    a1, a2, keep = _ment_groups(g1, g2, minv.device)
    pred = torch.empty((nb, nb2) if meas_axis2 >= 0 else (nb,), dtype=torch.float32, device=minv.device)
    h1, p1 = _host_i32(int_shape)
    h2, p2 = _host_f32(int_first)
    h3, p3 = _host_f32(int_step)
    with torch.cuda.device(minv.device):
        _lib.check(lib.mfb_ment_integrate_nd(d, _ptr(meas_coords), nb, int(meas_axis),
                                             _ptr(meas_coords2) if meas_axis2 >= 0 else None, nb2, int(meas_axis2),
                                             len(int_shape), p1, p2, p3, _ptr(minv), *a1, *a2, float(neg_half_inv_s2),
                                             float(log_norm), _ptr(pred), _stream()), "ment_integrate_nd")
    return pred


def cdf_sample(rho: torch.Tensor, shape, first_edge, cell, size: int, seed: int, offset: int = 0,
               jitter: bool = False, pad: float = 1.0e-15) -> torch.Tensor:
    """`size` particles from the piecewise-constant density rho (flattened grid, 'ij' order)."""
    lib = _lib.load()
    rho = _check_f32("rho", rho.reshape(-1))
    g = rho.numel()
    d = len(shape)
    cdf = torch.empty(g, dtype=torch.float64, device=rho.device)
    out = torch.empty((int(size), d), dtype=torch.float32, device=rho.device)
    a1, p1 = _host_i32(shape)
    a2, p2 = _host_f32(first_edge)
    a3, p3 = _host_f32(cell)
    with torch.cuda.device(rho.device):
        wbytes = lib.mfb_cdf_workspace_bytes(g)
        work = torch.empty(wbytes, dtype=torch.uint8, device=rho.device)
        _lib.check(lib.mfb_cdf_build(_ptr(rho), g, float(pad), _ptr(cdf), _ptr(work), wbytes, _stream()), "cdf_build")
        _lib.check(lib.mfb_cdf_sample(_ptr(cdf), g, _ptr(work), d, p1, p2, p3, 1 if jitter else 0,
                                      int(seed) & (2 ** 64 - 1), int(offset), int(size), _ptr(out), _stream()),
                   "cdf_sample")
    return out


def gs_update(table: torch.Tensor, meas: torch.Tensor, pred: torch.Tensor, lr: float, thresh: float) -> None:
    """In-place Gauss-Seidel update of one Lagrange table."""
    lib = _lib.load()
    n = table.numel()
    meas, pred = _check_f32("meas", meas), _check_f32("pred", pred)
    if not (table.is_cuda and table.is_contiguous() and table.dtype == torch.float32):
        raise RuntimeError("gs_update: table must be a contiguous CUDA float32 tensor (no CPU fallback)")
    with torch.cuda.device(table.device):
        _lib.check(lib.mfb_gs_update(_ptr(table), _ptr(meas), _ptr(pred), n, float(lr), float(thresh), _stream()),
                   "gs_update")


# --------------------------------------------------------------------------------------
# base noise on the device (generate/flows/zuko.py:15-16 -> torch.randn)
# --------------------------------------------------------------------------------------
class PhiloxStream:
    """torch's CUDA generator, mirrored for the library's randn kernel.

    ``normal_(out)`` fills ``out`` with exactly the values ``torch.randn(out.shape, device=...)`` would
    produce for the generator's current (seed, offset) and moves the generator on by the same amount, so
    library draws and torch draws interleave as if all of them were torch's.  While a CUDA graph is being
    captured the (seed, offset) pair is read from a device tensor instead, and the kernel advances it
    itself; ``sync()`` before a replay re-primes that tensor if the generator was re-seeded or used by
    somebody else in the meantime, ``consumed()`` after it books the replay's draws on the generator.
    """

    def __init__(self, device) -> None:
        self.device = torch.device(device)
        self.state: Optional[torch.Tensor] = None   # int64[2] on the device: seed, offset
        self._mirror = None                          # (seed, offset) the device tensor holds
        self._pinned = None
        self.captured_increment = 0                  # offset consumed by one replay of the captured draws

    def _generator(self):
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        return torch.cuda.default_generators[idx]

    @staticmethod
    def _as_i64(v: int) -> int:
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v

    def normal_(self, out: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        if not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous():
            raise RuntimeError("PhiloxStream.normal_: contiguous float32 CUDA tensor expected; there is no CPU fallback")
        numel = out.numel()
        if numel == 0:
            return out
        with torch.cuda.device(out.device):
            inc = int(lib.mfb_randn_offset_increment(numel))
            if torch.cuda.is_current_stream_capturing():
                if self.state is None:
                    raise RuntimeError("PhiloxStream: call begin_capture() before capturing a draw")
                _lib.check(lib.mfb_randn_philox_state(_ptr(out), numel, _ptr(self.state), 1, _stream()), "randn_philox_state")
                self.captured_increment += inc
            else:
                g = self._generator()
                seed, off = int(g.initial_seed()), int(g.get_offset())
                _lib.check(lib.mfb_randn_philox(_ptr(out), numel, ctypes.c_uint64(seed & ((1 << 64) - 1)),
                                                ctypes.c_uint64(off), _stream()), "randn_philox")
                g.set_offset(off + inc)
        return out

    def mark(self) -> int:
        """current offset of torch's generator (see ``rewind``)"""
        return int(self._generator().get_offset())

    def rewind(self, mark: int) -> None:
        """put torch's generator back to a marked offset (warm-up passes of a capture must not consume
        the caller's random stream)"""
        self._generator().set_offset(int(mark))

    # ---- graph replay support ---------------------------------------------------------------
    def begin_capture(self) -> None:
        """Allocate / prime the device state; call right before the graph capture starts."""
        if self.state is None:
            self.state = torch.zeros(2, dtype=torch.int64, device=self.device)
            # ring of pinned staging slots, each guarded by an event: a slot is rewritten by the host only after
            # the asynchronous copy that last read it has completed (re-seeding every step without synchronising
            # in between must not let a later state overtake an earlier replay)
            self._pinned = torch.zeros((8, 2), dtype=torch.int64).pin_memory()
            self._pinned_np = self._pinned.numpy()       # the same memory: element writes without a dispatcher trip
            self._slot_events = [None] * 8
            self._slot = 0
        self.captured_increment = 0
        self._mirror = None
        self.sync()

    def sync(self) -> None:
        """Make the device state equal to torch's generator (no-op in steady state)."""
        g = self._generator()
        cur = (int(g.initial_seed()), int(g.get_offset()))
        if self._mirror != cur:
            i = self._slot
            self._slot = (i + 1) % self._pinned.shape[0]
            ev = self._slot_events[i]
            if ev is not None:
                ev.synchronize()
            else:
                ev = self._slot_events[i] = torch.cuda.Event()
            self._pinned_np[i, 0] = self._as_i64(cur[0])
            self._pinned_np[i, 1] = self._as_i64(cur[1])
            self.state.copy_(self._pinned[i], non_blocking=True)
            ev.record()
            self._mirror = cur

    def consumed(self) -> None:
        """Book one replay of the captured draws on torch's generator."""
        if self.captured_increment:
            g = self._generator()
            off = int(g.get_offset()) + self.captured_increment
            g.set_offset(off)
            self._mirror = (self._mirror[0], off)
