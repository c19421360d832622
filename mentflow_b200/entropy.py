"""Negative-entropy estimators (mentflow/entropy.py).

``MonteCarloEntropyEstimator`` is one fused, deterministic double-precision reduction over
the particles (sum log q and sum |x|^2 in one pass) instead of a MultivariateNormal
triangular solve plus two means."""
import math
from typing import Any

import numpy as np
import torch

from . import ops
from .prior import Gaussian


_coef_cache = {}


def _coef(a: float, b: float, device) -> torch.Tensor:
    """Device-resident float64 pair (a, b): cached so that a step issues no host-to-device copy."""
    key = (a, b, str(device))
    t = _coef_cache.get(key)
    if t is None:
        if len(_coef_cache) > 64:
            _coef_cache.clear()
        t = _coef_cache[key] = torch.tensor([a, b], dtype=torch.float64, device=device)
    return t


class _MCEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, log_prob, inv_s2, log_norm, reducer):
        m = ops.moments(x.detach(), log_prob.detach(), with_cov=False)
        n = float(x.shape[0])
        if reducer is not None:
            n = reducer(m, n)
        ctx.save_for_backward(x)
        ctx.inv_s2, ctx.n = inv_s2, n
        # H = mean(log q) - mean(log prior),  log prior = -0.5 |x|^2 / s^2 + log_norm
        coef = _coef(1.0 / n, 0.5 * inv_s2 / n, x.device)
        return ((m * coef).sum() - log_norm).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        gl = (g / ctx.n).expand(x.shape[0])
        gx = x * (g * (ctx.inv_s2 / ctx.n)) if ctx.inv_s2 != 0.0 else None
        return gx, gl, None, None, None


class EntropyEstimator(torch.nn.Module):
    """Estimates negative entropy from samples and/or log probability (entropy.py:8-15)."""

    def __init__(self, prior: Any = None) -> None:
        super().__init__()
        self.prior = prior
        self.reducer = None

    def forward(self, x: torch.Tensor, log_prob: torch.Tensor = None) -> torch.Tensor:
        raise NotImplementedError


class EmptyEntropyEstimator(EntropyEstimator):
    def forward(self, x: torch.Tensor, log_prob: torch.Tensor = None) -> torch.Tensor:
        return 0.0


class MonteCarloEntropyEstimator(EntropyEstimator):
    """H = mean(log q) - mean(log prior(x))  (entropy.py:53-62)."""

    def forward(self, x: torch.Tensor, log_prob: torch.Tensor) -> torch.Tensor:
        if self.prior is None:
            return _MCEntropy.apply(x, log_prob, 0.0, 0.0, self.reducer)
        if isinstance(self.prior, Gaussian):
            return _MCEntropy.apply(x, log_prob, 1.0 / self.prior.scale ** 2, self.prior.log_norm, self.reducer)
        # arbitrary prior object: same estimator, prior evaluated by the caller's code
        h = _MCEntropy.apply(x, log_prob, 0.0, 0.0, self.reducer)
        return h - torch.mean(self.prior.log_prob(x))


class CovarianceEntropyEstimator(EntropyEstimator):
    """-3 log(2 pi e) - log(sqrt(det cov) + pad)  (entropy.py:27-38; the constant is the
    reference's, hard-coded for six dimensions).  The covariance comes from the moments kernel;
    no gradient is propagated (the reference's configs use the Monte-Carlo estimator)."""

    def __init__(self, prior: Any = None, pad: float = 1.0e-12) -> None:
        if prior is not None:
            raise ValueError("This class cannot estimate relative entropy (prior != None).")
        super().__init__(prior=prior)
        self.pad = pad

    def forward(self, x: torch.Tensor, log_prob: torch.Tensor = None) -> torch.Tensor:
        n, d = x.shape
        m = ops.moments(x.detach(), None, with_cov=True)
        mean = m[2:2 + d] / n
        second = m[2 + d:].reshape(d, d) / n
        cov = (second - torch.outer(mean, mean)) * (n / (n - 1.0))
        eps = torch.sqrt(torch.det(cov))
        h = -3.0 * np.log(2.0 * np.pi * np.e) - torch.log(eps + self.pad)
        return h.to(torch.float32)


class KNNEntropyEstimator(EntropyEstimator):
    def __init__(self, prior: Any = None, k: int = 5) -> None:
        if prior is not None:
            raise ValueError("This class cannot estimate relative entropy (prior != None).")
        super().__init__(prior=prior)
        self.k = k

    def forward(self, x, log_prob=None):
        raise NotImplementedError
