"""Prior distributions (API of mentflow/prior.py: ``Gaussian``, ``Uniform`` with ``log_prob`` and ``to``)."""
import math

import torch


class _Prior:
    """Shared bookkeeping: dimension, scale and the device the caller asked for (kept for API
    compatibility; ``log_prob`` follows the device of its argument)."""

    def __init__(self, ndim: int, scale: float, device) -> None:
        self.ndim, self.scale, self.device = ndim, scale, device

    def to(self, device):
        self.device = device
        return self


class Gaussian(_Prior):
    """Isotropic N(0, scale^2 I) (prior.py:4-26).  ``log_prob`` is the closed form of what the reference
    gets from ``MultivariateNormal(0, scale^2 I)``; on the training path it is never called: the
    Monte-Carlo entropy kernel folds it in (``entropy.MonteCarloEntropyEstimator``)."""

    def __init__(self, ndim: int = 2, scale: float = 1.0, device=None) -> None:
        super().__init__(ndim, scale, device)

    @property
    def log_norm(self) -> float:
        """log of the normalisation constant: -D log(scale) - D/2 log(2 pi)."""
        return -self.ndim * (math.log(self.scale) + 0.5 * math.log(2.0 * math.pi))

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        r2 = torch.sum(x * x, dim=1)
        return self.log_norm - 0.5 * r2 / (self.scale * self.scale)


class Uniform(_Prior):
    """Constant density on a cube of side ``scale`` (prior.py:29-45; the reference version raises
    NameError because it never imports numpy -- fixed here)."""

    def __init__(self, ndim: int = 2, scale: float = 100.0, device=None) -> None:
        super().__init__(ndim, scale, device)
        self.volume = scale ** ndim

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        value = -math.log(self.volume)
        return torch.full((x.shape[0],), value, dtype=torch.float32, device=x.device)
