"""Prior distributions (mentflow/prior.py)."""
import math

import torch


class Gaussian:
    """Isotropic N(0, scale^2 I) (prior.py:4-26).  ``log_prob`` is the closed form of
    MultivariateNormal(0, scale^2 I).log_prob; on the training path it is never called:
    the Monte-Carlo entropy kernel folds it in."""

    def __init__(self, ndim: int = 2, scale: float = 1.0, device=None) -> None:
        self.ndim = ndim
        self.scale = scale
        self.device = device

    def to(self, device):
        self.device = device
        return self

    @property
    def log_norm(self) -> float:
        return -self.ndim * math.log(self.scale) - 0.5 * self.ndim * math.log(2.0 * math.pi)

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        return -0.5 * torch.sum(x * x, dim=1) / (self.scale ** 2) + self.log_norm


class Uniform:
    """Constant density on a cube of side ``scale`` (prior.py:29-45; the reference version
    raises NameError because it never imports numpy -- fixed here)."""

    def __init__(self, ndim: int = 2, scale: float = 100.0, device=None) -> None:
        self.scale = scale
        self.ndim = ndim
        self.volume = scale ** ndim
        self.device = device

    def to(self, device):
        self.device = device
        return self

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        return torch.full((x.shape[0],), -math.log(self.volume), dtype=torch.float32, device=x.device)
