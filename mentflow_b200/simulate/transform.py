"""Transfer maps between the reconstruction point and the screens.

API of ``mentflow/simulate/transform.py`` (``Transform``, ``LinearTransform``,
``CompositeTransform``, ``rotation_matrix``).  On the hot path a ``LinearTransform`` is never
*applied*: ``simulate.forward`` reads ``.matrix`` and folds the measured row into the fused
projection kernel.  ``forward``/``inverse`` exist for callers that want the full (N, D)
image (classical MENT's integration mode, notebooks) and are plain matmuls.
"""
import math

import torch
import torch.nn as nn


def rotation_matrix(angle: float) -> torch.Tensor:
    """2x2 phase-space rotation [[c, s], [-s, c]] in float64, like the reference (:12-15)."""
    c, s = math.cos(angle), math.sin(angle)
    return torch.tensor([[c, s], [-s, c]], dtype=torch.float64)


class Transform(nn.Module):
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def inverse(self, u: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class LinearTransform(Transform):
    """u = x M^T (simulate/transform.py:58-75)."""

    def __init__(self, matrix: torch.Tensor) -> None:
        super().__init__()
        self.set_matrix(matrix)

    def set_matrix(self, matrix: torch.Tensor) -> None:
        self.matrix = matrix
        self.matrix_inv = torch.linalg.inv(matrix)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x @ self.matrix.T

    def inverse(self, u: torch.Tensor) -> torch.Tensor:
        return u @ self.matrix_inv.T

    def to(self, device):
        # the reference forgets matrix_inv here; both move.
        self.matrix = self.matrix.to(device)
        self.matrix_inv = self.matrix_inv.to(device)
        return self


class CompositeTransform(Transform):
    """Chain of transforms (simulate/transform.py:35-55); a chain of LinearTransforms is
    collapsed to one matrix by ``simulate.forward``."""

    def __init__(self, *transforms) -> None:
        super().__init__()
        self.transforms = nn.ModuleList(transforms)

    def forward(self, x):
        for t in self.transforms:
            x = t(x)
        return x

    def inverse(self, u):
        for t in reversed(self.transforms):
            u = t.inverse(u)
        return u

    def to(self, device):
        for t in self.transforms:
            t.to(device)
        return self

    def as_matrix(self):
        """Product matrix if every stage is linear, else None."""
        m = None
        for t in self.transforms:
            if not isinstance(t, LinearTransform):
                return None
            m = t.matrix if m is None else t.matrix @ m
        return m
