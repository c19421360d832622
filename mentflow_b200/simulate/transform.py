"""Transfer maps between the reconstruction point and the screens.

API of ``mentflow/simulate/transform.py`` (``Transform``, ``LinearTransform``,
``CompositeTransform``, ``MultipoleTransform``, ``ProjectionTransform``, ``rotation_matrix``).  On the hot path a ``LinearTransform`` is never
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


# Re(z^(n-1)), Im(z^(n-1)) of z = x + i y for the multipole orders the reference accepts, in the
# reference's own expanded form (simulate/transform.py:118-134) so that results agree to the last bit
_KICK_POLYNOMIALS = {
    3: lambda x, y: (x ** 2 - y ** 2, 2.0 * x * y),
    4: lambda x, y: (x ** 3 - 3.0 * y ** 2 * x, -(y ** 3) + 3.0 * x ** 2 * y),
    5: lambda x, y: (x ** 4 - 6.0 * x ** 2 * y ** 2 + y ** 4, 4.0 * x ** 3 * y - 4.0 * x * y ** 3),
}


class MultipoleTransform(Transform):
    """Thin multipole kick (simulate/transform.py:78-146).

    ``simulate.forward`` folds a chain ``linear* -> multipole -> linear*`` into the fused projection
    kernel (``multipole_terms``); ``forward`` is the stand-alone (N, D) image with the reference's
    semantics, quirks included: only orders 3, 4, 5 are accepted (the reference's ``if`` ladder falls
    into its ``else: raise`` for orders 1 and 2, :118-134), and the normal kick sets column 3 to
    ``column 1 + k Im(z^(n-1))`` -- from column 1, not column 3 (:144).
    """

    def __init__(self, order: int, strength: float, skew: bool = False) -> None:
        super().__init__()
        self.order, self.strength, self.skew = order, strength, skew

    def coefficient(self) -> float:
        """Integrated strength / (order - 1)!"""
        return float(self.strength) / math.factorial(int(self.order) - 1)

    def forward(self, particles: torch.Tensor) -> torch.Tensor:
        poly = _KICK_POLYNOMIALS.get(self.order)
        if poly is None:
            raise ValueError("MPS-compatible MultipoleTransform requires order <= 5.")
        four_d = particles.shape[1] > 2
        pos_x, mom_x = particles[:, 0], particles[:, 1]
        pos_y = particles[:, 2] if four_d else 0.0 * pos_x
        re, im = poly(pos_x, pos_y)
        k = self.coefficient()
        out = particles.clone()
        if self.skew:
            out[:, 1] = mom_x + k * im
            if four_d:
                out[:, 3] = particles[:, 3] + k * re
        else:
            out[:, 1] = mom_x - k * re
            if four_d:
                out[:, 3] = mom_x + k * im        # sic: the reference adds to column 1 here
        return out

    def inverse(self, u: torch.Tensor) -> torch.Tensor:
        """Momentum reversal, kick, momentum reversal (:145-146).  Unlike the reference, which flips
        the momenta of its argument in place, the input is left untouched."""
        return reverse_momentum(self.forward(reverse_momentum(u.clone())))


def reverse_momentum(x: torch.Tensor) -> torch.Tensor:
    """Flip the sign of every momentum column (1, 3, ...), in place like simulate/transform.py:18-21."""
    x[:, 1::2].neg_()
    return x


class ProjectionTransform(Transform):
    """(N, 1) projection on a direction, normalised at construction (simulate/transform.py:149-156)."""

    def __init__(self, direction: torch.Tensor) -> None:
        super().__init__()
        self.direction = direction / torch.norm(direction)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.sum(x * self.direction.to(x.device), dim=1)[:, None]
