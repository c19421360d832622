"""Screens: 1-D / 2-D profile diagnostics with the reference's interface.

Mirrors ``mentflow/diagnostics/diagnostics.py`` (``Histogram1D`` :71-131, ``Histogram2D``
:134-201, noise :50-68, ``Projection`` :204-211): same constructor arguments, the same public
buffers (``edges``, ``coords``, ``resolution``, ``bandwidth``) and the same mutable flags
(``kde``, ``noise``) that callers toggle.  The arithmetic is the CUDA library's: a diagnostic
is *data* (projection axis + bin geometry) consumed by the fused projection kernels; calling
it directly launches the same kernel with a single screen.
"""
from typing import Iterable, Optional, Tuple, Union

import torch

from .. import ops
from ..utils import coords_from_edges


def _uniform_geometry(edges: torch.Tensor, what: str) -> Tuple[float, float, float]:
    """(first centre, accurate centre spacing, reference delta).  Centres are computed in the
    reference's fp32 arithmetic; the spacing used to place deposits is the float64 mean
    spacing of those centres, while ``delta`` = fp32 c[1]-c[0] is what the reference multiplies
    by in its normalisation (it carries up to 1e-5 relative rounding error, which would shift
    far bins by 1e-3 bin widths if it were used as the grid pitch).  Raises if the bins are not
    equally spaced (the KDE kernel deposits on a regular grid)."""
    e = edges.detach().to("cpu", torch.float32)
    if e.ndim != 1 or e.numel() < 3:
        raise ValueError(f"{what}: need at least two bins")
    c = coords_from_edges(e)
    w = torch.diff(c)
    delta = float(c[1] - c[0])
    if delta <= 0 or float((w - delta).abs().max()) > 1.0e-4 * abs(delta):
        raise NotImplementedError(f"{what}: KDE screens need equally spaced bin edges")
    spacing = float((c[-1].double() - c[0].double()) / (c.numel() - 1))
    return float(c[0]), spacing, delta


class Diagnostic(torch.nn.Module):
    def __init__(self, device=None, seed: Optional[int] = None, ndim: Optional[int] = None) -> None:
        super().__init__()
        self.device = device
        self.seed = seed
        self.ndim = ndim

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class Histogram(Diagnostic):
    def __init__(self, noise: bool = False, noise_scale: float = 0.0, noise_type: str = "gaussian", **kws):
        super().__init__(**kws)
        self.noise = noise
        self.noise_scale = noise_scale
        self.noise_type = noise_type

    def set_noise(self, setting: bool) -> None:
        self.noise = setting

    def project(self, x):
        raise NotImplementedError

    def bin(self, x_proj):
        raise NotImplementedError

    def apply_noise(self, hist: torch.Tensor) -> torch.Tensor:
        """Multiplicative noise h (1 + scale * xi), xi drawn from a generator that is re-seeded
        on every call, clamped at zero (diagnostics.py:53-67; only used when generating data)."""
        if not (self.noise and self.noise_scale > 0.0):
            return hist
        rng = torch.Generator(device=hist.device)
        if self.seed is not None:
            rng.manual_seed(self.seed)
        rows = hist.shape[0]
        if self.noise_type == "uniform":
            frac = torch.rand(rows, generator=rng, device=hist.device) * (2.0 * self.noise_scale)
        elif self.noise_type == "gaussian":
            frac = torch.randn(rows, generator=rng, device=hist.device) * self.noise_scale
        else:
            frac = torch.zeros(hist.shape, dtype=torch.float32, device=hist.device)
        return torch.clamp(hist * (1.0 + frac), 0.0, None)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.apply_noise(self.bin(self.project(x)))


class Histogram1D(Histogram):
    def __init__(self, edges: torch.Tensor, bandwidth: Optional[float] = None, axis: int = 0,
                 direction: Optional[torch.Tensor] = None, kde: bool = True, **kws) -> None:
        super().__init__(**kws)
        self.axis = axis
        self.kde = kde
        self.ndim = 1
        self.direction = None if direction is None else direction / torch.norm(direction)
        self.bandwidth_rel = 0.5 if bandwidth is None else float(bandwidth)
        edges = edges.to(torch.float32)
        self.register_buffer("edges", edges)
        self.register_buffer("coords", coords_from_edges(edges))
        self.register_buffer("resolution", edges[1] - edges[0])
        self.register_buffer("bandwidth", self.bandwidth_rel * self.resolution)
        self._geom = None  # (c0, delta, sigma) python floats, filled on first use

    # -- data consumed by the fused kernels ------------------------------------------------
    @property
    def nbins(self) -> int:
        return self.edges.shape[0] - 1

    def geometry(self) -> Tuple[float, float, float, float]:
        """(c0, spacing, sigma, delta): one geometry record of the C ABI."""
        key = (self.edges.data_ptr(), self.edges._version, self.bandwidth.data_ptr(), self.bandwidth._version)
        if self._geom is None or self._geom[0] != key:     # recomputed when the public buffers change
            c0, spacing, delta = _uniform_geometry(self.edges, "Histogram1D")
            self._geom = (key, (c0, spacing, float(self.bandwidth.detach().cpu()), delta))
        return self._geom[1]

    def projection_vector(self, matrix: Optional[torch.Tensor], ndim: int, device) -> torch.Tensor:
        """Row vector w with x_proj = x . w for particles x *before* the linear map M:
        row ``axis`` of M, or M^T d_hat when a direction is set."""
        if matrix is None:
            matrix = torch.eye(ndim, dtype=torch.float32, device=device)
        matrix = matrix.to(device=device, dtype=torch.float32)
        if self.direction is None:
            return matrix[self.axis]
        return self.direction.to(device=device, dtype=torch.float32) @ matrix

    # -- reference interface ------------------------------------------------------------------
    def project(self, x: torch.Tensor) -> torch.Tensor:
        if self.direction is None:
            return x[:, self.axis]
        return torch.sum(x * self.direction.to(x.device), dim=1)

    def bin(self, x_proj: torch.Tensor) -> torch.Tensor:
        from ..simulate.simulate import profiles_1d
        one = torch.ones((1, 1), dtype=torch.float32, device=x_proj.device)
        return profiles_1d(x_proj.reshape(-1, 1), one, [self], kde=self.kde)[0]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ..simulate.simulate import profiles_1d
        w = self.projection_vector(None, x.shape[1], x.device)[None, :]
        return self.apply_noise(profiles_1d(x, w, [self], kde=self.kde)[0])

    def to(self, device):
        self.device = device
        if self.direction is not None:
            self.direction = self.direction.to(device)
        return super().to(device)


class Histogram2D(Histogram):
    def __init__(self, axis: Iterable[int], edges: Iterable[torch.Tensor],
                 bandwidth: Iterable[Optional[float]] = (None, None), kde: bool = True, **kws) -> None:
        super().__init__(**kws)
        self.axis = tuple(axis)
        self.kde = kde
        self.ndim = 2
        bw = [0.5 if b is None else float(b) for b in bandwidth]
        self.bandwidth_rel = (bw[0], bw[1])
        ex, ey = (e.to(torch.float32) for e in edges)
        self.register_buffer("edges_x", ex)
        self.register_buffer("edges_y", ey)
        self.register_buffer("coords_x", coords_from_edges(ex))
        self.register_buffer("coords_y", coords_from_edges(ey))
        self.register_buffer("resolution_x", ex[1] - ex[0])
        self.register_buffer("resolution_y", ey[1] - ey[0])
        self.register_buffer("bandwidth_x", bw[0] * self.resolution_x)
        self.register_buffer("bandwidth_y", bw[1] * self.resolution_y)
        self._geom = None

    @property
    def edges(self):
        return (self.edges_x, self.edges_y)

    @property
    def shape(self) -> Tuple[int, int]:
        return (self.edges_x.shape[0] - 1, self.edges_y.shape[0] - 1)

    def geometry(self):
        key = tuple((t.data_ptr(), t._version) for t in (self.edges_x, self.edges_y, self.bandwidth_x, self.bandwidth_y))
        if self._geom is None or self._geom[0] != key:     # recomputed when the public buffers change
            cx, sx, dx = _uniform_geometry(self.edges_x, "Histogram2D (x)")
            cy, sy, dy = _uniform_geometry(self.edges_y, "Histogram2D (y)")
            self._geom = (key, ((cx, sx, float(self.bandwidth_x.detach().cpu()), dx),
                                (cy, sy, float(self.bandwidth_y.detach().cpu()), dy)))
        return self._geom[1]

    def projection_vectors(self, matrix: Optional[torch.Tensor], ndim: int, device) -> torch.Tensor:
        if matrix is None:
            matrix = torch.eye(ndim, dtype=torch.float32, device=device)
        matrix = matrix.to(device=device, dtype=torch.float32)
        return matrix[list(self.axis)]

    def project(self, x: torch.Tensor) -> torch.Tensor:
        return x[:, list(self.axis)]

    def bin(self, x_proj: torch.Tensor) -> torch.Tensor:
        from ..simulate.simulate import profiles_2d
        eye = torch.eye(2, dtype=torch.float32, device=x_proj.device)[None]
        return profiles_2d(x_proj, eye, [self], kde=self.kde)[0]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ..simulate.simulate import profiles_2d
        w = self.projection_vectors(None, x.shape[1], x.device)[None]
        return self.apply_noise(profiles_2d(x, w, [self], kde=self.kde)[0])

    def to(self, device):
        self.device = device
        return super().to(device)


class Projection(Diagnostic):
    """Projects points onto axes, no density estimation (diagnostics.py:204-211)."""

    def __init__(self, axis: Union[int, Tuple[int]], **kws) -> None:
        super().__init__(**kws)
        self.axis = axis

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x[:, self.axis]
