"""Grid sampling of a density (mentflow/sample.py): evaluate the density on the cell centres of
a regular grid, draw cells from the resulting pmf, jitter uniformly inside each cell.

The reference materialises the res^D x D grid points, calls ``prob_func`` on them and uses
``torch.multinomial`` (limited to 2^24 categories).  Here the density of a ``MENT`` model is
evaluated straight from the grid index, the pmf is turned into a float64 CDF by a scan kernel
and particles are drawn by inverse-CDF search with an in-kernel Philox stream -- no category
limit.  The random stream necessarily differs from torch's CPU generator; parity is statistical.
"""
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import ops
from .utils import coords_from_edges, get_grid_points


def _draw_seed() -> int:
    """A fresh 63-bit seed from torch's global generator, so torch.manual_seed reproduces runs."""
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def sample_hist(hist: torch.Tensor, edges: Sequence[torch.Tensor], size: int, noise: float = 0.0,
                device=None, seed: Optional[int] = None) -> torch.Tensor:
    """Particles from a histogram (sample.py:34-57).  ``edges`` must be equally spaced."""
    if hist.ndim == 1 and isinstance(edges, torch.Tensor):
        edges = [edges]
    shape = list(hist.shape)
    first = [float(e[0]) for e in edges]
    cell = [float((e[-1].double() - e[0].double()) / (len(e) - 1)) for e in edges]
    x = ops.cdf_sample(hist, shape, first, cell, int(size), _draw_seed() if seed is None else seed,
                       jitter=bool(noise))
    return torch.squeeze(x)


class GridSampler:
    def __init__(self, limits: List[Tuple[float]], shape: Tuple[int], noise: float = 0.0, device=None,
                 store: bool = True) -> None:
        self.device = device
        self.shape = tuple(int(s) for s in shape)
        self.limits = limits
        self.ndim = len(limits)
        self.noise = noise
        self.store = store
        self.edges = [torch.linspace(limits[a][0], limits[a][1], self.shape[a] + 1) for a in range(self.ndim)]
        self.coords = [coords_from_edges(e) for e in self.edges]
        self.points = None
        self.calls = 0

    def send(self, x: torch.Tensor) -> torch.Tensor:
        return x.type(torch.float32).to(self.device)

    # geometry handed to the kernels (python floats, float64 arithmetic)
    def cell_sizes(self):
        return [(float(self.limits[a][1]) - float(self.limits[a][0])) / self.shape[a] for a in range(self.ndim)]

    def first_centres(self):
        return [float(self.limits[a][0]) + 0.5 * c for a, c in enumerate(self.cell_sizes())]

    def get_grid_points(self) -> torch.Tensor:
        if self.points is not None:
            return self.points
        points = self.send(get_grid_points(*self.coords))
        if self.store:
            self.points = points
        return points

    supports_offset = True    # __call__ takes the index of the first particle to draw (sharded MENT)

    def __call__(self, prob_func: Callable, size: int, seed: Optional[int] = None, offset: int = 0) -> torch.Tensor:
        """``size`` particles; particle s of the call uses Philox counter ``offset + s``, so ranks that pass the
        same seed (torch.manual_seed on every rank) and disjoint offsets draw disjoint slices of one stream."""
        owner = getattr(prob_func, "__self__", None)
        if owner is not None and hasattr(owner, "prob_on_grid") and getattr(prob_func, "__name__", "") == "prob":
            prob = owner.prob_on_grid(self)           # density from the grid index, nothing materialised
        else:
            prob = prob_func(self.get_grid_points())
        self.calls += 1
        first = [float(self.limits[a][0]) for a in range(self.ndim)]
        x = ops.cdf_sample(prob, list(self.shape), first, self.cell_sizes(), int(size),
                           _draw_seed() if seed is None else seed, offset=int(offset), jitter=bool(self.noise))
        return x

    def to(self, device):
        self.device = device
        self.edges = [self.send(e) for e in self.edges]
        self.coords = [self.send(c) for c in self.coords]
        self.points = None
        return self
