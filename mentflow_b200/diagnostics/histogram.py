"""Functional KDE interface of ``mentflow/diagnostics/histogram.py`` (:77-101) on the CUDA
kernels: ``kde_histogram_1d(x, bins, bandwidth)`` / ``kde_histogram_2d(x, y, bins, bandwidth)``
with *absolute* bandwidths, as in the reference."""
from typing import Iterable

import torch

from .. import ops


def _geom_rows(edges_list, sigmas, device):
    from .diagnostics import _uniform_geometry
    rows, ratio = [], 0.0
    for e, s in zip(edges_list, sigmas):
        c0, spacing, delta = _uniform_geometry(e, "kde_histogram")
        s = float(s)
        rows.append([c0, spacing, s, delta, 0.0, 0.0, 0.0, 0.0])
        ratio = max(ratio, s / spacing)
    return torch.tensor(rows, dtype=torch.float32, device=device), ratio


def kde_histogram_1d(x: torch.Tensor, bins: torch.Tensor, bandwidth: float = 1.0, epsilon: float = 1.0e-10):
    if epsilon != 1.0e-10:
        raise NotImplementedError("the normalisation pad is fixed at 1e-10 in the kernels")
    geom, ratio = _geom_rows([bins], [bandwidth], x.device)
    one = torch.ones((1, 1), dtype=torch.float32, device=x.device)
    return ops.project_kde1d(x.reshape(-1, 1), one, geom, ratio, bins.shape[0] - 1)[0]


def kde_histogram_2d(x: torch.Tensor, y: torch.Tensor, bins: Iterable[torch.Tensor],
                     bandwidth: Iterable[float] = (1.0, 1.0), epsilon: float = 1.0e-10):
    if epsilon != 1.0e-10:
        raise NotImplementedError("the normalisation pad is fixed at 1e-10 in the kernels")
    bins = list(bins)
    geom, ratio = _geom_rows(bins, list(bandwidth), x.device)
    xy = torch.stack([x, y], dim=1)
    eye = torch.eye(2, dtype=torch.float32, device=x.device)[None]
    return ops.project_kde2d(xy, eye, geom.reshape(1, 2, -1), ratio, bins[0].shape[0] - 1, bins[1].shape[0] - 1)[0]
