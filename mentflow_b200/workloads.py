"""Synthetic measurement set-ups with the parameter values of the reference's experiment
scripts.  Plain tensors only -- usable by tests, bench and the
oracle alike."""
import math
from typing import Dict

import numpy as np
import torch


def isotropic_1d(ndim: int = 6, num: int = 100, bins: int = 64, xmax: float = 3.5, seed: int = 0) -> Dict:
    """rec_nd_1d: `num` random unit directions; M = identity with row 0 replaced by the
    direction; one Histogram1D(axis=0) with `bins` bins on [-xmax, xmax]
    (experiments/rec_nd_1d/setup.py:28-69, run_gmm.sh:17-23)."""
    rng = torch.Generator().manual_seed(seed)
    dirs = torch.randn((num, ndim), generator=rng)
    dirs = dirs / torch.norm(dirs, dim=1)[:, None]
    mats = []
    for v in dirs:
        m = torch.eye(ndim)
        m[0, :] = v
        mats.append(m.float())
    return {"matrices": mats, "edges": torch.linspace(-xmax, xmax, bins + 1), "ndim": ndim}


def rotations_2d(num: int = 7, bins: int = 85, xmax: float = 3.5, max_angle: float = math.pi) -> Dict:
    """rec_2d/linear: rotations at linspace(0, max_angle, num, endpoint=False)
    (experiments/rec_2d/linear/setup.py:29-43, config/base.yaml:17-23)."""
    mats = []
    for a in np.linspace(0.0, max_angle, num, endpoint=False):
        c, s = math.cos(a), math.sin(a)
        mats.append(torch.tensor([[c, s], [-s, c]], dtype=torch.float32))
    return {"matrices": mats, "edges": torch.linspace(-xmax, xmax, bins + 1), "ndim": 2}


def corner_2d(ndim: int = 6, bins: int = 85, xmax: float = 3.5) -> Dict:
    """rec_nd_2d 'corner' optics: every axis pair (j < i) is swapped onto the measured axes
    (0, 2) by a permutation matrix (experiments/rec_nd_2d/setup.py:38-53)."""
    mats = []
    for i in range(ndim):
        for j in range(i):
            m = torch.eye(ndim)
            for k, l in zip((0, 2), (j, i)):
                p = torch.eye(ndim)
                p[k, k] = p[l, l] = 0.0
                p[k, l] = p[l, k] = 1.0
                m = p @ m
            mats.append(m.float())
    e = torch.linspace(-xmax, xmax, bins + 1)
    return {"matrices": mats, "edges": (e, e.clone()), "axis": (0, 2), "ndim": ndim}


def gaussian_mixture(n: int, ndim: int = 6, modes: int = 7, seed: int = 0, device="cpu") -> torch.Tensor:
    """Synthetic n-D Gaussian-mixture particles (stand-in for the reference's ground-truth
    samplers, mentflow/distributions; values only need to spread over the screens)."""
    rng = torch.Generator().manual_seed(seed)
    centres = torch.randn(modes, ndim, generator=rng) * 1.2
    which = torch.randint(0, modes, (n,), generator=rng)
    x = centres[which] + 0.45 * torch.randn(n, ndim, generator=rng)
    return x.float().to(device)
