"""TEST INFRASTRUCTURE -- torch-CPU restatement of the reference's non-flow hot path.

Every function cites the reference lines it restates (paths relative to
``/root/reference/mentflow``).  The functions are deliberately written in the *dense* form
the reference uses (an (N, B) kernel matrix per projection, one matmul per transform) so
that timing them is a fair "port" CPU baseline; ``chunk=`` bounds memory for large-N tests
without changing the arithmetic of each row.

Pinned against the reference's own code by ``tests/test_oracle_golden.py`` using the
vectors in ``tests/golden`` (made by ``oracle/make_goldens.py`` from the real package).
Nothing in ``mentflow_b200`` imports this file.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

KDE_EPS = 1.0e-10   # diagnostics/histogram.py:15,51
KL_PAD = 1.0e-12    # loss.py:15


# --------------------------------------------------------------------------------------
# screens (diagnostics) described as plain data
# --------------------------------------------------------------------------------------
@dataclass
class Screen1D:
    """diagnostics/diagnostics.py:71-131 (Histogram1D) as data."""
    edges: torch.Tensor                 # (B+1,) fp32
    bandwidth: float = 0.5              # relative to the bin width (:108-114)
    axis: int = 0
    direction: Optional[torch.Tensor] = None

    @property
    def sigma(self) -> torch.Tensor:    # :113-114  bandwidth * (edges[1] - edges[0])
        return self.bandwidth * (self.edges[1] - self.edges[0])


@dataclass
class Screen2D:
    """diagnostics/diagnostics.py:134-201 (Histogram2D) as data."""
    axis: Tuple[int, int]
    edges_x: torch.Tensor
    edges_y: torch.Tensor
    bandwidth: Tuple[float, float] = (0.5, 0.5)

    @property
    def sigma(self):
        return (self.bandwidth[0] * (self.edges_x[1] - self.edges_x[0]),
                self.bandwidth[1] * (self.edges_y[1] - self.edges_y[0]))


def centres(edges: torch.Tensor) -> torch.Tensor:
    """utils/grid.py:5-6."""
    return 0.5 * (edges[:-1] + edges[1:])


# --------------------------------------------------------------------------------------
# transforms + projection
# --------------------------------------------------------------------------------------
def linear_map(x: torch.Tensor, matrix: torch.Tensor) -> torch.Tensor:
    """simulate/transform.py:67-68: u = x M^T (full D x D product, as the reference does)."""
    return torch.matmul(x, matrix.T)


def multipole_kick(x: torch.Tensor, order: int, strength: float, skew: bool = False) -> torch.Tensor:
    """simulate/transform.py:78-146 (MultipoleTransform.forward), quirks included: the `if / if /
    if-elif-else` ladder (:118-134) raises for every order outside 3..5 (orders 1 and 2 fall into
    the final `else`), and the normal (non-skew) kick writes U[:, 3] = X[:, 1] + k Im(z^(n-1)),
    i.e. from column 1, not column 3 (:144).  k = strength / (order - 1)!."""
    if order not in (3, 4, 5):
        raise ValueError("MPS-compatible MultipoleTransform requires order <= 5.")
    u = x.clone()
    xx = x[:, 0]
    yy = x[:, 2] if x.shape[1] > 2 else 0.0 * x[:, 0]
    if order == 3:
        zr, zi = xx ** 2 - yy ** 2, 2.0 * xx * yy
    elif order == 4:
        zr, zi = xx ** 3 - 3.0 * yy ** 2 * xx, -(yy ** 3) + 3.0 * xx ** 2 * yy
    else:
        zr, zi = xx ** 4 - 6.0 * xx ** 2 * yy ** 2 + yy ** 4, 4.0 * xx ** 3 * yy - 4.0 * xx * yy ** 3
    k = strength / math.factorial(order - 1)
    if skew:
        u[:, 1] = x[:, 1] + k * zi
        if x.shape[1] > 2:
            u[:, 3] = x[:, 3] + k * zr
    else:
        u[:, 1] = x[:, 1] - k * zr
        if x.shape[1] > 2:
            u[:, 3] = x[:, 1] + k * zi
    return u


def projection_transform(x: torch.Tensor, direction: torch.Tensor) -> torch.Tensor:
    """simulate/transform.py:149-156: (N, 1) projection on the normalised direction."""
    d = direction / torch.norm(direction)
    return torch.sum(x * d, dim=1)[:, None]


def project_1d(u: torch.Tensor, axis: int = 0, direction: Optional[torch.Tensor] = None):
    """diagnostics/diagnostics.py:116-122 (direction is normalised at construction, :104-106)."""
    if direction is None:
        return u[:, axis]
    d = direction / torch.norm(direction)
    return torch.sum(u * d, dim=1)


# --------------------------------------------------------------------------------------
# differentiable KDE profiles
# --------------------------------------------------------------------------------------
def _kernel_matrix(values: torch.Tensor, coords: torch.Tensor, sigma) -> torch.Tensor:
    """diagnostics/histogram.py:37-38: exp(-0.5 ((v - c)/sigma)^2), shape (N, B)."""
    resid = values[:, None] - coords[None, :]
    return torch.exp(-0.5 * (resid / sigma).pow(2))


def kde_profile_1d(u: torch.Tensor, edges: torch.Tensor, sigma, chunk: Optional[int] = None):
    """diagnostics/histogram.py:77-86 -> :11-44.

    p_b = mean_n K_nb ;  p <- p / (sum_b p_b * delta + 1e-10), delta = c[1]-c[0].
    """
    c = centres(edges)
    n = u.shape[0]
    if chunk is None or n <= chunk:
        prob = torch.mean(_kernel_matrix(u, c, sigma), dim=0)
    else:
        acc = torch.zeros_like(c)
        for s in range(0, n, chunk):
            acc = acc + _kernel_matrix(u[s:s + chunk], c, sigma).sum(dim=0)
        prob = acc / n
    delta = c[1] - c[0]
    return prob / (torch.sum(prob * delta) + KDE_EPS)


def kde_sums_1d(u: torch.Tensor, edges: torch.Tensor, sigma, chunk: int = 65536):
    """Unnormalised S_b = sum_n K_nb in float64 (the quantity the CUDA kernel reduces and
    the ranks all-reduce).  Not a reference function: used for property tests."""
    c = centres(edges).double()
    acc = torch.zeros_like(c)
    for s in range(0, u.shape[0], chunk):
        acc += _kernel_matrix(u[s:s + chunk].double(), c, float(sigma)).sum(dim=0)
    return acc


def kde_profile_2d(ux, uy, edges_x, edges_y, sigma_x, sigma_y, chunk: Optional[int] = None):
    """diagnostics/histogram.py:89-101 -> :47-74.  P = Kx^T Ky (no 1/N), then
    P <- P / (sum P * dx * dy + 1e-10)."""
    cx, cy = centres(edges_x), centres(edges_y)
    n = ux.shape[0]
    if chunk is None or n <= chunk:
        prob = torch.matmul(_kernel_matrix(ux, cx, sigma_x).T, _kernel_matrix(uy, cy, sigma_y))
    else:
        prob = torch.zeros(cx.shape[0], cy.shape[0], dtype=ux.dtype)
        for s in range(0, n, chunk):
            prob = prob + torch.matmul(_kernel_matrix(ux[s:s + chunk], cx, sigma_x).T,
                                       _kernel_matrix(uy[s:s + chunk], cy, sigma_y))
    dx = cx[1] - cx[0]
    dy = cy[1] - cy[0]
    return prob / (torch.sum(prob * dx * dy) + KDE_EPS)


# --------------------------------------------------------------------------------------
# hard histograms (measurement generation / kde=False evaluation)
# --------------------------------------------------------------------------------------
def bin_index(u: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    """Bin rule of torch.histogram / np.histogramdd: bin i holds
    e[i] <= u < e[i+1], the last bin is closed on the right, everything else (and NaN)
    is dropped (returned as -1)."""
    nb = edges.shape[0] - 1
    idx = torch.searchsorted(edges, u.contiguous(), right=True) - 1
    idx = torch.where(u == edges[-1], torch.full_like(idx, nb - 1), idx)
    bad = (idx < 0) | (idx >= nb) | torch.isnan(u)
    return torch.where(bad, torch.full_like(idx, -1), idx)


def hist_counts_1d(u: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    idx = bin_index(u, edges)
    idx = idx[idx >= 0]
    return torch.bincount(idx, minlength=edges.shape[0] - 1)


def hist_density_1d(u: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    """diagnostics/diagnostics.py:128-131: torch.histogram(u, edges, density=True).hist."""
    return torch.histogram(u.contiguous(), edges, density=True).hist


def density_from_counts_1d(counts: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    """counts / counts.sum() / diff(edges) in fp32 (App. B.3; valid while counts < 2^24)."""
    c = counts.to(torch.float32)
    return c / c.sum() / torch.diff(edges)


def hist_counts_2d(uv: torch.Tensor, edges_x: torch.Tensor, edges_y: torch.Tensor):
    ix = bin_index(uv[:, 0], edges_x)
    iy = bin_index(uv[:, 1], edges_y)
    ok = (ix >= 0) & (iy >= 0)
    nx, ny = edges_x.shape[0] - 1, edges_y.shape[0] - 1
    flat = torch.bincount(ix[ok] * ny + iy[ok], minlength=nx * ny)
    return flat.reshape(nx, ny)


def hist_density_2d(uv: torch.Tensor, edges_x: torch.Tensor, edges_y: torch.Tensor):
    """diagnostics/diagnostics.py:192-201: np.histogramdd(density=True) in float64 -> fp32."""
    hist, _ = np.histogramdd(uv.detach().numpy(),
                             bins=[edges_x.numpy(), edges_y.numpy()], density=True)
    return torch.from_numpy(hist).type(torch.float32)


def apply_noise(hist: torch.Tensor, noise_scale: float, noise_type: str, seed: Optional[int]):
    """diagnostics/diagnostics.py:53-67: multiplicative noise from a freshly seeded generator."""
    rng = torch.Generator()
    if seed is not None:
        rng.manual_seed(seed)
    if noise_type == "uniform":
        frac = torch.rand(hist.shape[0], generator=rng) * 2.0 * noise_scale
    elif noise_type == "gaussian":
        frac = torch.randn(hist.shape[0], generator=rng) * noise_scale
    else:
        frac = torch.zeros(hist.shape)
    return torch.clamp(hist * (1.0 + frac), 0.0, None)


# --------------------------------------------------------------------------------------
# the simulation driver
# --------------------------------------------------------------------------------------
def simulate(x: torch.Tensor, matrices: Sequence[torch.Tensor], screens: Sequence[Sequence],
             kde: bool = True, chunk: Optional[int] = None) -> List[List[torch.Tensor]]:
    """simulate/simulate.py:29-33: for every transform, u = T(x.clone()); every diagnostic
    attached to that transform produces one profile."""
    out = []
    for matrix, row in zip(matrices, screens):
        u = linear_map(x.clone(), matrix)
        profiles = []
        for scr in row:
            if isinstance(scr, Screen1D):
                up = project_1d(u, scr.axis, scr.direction)
                if kde:
                    profiles.append(kde_profile_1d(up, scr.edges, scr.sigma, chunk))
                else:
                    profiles.append(hist_density_1d(up, scr.edges))
            else:
                uv = u[:, list(scr.axis)]
                if kde:
                    sx, sy = scr.sigma
                    profiles.append(kde_profile_2d(uv[:, 0], uv[:, 1], scr.edges_x, scr.edges_y,
                                                   sx, sy, chunk))
                else:
                    profiles.append(hist_density_2d(uv, scr.edges_x, scr.edges_y))
        out.append(profiles)
    return out


# --------------------------------------------------------------------------------------
# prior, entropy, discrepancy, loss
# --------------------------------------------------------------------------------------
def gaussian_log_prob(x: torch.Tensor, scale: float) -> torch.Tensor:
    """prior.py:10-26: MultivariateNormal(0, scale^2 I).log_prob(x) -- the same torch
    distribution object the reference builds, so the rounding is identical
    (= -0.5*|x|^2/scale^2 - D*log(scale) - D/2*log(2*pi))."""
    d = x.shape[1]
    loc = torch.zeros(d, dtype=torch.float32)
    cov = (torch.eye(d) * (scale ** 2)).type(torch.float32)
    dist = torch.distributions.MultivariateNormal(loc.to(x.dtype), cov.to(x.dtype))
    return dist.log_prob(x)


def gaussian_log_prob_closed_form(x: torch.Tensor, scale: float) -> torch.Tensor:
    """The closed form the CUDA kernels evaluate."""
    d = x.shape[1]
    return -0.5 * torch.sum(x * x, dim=1) / (scale ** 2) - d * math.log(scale) \
        - 0.5 * d * math.log(2.0 * math.pi)


def mc_entropy(x: torch.Tensor, log_q: torch.Tensor, prior_scale: Optional[float]):
    """entropy.py:58-62: H = mean(log q) - mean(log prior(x))  (negative entropy)."""
    h = torch.mean(log_q)
    if prior_scale is not None:
        h = h - torch.mean(gaussian_log_prob(x, prior_scale))
    return h


def cov_entropy(x: torch.Tensor, pad: float = 1.0e-12) -> torch.Tensor:
    """entropy.py:35-38 (the constant is hard-coded for 6D in the reference)."""
    eps = torch.sqrt(torch.det(torch.cov(x.T)))
    return -3.0 * np.log(2.0 * np.pi * np.e) - torch.log(eps + pad)


def kl_div(pred: torch.Tensor, targ: torch.Tensor, pad: float = KL_PAD) -> torch.Tensor:
    """loss.py:15-17: F.kl_div(log(pred+pad), targ, 'batchmean')
    = (1/pred.shape[0]) * sum targ*(log targ - log(pred+pad)), 0*log0 = 0."""
    return torch.nn.functional.kl_div(torch.log(pred + pad), targ, reduction="batchmean")


def mae(pred, targ):
    """loss.py:7-8."""
    return torch.mean(torch.abs(pred - targ))


def mse(pred, targ):
    """loss.py:11-12."""
    return torch.mean(torch.square(pred - targ))


def mentflow_loss(x, log_q, matrices, screens, measurements, prior_scale, penalty,
                  discrepancy=kl_div, chunk: Optional[int] = None):
    """core.py:95-117: L = H + mu * (sum_k D_k / K); returns (L, H, [D_k])."""
    h = mc_entropy(x, log_q, prior_scale)
    preds = simulate(x, matrices, screens, kde=True, chunk=chunk)
    d = [discrepancy(p, m) for prow, mrow in zip(preds, measurements) for p, m in zip(prow, mrow)]
    return h + penalty * (sum(d) / len(d)), h, d


# --------------------------------------------------------------------------------------
# classical MENT
# --------------------------------------------------------------------------------------
def lagrange_interp(table: torch.Tensor, coords: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """ment.py:45-52: scipy RegularGridInterpolator(method='linear', bounds_error=False,
    fill_value=0) on the bin *centres*, evaluated in float64 (1-D tables)."""
    t = table.double().numpy()
    c = coords.double().numpy()
    v = u.double().numpy()
    idx = np.clip(np.searchsorted(c, v, side="right") - 1, 0, len(c) - 2)
    w = (v - c[idx]) / (c[idx + 1] - c[idx])
    val = t[idx] * (1.0 - w) + t[idx + 1] * w
    val = np.where((v < c[0]) | (v > c[-1]) | np.isnan(v), 0.0, val)
    return torch.from_numpy(val)


def lagrange_interp_2d(table, coords_x, coords_y, uv):
    """Same, bilinear, for 2-D tables (ment.py:36-49 with two coordinate arrays)."""
    t = table.double().numpy()
    cx, cy = coords_x.double().numpy(), coords_y.double().numpy()
    vx, vy = uv[:, 0].double().numpy(), uv[:, 1].double().numpy()
    ix = np.clip(np.searchsorted(cx, vx, side="right") - 1, 0, len(cx) - 2)
    iy = np.clip(np.searchsorted(cy, vy, side="right") - 1, 0, len(cy) - 2)
    wx = (vx - cx[ix]) / (cx[ix + 1] - cx[ix])
    wy = (vy - cy[iy]) / (cy[iy + 1] - cy[iy])
    val = (t[ix, iy] * (1 - wx) * (1 - wy) + t[ix + 1, iy] * wx * (1 - wy)
           + t[ix, iy + 1] * (1 - wx) * wy + t[ix + 1, iy + 1] * wx * wy)
    out = (vx < cx[0]) | (vx > cx[-1]) | (vy < cy[0]) | (vy > cy[-1])
    return torch.from_numpy(np.where(out, 0.0, val))


def ment_prob(x: torch.Tensor, matrices, screens, tables, prior_scale: float) -> torch.Tensor:
    """ment.py:239-249: rho(x) = exp(log prior(x)) * prod_k clamp(h_k(proj(M_k x)), 0, 1e10),
    interpolation in fp64 then cast to fp32 (:155, :233)."""
    prob = torch.ones(x.shape[0], dtype=torch.float32)
    for matrix, srow, trow in zip(matrices, screens, tables):
        u = linear_map(x, matrix)
        for scr, table in zip(srow, trow):
            if isinstance(scr, Screen1D):
                h = lagrange_interp(table, centres(scr.edges), project_1d(u, scr.axis, scr.direction))
            else:
                h = lagrange_interp_2d(table, centres(scr.edges_x), centres(scr.edges_y),
                                       u[:, list(scr.axis)])
            prob = prob * torch.clamp(h.type(torch.float32), 0.0, 1.0e10)
    return prob * torch.exp(gaussian_log_prob(x, prior_scale))


def normalize_projection(projection: torch.Tensor, bin_volume) -> torch.Tensor:
    """ment.py:184-191."""
    return projection / projection.sum() / bin_volume


def gauss_seidel_table(table: torch.Tensor, meas: torch.Tensor, pred: torch.Tensor,
                       lr: float, thresh: float) -> torch.Tensor:
    """ment.py:360-367: pred[pred<thresh]=0; where meas!=0 and pred!=0:
    h <- h * (1 + lr*(meas/pred - 1)); other entries unchanged."""
    pred = torch.where(pred < thresh, torch.zeros_like(pred), pred)
    ok = (meas != 0.0) & (pred != 0.0)
    safe = torch.where(ok, pred, torch.ones_like(pred))
    return torch.where(ok, table * (1.0 + lr * (meas / safe - 1.0)), table)


def initial_table(meas: torch.Tensor) -> torch.Tensor:
    """ment.py:179: h0 = 1[g > 0]."""
    return (meas > 0.0).float()


def grid_points(coords: Sequence[torch.Tensor]) -> torch.Tensor:
    """utils/grid.py:9-10."""
    return torch.vstack([c.ravel() for c in torch.meshgrid(*coords, indexing="ij")]).T


def cell_pmf(prob_grid: torch.Tensor) -> torch.Tensor:
    """sample.py:27-31: pmf over grid cells = (ravel(rho) + 1e-15) / sum."""
    pdf = torch.ravel(prob_grid) + 1.0e-15
    return pdf / torch.sum(pdf)


def sample_grid(prob_grid: torch.Tensor, edges: Sequence[torch.Tensor], size: int,
                generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """sample.py:34-57 (noise=0): multinomial over cells, then uniform jitter per axis."""
    idx = cell_pmf(prob_grid).multinomial(num_samples=size, replacement=True, generator=generator)
    sub = torch.unravel_index(idx, prob_grid.shape)
    cols = []
    for axis, e in enumerate(edges):
        lb, ub = e[sub[axis]], e[sub[axis] + 1]
        cols.append(lb + (ub - lb) * torch.rand(size, generator=generator))
    return torch.stack(cols, dim=1)
