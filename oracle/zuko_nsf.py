"""TEST INFRASTRUCTURE -- torch restatement of zuko 1.3.1's neural spline flow (NSF) as
mentflow builds and calls it.

**Parity vs. the real zuko package is UNPINNED**: zuko==1.3.1 (``pyproject.toml:11`` of the
reference) is a third-party dependency that is neither vendored under ``/root/reference`` nor
installable here (no network).  This file restates its published algorithm

    zuko/flows/autoregressive.py   MAF, MaskedAutoregressiveTransform
    zuko/flows/spline.py           NSF  (univariate=MonotonicRQSTransform, 3*bins-1 params)
    zuko/nn.py                     MaskedMLP, MaskedLinear
    zuko/transforms.py             MonotonicRQSTransform, AutoregressiveTransform
    zuko/distributions.py          NormalizingFlow, DiagNormal

and is anchored on the reference's call sites: ``generate/build.py:36-46`` (constructs
``zuko.flows.NSF(features, hidden_features=[units]*layers, transforms, bins)`` then inverts it
with ``Flow(flow.transform.inv, flow.base)``) and ``generate/flows/zuko.py:15-53`` (``rsample``,
``rsample_and_log_prob``, ``log_prob``, ``transform.inv``, per-layer ``transforms``).  It is
defended by property tests (``tests/test_oracle_nsf.py``: monotone, invertible, ladj equals
autograd Jacobian, autoregressive triangularity, identity outside +-bound, mask counts).
``cross_check_against_zuko()`` diffs it against the real package wherever that is importable.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# masks (zuko/nn.py MaskedMLP.__init__, zuko/flows/autoregressive.py)
# --------------------------------------------------------------------------------------
def layer_order(features: int, layer: int, passes: int = None) -> torch.Tensor:
    """MAF without randperm: arange for even layers, flipped arange for odd layers; then
    MaskedAutoregressiveTransform groups the order into `passes` classes: order // ceil(features / passes)
    (passes = features: fully autoregressive, passes = 2: coupling)."""
    order = torch.arange(features)
    order = order if layer % 2 == 0 else torch.flipud(order)
    if passes is not None:
        passes = min(max(int(passes), 1), features)
        order = torch.div(order, math.ceil(features / passes), rounding_mode="floor")
    return order


def masked_mlp_masks(order: torch.Tensor, total: int, hidden: Sequence[int]) -> List[torch.Tensor]:
    """Binary masks of every MaskedLinear of one autoregressive layer.

    adjacency[i, j] = order[i] > order[j] (output feature i may see input feature j), each
    row repeated ``total`` (= 3*bins-1) times; hidden units cycle through the reachable
    dependency classes; output rows are mapped back through ``inverse``.
    """
    out_order = torch.repeat_interleave(order, total)
    adjacency = out_order[:, None] > order[None, :]
    uniq, inverse = torch.unique(adjacency, dim=0, return_inverse=True)
    precedence = uniq.int() @ uniq.int().t() == uniq.sum(dim=-1)
    masks = []
    indices = None
    sizes = list(hidden) + [adjacency.shape[0]]
    for i, width in enumerate(sizes):
        mask = uniq if i == 0 else precedence[:, indices]
        if i < len(hidden):
            reachable = mask.sum(dim=-1).nonzero().squeeze(dim=-1)
            indices = reachable[torch.arange(width) % len(reachable)]
            mask = mask[indices]
        else:
            mask = mask[inverse]
        masks.append(mask.clone())
    return masks


class MaskedLinear(nn.Linear):
    """nn.Linear (default init over the full, unmasked shape) with F.linear(x, mask*W, b)."""

    def __init__(self, mask: torch.Tensor):
        super().__init__(mask.shape[1], mask.shape[0])
        self.register_buffer("mask", mask)

    def forward(self, x):
        return F.linear(x, self.mask * self.weight, self.bias)


# --------------------------------------------------------------------------------------
# monotonic rational-quadratic spline (zuko/transforms.py MonotonicRQSTransform)
# --------------------------------------------------------------------------------------
class RQSpline:
    def __init__(self, widths, heights, derivatives, bound: float = 5.0, slope: float = 1e-3):
        log_slope = math.log(slope)
        widths = widths / (1 + abs(2 * widths / log_slope))
        heights = heights / (1 + abs(2 * heights / log_slope))
        derivatives = derivatives / (1 + abs(derivatives / log_slope))
        widths = F.pad(F.softmax(widths, dim=-1), (1, 0), value=0)
        heights = F.pad(F.softmax(heights, dim=-1), (1, 0), value=0)
        derivatives = F.pad(derivatives, (1, 1), value=0)
        self.horizontal = bound * (2 * torch.cumsum(widths, dim=-1) - 1)
        self.vertical = bound * (2 * torch.cumsum(heights, dim=-1) - 1)
        self.derivatives = torch.exp(derivatives)
        self.bins = self.derivatives.shape[-1] - 1

    def _bin(self, k):
        mask = torch.logical_and(0 <= k, k < self.bins)
        k = (k % self.bins)[..., None]
        def take(t, kk):
            shape = torch.broadcast_shapes(t.shape[:-1], kk.shape[:-1])
            return t.expand(*shape, t.shape[-1]).gather(-1, kk.expand(*shape, 1)).squeeze(-1)

        x0, x1 = take(self.horizontal, k), take(self.horizontal, k + 1)
        y0, y1 = take(self.vertical, k), take(self.vertical, k + 1)
        d0, d1 = take(self.derivatives, k), take(self.derivatives, k + 1)
        s = (y1 - y0) / (x1 - x0)
        return mask, x0, x1, y0, y1, d0, d1, s

    @staticmethod
    def _search(seq, value):
        return torch.sum(seq < value[..., None], dim=-1)

    def call_and_ladj(self, x):
        k = self._search(self.horizontal, x) - 1
        mask, x0, x1, y0, y1, d0, d1, s = self._bin(k)
        z = mask * (x - x0) / (x1 - x0)
        den = s + (d0 + d1 - 2 * s) * z * (1 - z)
        y = y0 + (y1 - y0) * (s * z ** 2 + d0 * z * (1 - z)) / den
        jac = s ** 2 * (2 * s * z * (1 - z) + d0 * (1 - z) ** 2 + d1 * z ** 2) / den ** 2
        return torch.where(mask, y, x), mask * jac.log()

    def inverse(self, y):
        k = self._search(self.vertical, y) - 1
        mask, x0, x1, y0, y1, d0, d1, s = self._bin(k)
        y_ = mask * (y - y0)
        a = (y1 - y0) * (s - d0) + y_ * (d0 + d1 - 2 * s)
        b = (y1 - y0) * d0 - y_ * (d0 + d1 - 2 * s)
        c = -s * y_
        z = 2 * c / (-b - (b ** 2 - 4 * a * c).sqrt())
        return torch.where(mask, x0 + z * (x1 - x0), y)


# --------------------------------------------------------------------------------------
# one autoregressive layer and the whole flow
# --------------------------------------------------------------------------------------
class AutoregressiveSplineLayer(nn.Module):
    """zuko MaskedAutoregressiveTransform(univariate=MonotonicRQSTransform, passes=features)."""

    def __init__(self, features: int, hidden: Sequence[int], bins: int, order: torch.Tensor):
        super().__init__()
        self.features, self.bins = features, bins
        self.total = 3 * bins - 1
        self.register_buffer("order", order.clone())
        layers: List[nn.Module] = []
        for mask in masked_mlp_masks(order, self.total, hidden):
            layers += [MaskedLinear(mask), nn.ReLU()]
        self.hyper = nn.Sequential(*layers[:-1])

    def spline(self, v: torch.Tensor) -> RQSpline:
        phi = self.hyper(v).unflatten(-1, (self.features, self.total))
        w, h, d = phi.split([self.bins, self.bins, self.bins - 1], dim=-1)
        return RQSpline(w, h, d)

    def call_and_ladj(self, v):
        y, ladj = self.spline(v).call_and_ladj(v)
        return y, ladj.sum(dim=-1)

    def forward(self, v):
        return self.call_and_ladj(v)[0]

    def inverse(self, y):
        v = torch.zeros_like(y)
        for _ in range(self.features):
            v = self.spline(v).inverse(y)
        return v


class NSFOracle(nn.Module):
    """The flow object mentflow's WrappedZukoFlow drives (generate/flows/zuko.py:10-53):
    sampling direction z -> x is a single conditioner pass per layer."""

    def __init__(self, features: int, hidden_units: int = 64, hidden_layers: int = 3,
                 transforms: int = 5, bins: int = 20, passes: int = None):
        super().__init__()
        self.features = features
        self.layers = nn.ModuleList([
            AutoregressiveSplineLayer(features, [hidden_units] * hidden_layers, bins,
                                      layer_order(features, i, passes))
            for i in range(transforms)])

    # base = DiagNormal(0, 1)
    def base_log_prob(self, z):
        return -0.5 * (z ** 2).sum(dim=-1) - 0.5 * self.features * math.log(2 * math.pi)

    def sample_base(self, n, generator=None):
        p = next(self.parameters())
        return torch.randn(n, self.features, generator=generator, dtype=p.dtype, device=p.device)

    # generate/flows/zuko.py:28-29, 34-41
    def forward_steps(self, z):
        out = [z]
        for layer in self.layers:
            out.append(layer(out[-1]))
        return out

    def forward(self, z):
        return self.forward_steps(z)[-1]

    # generate/flows/zuko.py:24-26 -> NormalizingFlow.rsample_and_log_prob
    def forward_and_log_prob(self, z) -> Tuple[torch.Tensor, torch.Tensor]:
        v, total = z, torch.zeros(z.shape[0], dtype=z.dtype, device=z.device)
        for layer in self.layers:
            v, ladj = layer.call_and_ladj(v)
            total = total + ladj
        return v, self.base_log_prob(z) - total

    def sample_and_log_prob(self, n, generator=None):
        return self.forward_and_log_prob(self.sample_base(n, generator))

    # generate/flows/zuko.py:31-32, 43-50
    def inverse_steps(self, x):
        out = [x]
        for layer in reversed(self.layers):
            out.append(layer.inverse(out[-1]))
        return out

    def inverse(self, x):
        return self.inverse_steps(x)[-1]

    # generate/flows/zuko.py:21-22 -> NormalizingFlow.log_prob
    def log_prob(self, x):
        v, total = x, torch.zeros(x.shape[0], dtype=x.dtype, device=x.device)
        for layer in reversed(self.layers):
            v = layer.inverse(v)
            _, ladj = layer.call_and_ladj(v)
            total = total + ladj
        return self.base_log_prob(v) - total


def cross_check_against_zuko(features=6, seed=0, n=4096) -> float:
    """if zuko is importable, copy its weights into the restatement and
    return the max abs difference of (x, log q) in float64.  Raises ImportError otherwise."""
    import zuko  # noqa: F401  (absent in this image)
    torch.manual_seed(seed)
    flow = zuko.flows.NSF(features, transforms=5, hidden_features=[64] * 3, bins=20).double()
    mine = NSFOracle(features).double()
    theirs = [p for p in flow.parameters()]
    with torch.no_grad():
        for dst, src in zip(mine.parameters(), theirs):
            dst.copy_(src)
    z = torch.randn(n, features, dtype=torch.float64)
    inv = zuko.flows.Flow(flow.transform.inv, flow.base)()
    x_ref, ladj = inv.transform.inv.call_and_ladj(z)
    lq_ref = inv.base.log_prob(z) - ladj
    x, lq = mine.forward_and_log_prob(z)
    return max((x - x_ref).abs().max().item(), (lq - lq_ref).abs().max().item())
