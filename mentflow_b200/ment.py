"""Classical MENT (mentflow/ment.py): rho(x) = prior(x) * prod_k h_k(M_k x) with one Lagrange
table per measured profile, updated by Gauss-Seidel sweeps.

Interface of the reference's ``MENT`` / ``LagrangeFunction``; the arithmetic runs in the CUDA
library: ``prob`` is one fused kernel over all measurements (the reference does a device->numpy
->scipy->device round trip per measurement), sampling is a scan + inverse-CDF kernel, the
predicted profile of the sampled particles is the fused projection+KDE kernel, and the table
update is one elementwise kernel.  1-D screens (Histogram1D) are supported.
"""
import math
from typing import Any, Callable, List, Optional, Tuple

import torch

from . import ops
from .diagnostics import Histogram1D
from .loss import kl_divergence
from .prior import Gaussian, Uniform
from .simulate import forward as simulate_forward
from .simulate.simulate import _linear_matrix
from .utils import coords_from_edges, get_grid_points, unravel


class LagrangeFunction:
    """One h_k: values on the bin centres, linear interpolation in between, zero outside
    (ment.py:20-52; scipy RegularGridInterpolator(method="linear", fill_value=0))."""

    def __init__(self, coords, values: torch.Tensor, **interpolation_kws) -> None:
        method = interpolation_kws.get("method", "linear")
        if method != "linear":
            raise NotImplementedError("only linear interpolation of the Lagrange functions is implemented")
        self.coords = coords
        self.values = values

    def set_values(self, values: torch.Tensor) -> None:
        self.values = values

    def __call__(self, u: torch.Tensor) -> torch.Tensor:
        u = u.reshape(-1, 1).to(torch.float32)
        one = torch.ones((1, 1), dtype=torch.float32, device=u.device)
        return ops.ment_prob(u, one, self.coords.reshape(1, -1).to(u.device), self.values.reshape(1, -1).to(u.device),
                             0.0, 0.0)


class MENT:
    def __init__(self, ndim: int, transforms: List[Callable], diagnostics: List[List[Callable]],
                 measurements: List[List[torch.Tensor]], discrepancy_function: Callable = kl_divergence,
                 prior: Any = None, interpolation: str = "linear", mode: str = "integrate",
                 integration_limits=None, integration_shape=None, sampler: Optional[Callable] = None,
                 n_samples: int = 1000000, device=None, verbose: bool = False) -> None:
        self.device = device
        self.verbose = verbose
        self.mode = mode
        self.ndim = ndim
        self.epoch = 0
        self.transforms = transforms
        self.diagnostics = [[]] if diagnostics is None else diagnostics
        self.measurements = [[]] if measurements is None else measurements
        self.discrepancy_function = discrepancy_function
        # the reference's default refers to an undefined `UniformPrior` (ment.py:143)
        self.prior = prior if prior is not None else Uniform(ndim=ndim, scale=100.0)
        self.integration_limits = integration_limits
        self.integration_shape = integration_shape
        self.sampler = sampler
        self.n_samples = int(n_samples)
        self.interpolation = interpolation
        self.reducer = None
        self._packed = None
        for row in self.diagnostics:
            for d in row:
                if not isinstance(d, Histogram1D):
                    raise NotImplementedError("MENT on the CUDA path supports 1-D screens (Histogram1D)")
        self.lagrange_functions = self.initialize_lagrange_functions()

    # ------------------------------------------------------------------ bookkeeping
    def send(self, x: torch.Tensor) -> torch.Tensor:
        return x.type(torch.float32).to(self.device)

    def set_diagnostics(self, diagnostics):
        self.diagnostics = [[]] if diagnostics is None else diagnostics
        self._packed = None
        return self.diagnostics

    def set_measurements(self, measurements):
        self.measurements = [[]] if measurements is None else measurements
        return self.measurements

    def initialize_lagrange_functions(self):
        """h_k = 1 where the measurement is positive, else 0 (ment.py:169-182)."""
        self.lagrange_functions = []
        for index in range(len(self.measurements)):
            row = []
            for measurement, diagnostic in zip(self.measurements[index], self.diagnostics[index]):
                coords = coords_from_edges(diagnostic.edges)
                values = (measurement > 0.0).float()
                row.append(LagrangeFunction(coords, values, method=self.interpolation))
            self.lagrange_functions.append(row)
        self._packed = None
        return self.lagrange_functions

    def _slots(self):
        return [(i, j) for i in range(len(self.diagnostics)) for j in range(len(self.diagnostics[i]))]

    def _pack(self, device):
        """Static part of the kernel arguments: projection rows and bin centres of every table."""
        if self._packed is not None and self._packed["device"] == device:
            return self._packed
        proj, coords = [], []
        nb = None
        for i, j in self._slots():
            d = self.diagnostics[i][j]
            matrix = _linear_matrix(self.transforms[i])
            if matrix is NotImplemented:
                raise NotImplementedError("MENT on the CUDA path needs linear transforms")
            proj.append(d.projection_vector(matrix, self.ndim, device))
            c = coords_from_edges(d.edges.to(torch.float32)).to(device)
            nb = c.shape[0] if nb is None else nb
            if c.shape[0] != nb:
                raise NotImplementedError("all screens of a MENT model must have the same number of bins")
            coords.append(c)
        self._packed = {"device": device, "proj": torch.stack(proj).contiguous(),
                        "coords": torch.stack(coords).contiguous()}
        return self._packed

    def _tables(self, device) -> torch.Tensor:
        return torch.stack([self.lagrange_functions[i][j].values.to(device=device, dtype=torch.float32).reshape(-1)
                            for i, j in self._slots()]).contiguous()

    def _prior_args(self) -> Tuple[float, float]:
        if isinstance(self.prior, Gaussian):
            return -0.5 / self.prior.scale ** 2, self.prior.log_norm
        if isinstance(self.prior, Uniform):
            return 0.0, -math.log(self.prior.volume)
        raise NotImplementedError("MENT on the CUDA path supports the Gaussian and Uniform priors")

    # ------------------------------------------------------------------ density
    def prob(self, x: torch.Tensor) -> torch.Tensor:
        """rho(x) (ment.py:239-249)."""
        pk = self._pack(x.device)
        a, b = self._prior_args()
        return ops.ment_prob(x, pk["proj"], pk["coords"], self._tables(x.device), a, b)

    def prob_on_grid(self, sampler) -> torch.Tensor:
        """rho on the cell centres of a GridSampler grid, straight from the grid index."""
        device = self.device if self.device is not None else "cuda"
        pk = self._pack(torch.device(device))
        a, b = self._prior_args()
        return ops.ment_prob_grid(list(sampler.shape), sampler.first_centres(), sampler.cell_sizes(), pk["proj"],
                                  pk["coords"], self._tables(pk["proj"].device), a, b)

    def log_prob(self, x: torch.Tensor, pad: float = 1.0e-12) -> torch.Tensor:
        return torch.log(self.prob(x) + pad)

    def evaluate_lagrange_function(self, u: torch.Tensor, index: int, diag_index: int) -> torch.Tensor:
        diagnostic = self.diagnostics[index][diag_index]
        return self.lagrange_functions[index][diag_index](diagnostic.project(u))

    def sample(self, size: int) -> torch.Tensor:
        return self.send(self.sampler(self.prob, int(size)))

    def sample_and_log_prob(self, size: int):
        x = self.sample(size)
        return x, self.log_prob(x)

    def discrepancy_vector(self, predictions) -> List[torch.Tensor]:
        return [self.discrepancy_function(pred, meas)
                for pred, meas in zip(unravel(predictions), unravel(self.measurements))]

    # ------------------------------------------------------------------ simulation of one profile
    def normalize_projection(self, projection: torch.Tensor, index: int, diag_index: int) -> torch.Tensor:
        diagnostic = self.diagnostics[index][diag_index]
        bin_volume = diagnostic.edges[1] - diagnostic.edges[0]
        return projection / projection.sum() / bin_volume

    def get_meas_points(self, index: int, diag_index: int) -> torch.Tensor:
        return coords_from_edges(self.diagnostics[index][diag_index].edges)

    def get_integration_points(self, index: int, diag_index: int) -> torch.Tensor:
        limits = self.integration_limits[index][diag_index]
        shape = self.integration_shape[index][diag_index]
        coords = [self.send(torch.linspace(limits[k][0], limits[k][1], shape[k])) for k in range(len(shape))]
        return coords[0] if len(coords) == 1 else self.send(get_grid_points(*coords))

    def _simulate_integrate(self, index: int, diag_index: int) -> torch.Tensor:
        """pred[b] = sum over the integration grid of rho(M^-1 [pixel; grid]) (ment.py:267-317)."""
        limits = self.integration_limits[index][diag_index]
        shape = [int(s) for s in self.integration_shape[index][diag_index]]
        diagnostic = self.diagnostics[index][diag_index]
        transform = self.transforms[index]
        device = torch.device(self.device if self.device is not None else "cuda")
        pk = self._pack(device)
        a, b = self._prior_args()
        first = [float(limits[k][0]) for k in range(len(shape))]
        step = [(float(limits[k][1]) - float(limits[k][0])) / max(shape[k] - 1, 1) for k in range(len(shape))]
        meas_coords = coords_from_edges(diagnostic.edges.to(torch.float32)).to(device)
        pred = ops.ment_integrate(self.ndim, meas_coords, diagnostic.axis, shape, first, step,
                                  transform.matrix_inv.to(device=device, dtype=torch.float32).contiguous(), pk["proj"],
                                  pk["coords"], self._tables(device), a, b)
        return self.normalize_projection(pred, index, diag_index)

    def _simulate_sample(self, index: int, diag_index: int) -> torch.Tensor:
        """sample -> transform -> KDE profile -> normalise (ment.py:319-326)."""
        x = self.sample(int(self.n_samples))
        pred = simulate_forward(x, [self.transforms[index]], [[self.diagnostics[index][diag_index]]],
                                reducer=self.reducer)[0][0]
        return self.normalize_projection(pred, index, diag_index)

    def simulate(self, index: int, diag_index: int, **kws) -> torch.Tensor:
        if self.mode == "integrate":
            return self._simulate_integrate(index, diag_index, **kws)
        if self.mode == "sample":
            return self._simulate_sample(index, diag_index, **kws)
        raise ValueError(f"Invalid mode {self.mode}")

    def simulate_all(self, **kws):
        if self.mode == "integrate":
            return [[self._simulate_integrate(i, j, **kws) for j in range(len(self.diagnostics[i]))]
                    for i in range(len(self.diagnostics))]
        x = self.sample(self.n_samples)
        return simulate_forward(x, self.transforms, self.diagnostics, reducer=self.reducer)

    # ------------------------------------------------------------------ Gauss-Seidel relaxation
    def gauss_seidel_update(self, lr: float = 1.0, thresh: float = 1.0e-10, **kws) -> None:
        """h <- h * (1 + lr * (g / g* - 1)) where g != 0 and g* >= thresh, one measurement after the
        other, each update feeding the next simulation (ment.py:336-371)."""
        for index in range(len(self.transforms)):
            if self.verbose:
                print(f"index={index}")
            for diag_index in range(len(self.diagnostics[index])):
                lf = self.lagrange_functions[index][diag_index]
                measurement = self.measurements[index][diag_index]
                prediction = self.simulate(index, diag_index, **kws)
                values = lf.values.to(device=prediction.device, dtype=torch.float32).contiguous().clone()
                ops.gs_update(values.view(-1), measurement.to(prediction.device).reshape(-1), prediction.reshape(-1),
                              lr, thresh)
                lf.set_values(values)
        self.epoch += 1

    gauss_seidel_step = gauss_seidel_update   # name used by BASELINE.json's north_star

    # ------------------------------------------------------------------ persistence
    def save(self, path: str) -> None:
        state = {"lagrange_functions": self.lagrange_functions, "epoch": self.epoch, "transforms": self.transforms,
                 "diagnostics": self.diagnostics, "measurements": self.measurements, "prior": self.prior,
                 "ndim": self.ndim, "sampler": self.sampler}
        torch.save(state, path)

    def load(self, path: str, device=None) -> None:
        state = torch.load(path, map_location=device, weights_only=False)
        for key in ("lagrange_functions", "epoch", "transforms", "diagnostics", "measurements", "prior", "ndim",
                    "sampler"):
            setattr(self, key, state[key])
        self._packed = None
        self.to(device)

    def to(self, device):
        self.device = device
        self._packed = None
        if self.transforms is not None:
            self.transforms = [t.to(device) for t in self.transforms]
        if self.diagnostics is not None:
            self.diagnostics = [[d.to(device) for d in row] for row in self.diagnostics]
        if self.measurements is not None:
            self.measurements = [[m.to(device) for m in row] for row in self.measurements]
        if self.sampler is not None:
            self.sampler = self.sampler.to(device)
        if self.prior is not None:
            self.prior = self.prior.to(device)
        for row in self.lagrange_functions:
            for lf in row:
                lf.values = self.send(lf.values)
                lf.coords = self.send(lf.coords)
        return self
