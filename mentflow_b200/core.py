"""MENT-Flow model: entropy-penalised reconstruction loss (mentflow/core.py:18-161)."""
from typing import Callable, Iterator, List, Optional, Tuple

import torch
import torch.nn as nn

from .loss import kl_divergence
from .simulate import forward as simulate_forward
from .utils import unravel


class MENTFlow(nn.Module):
    """Generative maximum-entropy tomography solver with the reference's constructor and
    methods.  ``loss(batch_size)`` runs: flow sample + log-density (one fused kernel per
    layer) -> entropy reduction -> fused projection + KDE for all screens -> discrepancies."""

    def __init__(self, transforms, diagnostics, measurements, generator, prior=None, entropy_estimator=None,
                 discrepancy_function: Callable = kl_divergence, penalty_parameter: float = 10.0) -> None:
        super().__init__()
        self.transforms = transforms
        self.diagnostics = self.set_diagnostics(diagnostics)
        self.measurements = self.set_measurements(measurements)
        self.generator = generator
        self.prior = prior
        self.entropy_estimator = entropy_estimator
        if isinstance(discrepancy_function, str):
            # the reference's default is the *string* "kld" (core.py:28), which cannot be called
            from . import loss as _loss
            discrepancy_function = {"kld": _loss.kl_divergence, "mae": _loss.mean_absolute_error,
                                    "mse": _loss.mean_square_error}[discrepancy_function]
        self.discrepancy_function = discrepancy_function
        self.penalty_parameter = penalty_parameter
        self.reducer = None  # set by mentflow_b200.distributed.shard_model

    def set_diagnostics(self, diagnostics):
        self.diagnostics = [[]] if diagnostics is None else diagnostics
        return self.diagnostics

    def set_measurements(self, measurements):
        self.measurements = [[]] if measurements is None else measurements
        return self.measurements

    def sample(self, size: int) -> torch.Tensor:
        return self.generator.sample(int(size))

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        return self.generator.log_prob(x)

    def sample_and_log_prob(self, size: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.generator.sample_and_log_prob(int(size))

    def sample_and_entropy(self, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
        x, log_prob = self.sample_and_log_prob(n)
        return x, self.entropy_estimator(x, log_prob)

    def discrepancy_vector(self, predictions) -> List[torch.Tensor]:
        return [self.discrepancy_function(pred, meas)
                for pred, meas in zip(unravel(predictions), unravel(self.measurements))]

    def loss_from_particles(self, x: torch.Tensor, log_prob: torch.Tensor):
        """The part of ``loss`` after sampling (used by parity tests that fix the particles)."""
        H = self.entropy_estimator(x, log_prob)
        predictions = simulate_forward(x, self.transforms, self.diagnostics, reducer=self.reducer)
        D = self.discrepancy_vector(predictions)
        L = H + self.penalty_parameter * (sum(D) / len(D))
        return L, H, D

    def loss(self, batch_size: int):
        """L = H + mu * mean_k D_k; returns (L, H, [D_k])  (core.py:95-117)."""
        x, log_prob = self.sample_and_log_prob(batch_size)
        return self.loss_from_particles(x, log_prob)

    def parameters(self, recurse: bool = True) -> Iterator[nn.Parameter]:
        return self.generator.parameters()

    def save(self, path) -> None:
        state = {
            "generator": self.generator.state_dict(),
            "entropy_estimator": self.entropy_estimator,
            "transforms": self.transforms,
            "diagnostics": self.diagnostics,
            "measurements": self.measurements,
        }
        torch.save(state, path)

    def load(self, path, device=None):
        state = torch.load(path, map_location=device, weights_only=False)
        try:
            self.generator.load_state_dict(state["generator"])
        except RuntimeError:
            raise RuntimeError("Error loading generative model. Architecture mismatch?")
        self.entropy_estimator = state["entropy_estimator"]
        self.transforms = state["transforms"]
        self.diagnostics = state["diagnostics"]
        self.measurements = state["measurements"]
        self.to(device)

    def to(self, device):
        if self.transforms is not None:
            self.transforms = [t.to(device) for t in self.transforms]
        if self.diagnostics is not None:
            self.diagnostics = [[d.to(device) for d in row] for row in self.diagnostics]
        if self.measurements is not None:
            self.measurements = [[m.to(device) for m in row] for row in self.measurements]
        if self.generator is not None:
            self.generator = self.generator.to(device)
        return self
