"""MENT-Flow model: entropy-penalised reconstruction loss (mentflow/core.py:18-161)."""
from typing import Callable, Iterator, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
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

    _BATCHED = None

    def _batched_discrepancy(self, stacked, n_slots):
        """All D_k in a handful of kernels when the discrepancy is one of the library's (the
        reference evaluates K small expressions in a Python loop, core.py:89-93)."""
        from . import loss as _loss
        table = {_loss.kl_divergence: _loss.kl_divergence_batched,
                 _loss.mean_absolute_error: lambda p, t: (p - t).abs().reshape(p.shape[0], -1).mean(dim=1),
                 _loss.mean_square_error: lambda p, t: (p - t).square().reshape(p.shape[0], -1).mean(dim=1)}
        fn = table.get(self.discrepancy_function)
        if fn is None or not stacked or sum(len(e[0]) for e in stacked) != n_slots:
            return None
        held = [m for row in self.measurements for m in row]
        key = tuple((id(m), getattr(m, "_version", 0)) for m in held)
        if self._BATCHED is None or self._BATCHED[0] != key:
            offsets, base = [], 0
            for row in self.measurements:
                offsets.append(base)
                base += len(row)
            groups = []
            for slots, prof, *_ in stacked:
                meas = torch.stack([self.measurements[i][j] for i, j in slots]).to(prof.device)
                index = torch.tensor([offsets[i] + j for i, j in slots], device=prof.device)
                groups.append((meas, index))
            self._BATCHED = (key, groups, held)   # `held` keeps the ids in the key owned by these tensors
        out = None
        for (slots, prof, *fused), (meas, index) in zip(stacked, self._BATCHED[1]):
            # one-dimensional KDE screens arrive with their KL already evaluated by the kernel
            d = fused[0] if fused and fused[0] is not None else fn(prof, meas)
            if len(stacked) == 1 and index.numel() == n_slots:
                return d                      # single group in natural order (the usual case)
            if out is None:
                out = torch.zeros(n_slots, dtype=d.dtype, device=d.device)
            out = out.index_copy(0, index, d)
        return out

    def loss_from_particles(self, x: torch.Tensor, log_prob: torch.Tensor):
        """The part of ``loss`` after sampling (used by parity tests that fix the particles)."""
        est = self.entropy_estimator
        # sharded particles: the entropy's two moment sums are stashed on the reducer and travel at the tail
        # of the all-reduce of the profile sums -- one collective on the critical path instead of two
        pending = est.begin(x, log_prob) if (self.reducer is not None and hasattr(est, "begin")) else None
        if pending is None:
            H = est(x, log_prob)
        stacked = []
        from . import loss as _loss
        targets = self.measurements if self.discrepancy_function is _loss.kl_divergence else None
        predictions = simulate_forward(x, self.transforms, self.diagnostics, reducer=self.reducer, stacked=stacked,
                                       kl_targets=targets)
        if pending is not None:
            H = est.finish(pending)
        n_slots = sum(len(row) for row in predictions)
        dvec = self._batched_discrepancy(stacked, n_slots)
        if dvec is not None:
            D = list(dvec.unbind(0))
            if dvec.is_cuda and dvec.dtype == torch.float32 and (not torch.is_tensor(H) or H.dtype == torch.float32):
                # the scalar tail as one launch: it is on the critical path of every step
                h = H if torch.is_tensor(H) else (None if H == 0.0 else torch.full((), float(H), device=dvec.device))
                return ops.LossTail.apply(h, dvec, float(self.penalty_parameter)), H, D
            mean_d = dvec.mean()
        else:
            D = self.discrepancy_vector(predictions)
            mean_d = sum(D) / len(D)
        L = H + self.penalty_parameter * mean_d
        return L, H, D

    def loss(self, batch_size: int):
        """L = H + mu * mean_k D_k; returns (L, H, [D_k])  (core.py:95-117)."""
        x, log_prob = self.sample_and_log_prob(batch_size)
        return self.loss_from_particles(x, log_prob)

    def parameters(self, recurse: bool = True) -> Iterator[nn.Parameter]:
        return self.generator.parameters()

    # checkpoint layout of the reference (core.py:122-143): one dict with the generator's state_dict and the
    # pickled measurement setup under these keys
    _CHECKPOINT_OBJECTS = ("entropy_estimator", "transforms", "diagnostics", "measurements")

    def save(self, path) -> None:
        state = {name: getattr(self, name) for name in self._CHECKPOINT_OBJECTS}
        state["generator"] = self.generator.state_dict()
        torch.save(state, path)

    def load(self, path, device=None):
        state = torch.load(path, map_location=device, weights_only=False)
        try:
            self.generator.load_state_dict(state["generator"])
        except RuntimeError as err:
            raise RuntimeError("Error loading generative model. Architecture mismatch?") from err
        for name in self._CHECKPOINT_OBJECTS:
            setattr(self, name, state[name])
        self.to(device)

    def to(self, device):
        """Move transforms, screens, measurements and the generator (core.py:145-159)."""
        def moved(rows):
            return None if rows is None else [[item.to(device) for item in row] for row in rows]

        if self.transforms is not None:
            self.transforms = [t.to(device) for t in self.transforms]
        self.diagnostics = moved(self.diagnostics)
        self.measurements = moved(self.measurements)
        if self.generator is not None:
            self.generator = self.generator.to(device)
        return self
