"""CUDA-graph capture of the forward hot path for a fixed batch size.

The reference's trainer evaluates ``model.loss(batch_size)`` a few hundred times per epoch with the
same shapes (train/train.py:163-214, and the no-grad convergence test at :235-261).  One pass is
~25 kernel launches; at the reference's batch sizes (2.5e4-1e5 particles) the launches, not the
kernels, bound the rate.  ``GraphedLoss`` captures one pass -- flow sample +
log-density, entropy, projections + KDE, discrepancies, loss -- into a CUDA graph and replays it.

Forward only (no autograd through a replay).  The capture is keyed on the generator's parameter
versions: after an optimiser step the next call re-captures.
"""
from __future__ import annotations

from typing import Optional

import torch


class GraphedLoss:
    def __init__(self, model, batch_size: int, warmup: int = 2) -> None:
        self.model = model
        self.batch_size = int(batch_size)
        self.warmup = max(1, int(warmup))
        gen = model.generator
        dev = next(gen.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedLoss needs a CUDA model; mentflow_b200 has no CPU fallback")
        self.device = dev
        self.z = torch.empty((self.batch_size, gen.features), dtype=torch.float32, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._key = None
        self._held = None
        self.out = None

    def _step(self):
        gen = self.model.generator
        x, logq = gen.forward_and_log_prob(self.z)
        return self.model.loss_from_particles(x, logq)

    def _weights_key(self):
        gen = self.model.generator
        return tuple((p.data_ptr(), p._version) for p in gen.parameters()) + (float(self.model.penalty_parameter),)

    def _capture(self) -> None:
        gen = self.model.generator
        with torch.cuda.device(self.device), torch.no_grad():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(self.warmup):
                    self._step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            # the packed weights / operand images are built during warm-up (outside the graph) and
            # must outlive it
            self._held = getattr(gen, "_pack_cache", None)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self._step()
        self._key = self._weights_key()

    def __call__(self, z: Optional[torch.Tensor] = None):
        """(L, H, [D_k]) of one pass.  ``z``: base noise (device or pinned host tensor of shape
        (batch_size, D)); None draws it on the device."""
        if self.graph is None or self._key != self._weights_key():
            self._capture()
        if z is None:
            self.z.normal_()
        else:
            self.z.copy_(z, non_blocking=True)
        self.graph.replay()
        return self.out
