"""Contract of a trainable particle generator, as ``MENTFlow``, the trainers and the notebooks of the
reference use it (mentflow/generate/base.py:8-26 plus the ``Distribution`` protocol of
mentflow/types_.py:13-25).  All tensors are float32 on the generator's device; ``n`` particles, ``D``
phase-space dimensions."""
import torch

# method -> what a subclass has to return
_CONTRACT = {
    "sample": "sample(n) -> (n, D) particles",
    "log_prob": "log_prob(x: (n, D)) -> (n,) log-density of the generator at x",
    "sample_and_log_prob": "sample_and_log_prob(n) -> ((n, D) particles, (n,) their log-density)",
    "forward": "forward(z: (n, D)) -> (n, D): base noise to particles",
    "inverse": "inverse(x: (n, D)) -> (n, D): particles back to base noise",
    "forward_steps": "forward_steps(z) -> list of (n, D): z and the output of every layer",
    "inverse_steps": "inverse_steps(x) -> list of (n, D): x and the output of every inverted layer",
    "sample_base": "sample_base(n) -> (n, D) draws of the base distribution",
    "dim": "dim() -> D",
}


def _required(name: str, contract: str):
    def method(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} must implement {contract}")

    method.__name__ = method.__qualname__ = name
    method.__doc__ = contract
    return method


class GenerativeModel(torch.nn.Module):
    """Base class of the generators (``NSFGenerator``); every method of the contract raises until a
    subclass provides it."""


for _name, _text in _CONTRACT.items():
    setattr(GenerativeModel, _name, _required(_name, _text))
del _name, _text
