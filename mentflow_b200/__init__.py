"""mentflow_b200 -- B200-native implementation of MENT-Flow's reconstruction hot path behind
the reference package's Python API (module names follow ``mentflow``)."""
from . import diagnostics, entropy, generate, graphs, loss, ment, prior, sample, simulate, utils
from .core import MENTFlow
from .utils import unravel

__all__ = ["MENTFlow", "diagnostics", "entropy", "generate", "graphs", "loss", "ment", "prior", "sample", "simulate", "utils",
           "unravel"]
