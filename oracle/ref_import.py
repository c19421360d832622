"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference package.

Imports ``/root/reference/mentflow`` in THIS container so that golden vectors can be
generated from the reference's own code (``oracle/make_goldens.py``).  The reference
depends on third-party packages that are not installed here (zuko, ot, skimage,
matplotlib, psdist, ultraplot); they are replaced by inert placeholder modules, which is
enough for everything on the hot path except the zuko flow itself.

``/root/reference`` does not exist on the GPU box: nothing under ``tests/ -m gpu``,
``bench.py`` or ``__graft_entry__.smoke()`` may call this module.  It is used only by
``oracle/make_goldens.py`` and by CPU tests that are skipped when the reference is absent.
"""
import importlib
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("MENTFLOW_REFERENCE_ROOT", "/root/reference")

_PLACEHOLDERS = [
    "zuko", "zuko.flows", "ot", "ot.lp", "skimage", "skimage.io", "matplotlib",
    "matplotlib.pyplot", "psdist", "psdist.plot", "ultraplot", "tqdm.notebook",
]


class _Inert(types.ModuleType):
    """A module whose every attribute is another inert module (and is callable)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Inert(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *args, **kwargs):
        return self


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mentflow"))


def load():
    """Return the reference ``mentflow`` package (raises if the tree is absent)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "mentflow" in sys.modules:
        return sys.modules["mentflow"]
    for name in _PLACEHOLDERS:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Inert(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("mentflow")
