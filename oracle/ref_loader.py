"""Import the live reference (read-only checkout) for golden-vector generation.

TEST INFRASTRUCTURE ONLY.  Works only where ``/root/reference`` exists (the
build container); the GPU box never has it, so nothing that runs there may
call :func:`load_reference` -- the committed fixtures under ``tests/golden/``
are what travels.

The reference imports plotting / graph packages at module import time
(scedar/eda/plot.py:4-20) that the hot path never calls and that are absent
from this image; permissive stub modules are pre-seeded for them.
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"

_STUBBED = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
    "matplotlib.gridspec", "matplotlib.patches", "matplotlib.cm",
    "mpl_toolkits", "mpl_toolkits.axes_grid1",
    "mpl_toolkits.axes_grid1.inset_locator", "seaborn", "fa2", "igraph",
    "umap", "nmslib", "xgboost", "leidenalg",
]


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Stub(self.__name__ + "." + name)
        sys.modules[child.__name__] = child
        return child

    def __call__(self, *args, **kwargs):
        return None


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scedar"))


def load_reference():
    """Return the reference's ``SampleDistanceMatrix`` class."""
    if not reference_available():
        raise RuntimeError("live reference not present at " + REFERENCE_ROOT)
    for name in _STUBBED:
        if name in sys.modules:
            continue
        m = _Stub(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__path__ = []
        sys.modules[name] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from scedar.eda.sdm import SampleDistanceMatrix
    return SampleDistanceMatrix
