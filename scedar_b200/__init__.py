"""scedar_b200 -- the cell x cell distance matrix / kNN hot path of scedar on
NVIDIA B200 (sm_100a).

``SampleDistanceMatrix`` is the drop-in for ``scedar.eda.SampleDistanceMatrix``
on this path; ``device`` holds the tensor-level operators; ``stripes`` shards
the path by row stripes over several GPUs.  The CUDA library behind
``include/scedar_b200.h`` does all arithmetic -- there is no CPU fallback.
"""
__version__ = "0.1.0"


def __getattr__(name):
    # torch (and CUDA) are imported only when the compute API is touched
    if name == "SampleDistanceMatrix":
        from .sdm import SampleDistanceMatrix
        return SampleDistanceMatrix
    if name == "HClustTree":
        from .hclust import HClustTree
        return HClustTree
    if name in ("device", "sdm", "stripes", "synth", "_capi", "knn", "patch",
                "hclust"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
