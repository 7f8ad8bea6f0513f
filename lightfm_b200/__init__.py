"""lightfm_b200 -- LightFM's fit_partial / predict hot path as sm_100a CUDA kernels.

``from lightfm_b200 import LightFM`` is a drop-in for ``from lightfm import LightFM``.
The first access loads ``csrc/libfm_cuda.so`` (build it with
``python -m lightfm_b200._build``); there is no CPU fallback.
"""
import importlib

__version__ = "0.1.0"
__all__ = ["LightFM", "evaluation", "cross_validation", "data", "__version__"]


def __getattr__(name):
    # Lazy so that `python -m lightfm_b200._build` can run before the library exists.
    if name == "LightFM":
        return importlib.import_module(".lightfm", __name__).LightFM
    if name in ("evaluation", "cross_validation", "data", "lightfm", "_lightfm_fast"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
