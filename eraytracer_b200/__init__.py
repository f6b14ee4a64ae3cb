"""eraytracer_b200 — B200-native hot path of plouj/eraytracer behind its render seam.

Only what the path needs lives here: csrc/ (CUDA kernels + the C ABI of
include/ert_b200.h), the ctypes binding, the scene flattening and the host-side
mirror of raytracer.erl's entry points.  The shared object is loaded on first
use; there is no CPU fallback anywhere in this package.
"""
from . import _lib            # noqa: F401
from . import scene           # noqa: F401
from . import ppm             # noqa: F401
from . import raytracer       # noqa: F401
from ._lib import BadArg, ErtError, Scene, device_count, load  # noqa: F401

__all__ = ["raytracer", "scene", "ppm", "Scene", "ErtError", "BadArg", "device_count", "load"]
