"""praline_b200 -- B200 (sm_100a) drop-in for PRALINE's pairwise DP alignment core.

Public host API (see engine.py); the CUDA library is loaded lazily and there is no CPU
fallback.  The PRALINE plug-in components live in praline_b200.plugin (imported on demand,
they need the `praline` package).
"""
from ._lib import PralineGpuError, LIB_PATH, EXPORTS  # noqa: F401
from . import matrices, synth  # noqa: F401


def get_engine(device=0):
    from .engine import get_engine as _g
    return _g(device)
