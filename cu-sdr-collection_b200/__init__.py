"""B200-native GNSS correlator engine — drop-in for the ``acquisition()`` / ``tracking()`` hot
path of gnsscusdr/CU-SDR-Collection (GPS/GPS_L1CA/include/acquisition.m, tracking.m).

The compute lives in ``libgnsscorr.so`` (hand-written sm_100a CUDA behind the C ABI declared in
``include/gnsscorr.h``).  This package is the host-side mirror of the reference's MATLAB
interface for that path: same function names, argument meaning and result fields.  There is
no CPU fallback: every compute call raises if the library or a B200 is missing.
"""
from .settings import Settings, init_settings, num_to_process   # noqa: F401
from .engine import Engine, MultiEngine, GnssCorrError, lib_path       # noqa: F401
from .acquisition import acquisition                      # noqa: F401
from .tracking import tracking, TRACK_FIELDS              # noqa: F401
from .prerun import preRun                                # noqa: F401

__all__ = ["Settings", "init_settings", "Engine", "MultiEngine", "GnssCorrError", "acquisition", "tracking",
           "preRun", "TRACK_FIELDS", "lib_path"]
