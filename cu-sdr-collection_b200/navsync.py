"""Bit and frame synchronisation front end of ``postNavigation`` for GPS L1 C/A - host mirror of the first half of
GPS/GPS_L1CA/include/NAVdecoding.m (:69-170): ``subFrameStart`` and the navigation bits of five subframes per channel from
``trackResults(ch).I_P``, computed on the GPU through ``gc_nav_sync``.  Ephemeris decoding (:172-185) stays scalar host code."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .engine import Engine
from .settings import Settings


def nav_sync(trackResults: list, settings: Settings, engine: Engine):
    """Returns ``(subFrameStart, navBits)``: ``subFrameStart[ch]`` 1-based (0 = no valid preamble, NAVdecoding.m:143-146),
    ``navBits[ch]`` the 1501 bits (:152-166) as 0/1 or ``None`` where subFrameStart-20 .. subFrameStart+29999 leaves the record."""
    n = int(settings.msToProcess)
    ip = np.ascontiguousarray(np.stack([np.asarray(tr["I_P"], dtype=np.float64)[:n] for tr in trackResults]))
    nch = ip.shape[0]
    sfs = np.zeros(nch, dtype=np.int32)
    bits = np.zeros((nch, 1501), dtype=np.uint8)                    # GC_NAV_BITS: D30* of the previous subframe + 1500 bits
    valid = np.zeros(nch, dtype=np.int32)
    rc = engine.lib.gc_nav_sync(engine._h, nch, n, ip.ctypes.data_as(C.POINTER(C.c_double)), sfs.ctypes.data_as(C.POINTER(C.c_int32)),
                                bits.ctypes.data_as(C.POINTER(C.c_uint8)), valid.ctypes.data_as(C.POINTER(C.c_int32)))
    engine._check(rc, "gc_nav_sync")
    return sfs, [bits[ch] if valid[ch] else None for ch in range(nch)]
