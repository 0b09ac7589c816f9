"""``acqResults = acquisition(longSignal, settings)`` — host mirror of
GPS/GPS_L1CA/include/acquisition.m (signature :1, result fields :130-134, console line
:154,209,286,292).  The search itself runs on the GPU through ``gc_acquire_host``."""
from __future__ import annotations

import sys

import numpy as np

from .engine import Engine, GnssCorrError
from .settings import Settings


def _to_file_samples(longSignal: np.ndarray, settings: Settings, swapped: bool = False) -> np.ndarray:
    """longSignal -> the file's own samples: int8 or int16 (settings.dataType), I,Q interleaved for fileType 2, one value
    per sample for fileType 1 (postProcessing.m:83-96).  ``swapped``: the GLONASS read path builds ``data2 + 1i*data1``
    (GLO_GL1/include/postProcessing.m:94), so real/imag are Q/I."""
    x = np.asarray(longSignal)
    if settings.fileType == 3:
        from .synth import pack_cplx2
        return pack_cplx2(x)
    dt = np.int16 if settings.dataType == "int16" else np.int8
    lim = 32768 if dt is np.int16 else 128
    if settings.fileType == 1:
        if np.iscomplexobj(x):
            raise GnssCorrError("fileType 1 expects a real-valued longSignal")
        if not (np.all(x == np.rint(x)) and np.max(np.abs(x)) <= lim):
            raise GnssCorrError("longSignal is not integer valued in the range of settings.dataType")
        return x.astype(dt)
    if not np.iscomplexobj(x):
        raise GnssCorrError("fileType 2 expects a complex longSignal (I + 1i*Q)")
    re, im = x.real, x.imag
    if not (np.all(re == np.rint(re)) and np.all(im == np.rint(im)) and
            np.max(np.abs(re)) <= lim and np.max(np.abs(im)) <= lim):
        raise GnssCorrError("longSignal is not integer valued in the range of settings.dataType; the accelerated path needs the raw samples")
    iq = np.empty(2 * x.size, dtype=dt)
    iq[0::2] = (im if swapped else re).astype(dt)
    iq[1::2] = (re if swapped else im).astype(dt)
    return iq


def acquisition(longSignal, settings: Settings, engine: Engine | None = None, verbose: bool = True) -> dict:
    """Same contract as the reference function: ``longSignal`` is the complex row vector
    postProcessing.m:88-96 builds from the first max(42, acqNonCohTime+2) code periods (an int8
    I,Q-interleaved array is accepted too); returns ``acqResults`` with 1x32 ``carrFreq``,
    ``codePhase`` and ``peakMetric`` (``carrFreq == 0`` means not acquired).  GLONASS
    (GLO_GL1/include/acquisition.m): ``acqSatelliteList`` holds frequency numbers K, the vectors are
    1x21 and indexed K+8 (K+7 here, 0-based)."""
    own = engine is None
    eng = engine or Engine(settings)
    try:
        x = np.asarray(longSignal)
        iq = x if x.dtype in (np.int8, np.int16, np.uint8) else _to_file_samples(x, settings, swapped=settings.is_glonass and settings.fileType == 2)
        r = eng.acquire(settings.acqSatelliteList, host_iq=iq)
    finally:
        if own:
            eng.close()
    if settings.signal in ("BDS_B2a", "BDS_B1C"):  # acqResults vectors are 1 x max(acqSatelliteList) (BDS/B2a/include/acquisition.m:128-132)
        n = max(settings.acqSatelliteList)
        r = {k: v[:n] for k, v in r.items()}
    if verbose:                                   # acquisition.m:154,209,286,292
        sys.stdout.write("(")
        off = 7 if settings.is_glonass else -1
        for sv in settings.acqSatelliteList:
            sys.stdout.write("%02d " % sv if r["carrFreq"][sv + off] != 0 else ". ")
        sys.stdout.write(")\n")
    return r
