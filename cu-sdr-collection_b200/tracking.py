"""``[trackResults, channel] = tracking(fid, channel, settings)`` — host mirror of
GPS/GPS_L1CA/include/tracking.m (signature :1, result struct :48-86, status handling :365).
The epoch loop runs on the GPU through ``gc_track`` / ``gc_track_file``."""
from __future__ import annotations

import numpy as np

from .engine import GC_SV_NONE, Engine
from .settings import Settings, num_to_process

# row order of the C ABI's output block == GC_F_* in include/gnsscorr.h
TRACK_FIELDS = ["absoluteSample", "codeFreq", "carrFreq", "I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L",
                "dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt", "remCodePhase", "remCarrPhase"]


def tracking(fid, channel: list, settings: Settings, engine: Engine | None = None):
    """``fid`` is an open binary file object (its ``name`` is used, as the MATLAB wrapper uses
    ``fopen(fid)``), or ``None`` when ``engine`` already holds the record.  Returns
    ``(trackResults, channel)``; ``trackResults[ch]`` has the reference's fields, ``status`` is
    copied from the channel only when every epoch was processed (tracking.m:365), and a record
    that runs out stops the whole call like the reference's early ``return`` (:241-245)."""
    own = engine is None
    eng = engine or Engine(settings)
    try:
        n = num_to_process(settings)                                        # tracking.m:90 (GAL_E1C tracking.m:48)
        nch = settings.numberOfChannels
        if settings.is_glonass:     # a GLONASS channel is live when status ~= '-' and is identified by K (GLO tracking.m:137-141)
            prn = [int(c["K"]) if c["status"] != "-" else GC_SV_NONE for c in channel[:nch]]
        else:
            prn = [int(c["PRN"]) for c in channel[:nch]]
        af = [float(c["acquiredFreq"]) for c in channel[:nch]]
        cp = [float(c["codePhase"]) for c in channel[:nch]]
        path = fid.name if fid is not None else None
        cf0 = [float(c["codeFreq"]) for c in channel[:nch]] if settings.signal == "BDS_B3I" else None   # B3I tracking.m:57
        out, vv, vi, done = eng.track(prn, af, cp, n, path=path, code_freq0=cf0)
    finally:
        if own:
            eng.close()
    results = []
    for ch in range(nch):
        tr = {"status": "-"}
        for i, f in enumerate(TRACK_FIELDS):
            tr[f] = out[ch, i]
        tr["CNo"] = {"VSMValue": vv[ch], "VSMIndex": vi[ch]}
        live = prn[ch] != (GC_SV_NONE if settings.is_glonass else 0)
        if live:
            tr["PRN"] = prn[ch]                                            # :138 (GLONASS stores K here)
        if live and done[ch] == n:
            tr["status"] = channel[ch]["status"]                           # :365
        tr["epochsDone"] = int(done[ch])
        results.append(tr)
    if any(prn[ch] != (GC_SV_NONE if settings.is_glonass else 0) and done[ch] < n for ch in range(nch)):
        print("Not able to read the specified number of samples  for tracking, exiting!")   # :242
    return results, channel
