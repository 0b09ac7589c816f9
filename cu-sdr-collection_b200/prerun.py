"""``channel = preRun(acqResults, settings)`` — GPS/GPS_L1CA/include/preRun.m:44-72.
Caller glue between the two hot functions (host side, scalar)."""
from __future__ import annotations

import numpy as np

from .settings import Settings


def preRun(acqResults: dict, settings: Settings) -> list:
    if settings.is_glonass:                                                     # GLO_GL1/include/preRun.m:44-72
        channel = [dict(K=0, acquiredFreq=0.0, codePhase=0, status="-") for _ in range(settings.numberOfChannels)]
        order = np.argsort(-np.asarray(acqResults["peakMetric"]), kind="stable")
        n = min(settings.numberOfChannels, int(np.sum(np.asarray(acqResults["carrFreq"]) != 0)))
        for ii in range(n):
            p = int(order[ii])
            channel[ii] = dict(K=p + 1 - 8, acquiredFreq=float(acqResults["carrFreq"][p]),     # Kindexes(ii)-8
                               codePhase=int(acqResults["codePhase"][p]), status="T")
        return channel
    channel = [dict(PRN=0, acquiredFreq=0.0, codePhase=0, status="-")
               for _ in range(settings.numberOfChannels)]                       # preRun.m:44-57
    # [junk, PRNindexes] = sort(peakMetric, 2, 'descend')  — stable, first index wins ties (:60)
    order = np.argsort(-np.asarray(acqResults["peakMetric"]), kind="stable")
    n = min(settings.numberOfChannels, int(np.sum(np.asarray(acqResults["carrFreq"]) != 0)))   # :65
    for ii in range(n):
        p = int(order[ii])
        channel[ii] = dict(PRN=p + 1, acquiredFreq=float(acqResults["carrFreq"][p]),
                           codePhase=int(acqResults["codePhase"][p]), status="T")              # :66-71
        if settings.signal == "GPS_L2C" and "CLCodePhase" in acqResults:     # GPS_L2C/include/preRun.m: channel.CLCodePhase
            channel[ii]["CLCodePhase"] = int(acqResults["CLCodePhase"][p])
        if settings.signal in ("BDS_B3I", "BDS_B1C") or settings.is_fam5:   # carrier-aided code NCO centre (BDS/B3I/include/preRun.m:71-73, GPS_L5C :69-71)
            channel[ii]["codeFreq"] = settings.codeFreqBasis + \
                (channel[ii]["acquiredFreq"] - settings.IF) / settings.carrFreqBasis * settings.codeFreqBasis
    if settings.signal in ("BDS_B3I", "BDS_B1C") or settings.is_fam5:
        for c in channel:
            c.setdefault("codeFreq", 0.0)
    return channel
