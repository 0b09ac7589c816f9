"""``[trackResults, channel] = tracking(fid, channel, settings)`` — host mirror of
GPS/GPS_L1CA/include/tracking.m (signature :1, result struct :48-86, status handling :365).
The epoch loop runs on the GPU through ``gc_track`` / ``gc_track_file``."""
from __future__ import annotations

import numpy as np

from .engine import GC_PARAM_B1C_WB_FACTOR, GC_SV_NONE, Engine
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
        cf0 = [float(c["codeFreq"]) for c in channel[:nch]] if (settings.signal in ("BDS_B3I", "BDS_B1C") or settings.is_fam5) else None   # B3I tracking.m:57, GPS_L5C :173, B1C NB_tracking.m:163
        l2c_pilot = settings.signal == "GPS_L2C" and int(settings.pilotTRKflag) == 1
        clp = [int(c.get("CLCodePhase", 0)) if c["PRN"] != 0 else 1 for c in channel[:nch]] if l2c_pilot else None
        if settings.signal == "BDS_B1C" and int(settings.pilotTRKflag) == 2:   # factor = CalcWeighingFactor(settings), WB_tracking.m:124
            eng.set_param(GC_PARAM_B1C_WB_FACTOR, settings.wbFactor if settings.wbFactor is not None else calc_weighing_factor(settings))
        out, vv, vi, done = eng.track(prn, af, cp, n, path=path, code_freq0=cf0, cl_code_phase=clp)
        # BDS B2a / B1C: DataCNo / DataPLD / PilotCNo / PilotPLD / B2a_CNo every CNoInterval epochs, computed on the device from the
        # rows the kernel wrote (Calc_CNo_PLD.m:38-100 + the 0.5/0.5 smoothing of BDS/B2a/include/tracking.m:409-431)
        pld = eng.cno_pld(nch, n) if settings.signal in ("BDS_B2a", "BDS_B1C") else None
    finally:
        if own:
            eng.close()
    results = []
    for ch in range(nch):
        tr = {"status": "-"}
        for i, f in enumerate(TRACK_FIELDS):
            tr[f] = out[ch, i]
        if settings.signal == "GPS_L2C":
            # the loop runs in half chips (code NCO at 2*codeFreqBasis); the recorded values are chips, and absoluteSample
            # is the fractional sample of the code start (GPS_L2C/include/tracking.m:223, 250, 376, 382-383)
            e = int(done[ch])
            step = tr["codeFreq"][:e] / settings.samplingFreq
            tr["absoluteSample"] = tr["absoluteSample"].copy()
            tr["absoluteSample"][:e] = tr["absoluteSample"][:e] + 1 - tr["remCodePhase"][:e] / step
            for f in ("remCodePhase", "codeFreq", "dllDiscr", "dllDiscrFilt"):
                tr[f] = tr[f] / 2
        if out.shape[1] >= 17:                                             # GPS_L5C tracking.m:57-60, 323-324
            tr["Pilot_I_P"], tr["Pilot_Q_P"] = out[ch, 15], out[ch, 16]
        if out.shape[1] == 21:                                             # GPS_L2C tracking.m:396-402; B1C WB_tracking.m:409-414
            tr["Pilot_I_E"], tr["Pilot_I_L"], tr["Pilot_Q_E"], tr["Pilot_Q_L"] = out[ch, 17], out[ch, 18], out[ch, 19], out[ch, 20]
        if pld is not None:                                                # BDS/B2a/include/tracking.m:66-72, 409-431; B1C NB_tracking.m:65-69, 400-418
            tr["DataCNo"], tr["DataPLD"] = pld[ch, 0], pld[ch, 1]
            if int(settings.pilotTRKflag) >= 1:
                tr["PilotCNo"], tr["PilotPLD"] = pld[ch, 2], pld[ch, 3]
                tr["B1C_CNo" if settings.signal == "BDS_B1C" else "B2a_CNo"] = pld[ch, 4]
        else:
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


def calc_weighing_factor(settings: Settings) -> float:
    """``factor = CalcWeighingFactor(settings)`` (BDS/B1C/include/CalcWeighingFactor.m:45-82): weight of the data channel in
    the full-band code discriminator, from the BOC(1,1) and QMBOC power spectra integrated over the front-end bandwidth
    ``settings.FEBW``.  A settings-derived scalar (host side; the MATLAB wrapper calls the reference's own function)."""
    from scipy.integrate import quad
    fc = settings.codeFreqBasis
    Tc, Br = 1 / fc, settings.FEBW

    def boc(f, m):
        a = np.pi / (2 * m)
        return Tc * (np.sin(a * f / fc) * np.sin(np.pi * f / fc) / np.cos(a * f / fc) * fc / f / np.pi) ** 2

    def integ(g):
        return 2 * quad(g, 0, Br / 2, limit=400, epsabs=0, epsrel=1e-12)[0]

    g11 = lambda f: boc(f, 1)
    gp = lambda f: 29 / 33 * boc(f, 1) + 4 / 33 * boc(f, 6)
    p11, p11_2 = integ(g11), integ(lambda f: g11(f) * f ** 2)
    pp, pp_2 = integ(gp), integ(lambda f: gp(f) * f ** 2)
    t1 = 11 * p11 * (p11_2 / p11)
    t2 = 33 * pp * (pp_2 / pp)
    return float(t1 / (t1 + t2))
