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
        cf0 = [float(c["codeFreq"]) for c in channel[:nch]] if (settings.signal in ("BDS_B3I", "BDS_B1C") or settings.is_fam5) else None   # B3I tracking.m:57, GPS_L5C :173, B1C NB_tracking.m:163
        out, vv, vi, done = eng.track(prn, af, cp, n, path=path, code_freq0=cf0)
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
        if out.shape[1] == 17:                                             # GPS_L5C tracking.m:57-60, 323-324
            tr["Pilot_I_P"], tr["Pilot_Q_P"] = out[ch, 15], out[ch, 16]
        if settings.signal in ("BDS_B2a", "BDS_B1C"):                      # BDS/B2a/include/tracking.m:66-72, 336-352; B1C NB_tracking.m:65-69, 340-355
            tr.update(_b2a_cno_pld(tr, settings, int(done[ch])))
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


def _b2a_cno_pld(tr: dict, settings: Settings, done: int) -> dict:
    """DataCNo / DataPLD (/ PilotCNo / PilotPLD / B2a_CNo) every settings.CNoInterval epochs from the recorded prompt
    rows - BDS/B2a/include/Calc_CNo_PLD.m:38-76 and the 0.5/0.5 smoothing of tracking.m:340-349.  Scalar host work
    on rows the GPU produced (40..200 values per call)."""
    n_int = int(settings.CNo_VSMinterval)
    nv = num_to_process(settings) // n_int
    pilot = int(settings.pilotTRKflag) == 1
    total = "B1C_CNo" if settings.signal == "BDS_B1C" else "B2a_CNo"
    res = {"DataCNo": np.zeros(nv), "DataPLD": np.zeros(nv)}
    if pilot:
        res.update({"PilotCNo": np.zeros(nv), "PilotPLD": np.zeros(nv), total: np.zeros(nv)})
    T = settings.intTime
    prev = np.zeros(3)

    def one(I, Q):
        Z = I ** 2 + Q ** 2
        Zm, Zv = np.mean(Z), np.var(Z, ddof=1)
        with np.errstate(invalid="ignore", divide="ignore"):
            Pav = np.sqrt(np.complex128(Zm ** 2 - Zv))
            Nv = 0.5 * (Zm - Pav)
            cno = np.abs((1 / T) * Pav / (2 * Nv))
            a = (np.sum(I[I > 0]) - np.sum(I[I < 0])) ** 2
            return cno, (a - np.sum(Q) ** 2) / (a + np.sum(Q) ** 2)

    for v in range(1, nv + 1):
        hi = v * n_int
        if hi > done:
            break
        cur = np.zeros(3)
        with np.errstate(invalid="ignore", divide="ignore"):
            d_cno, d_pld = one(tr["I_P"][hi - n_int:hi], tr["Q_P"][hi - n_int:hi])
            cur[0] = 10 * np.log10(d_cno)
            p_cno = 0.0
            if pilot:                                  # Calc_CNo_PLD.m:60-61: the pilot rows swap roles
                p_cno, p_pld = one(tr["Pilot_Q_P"][hi - n_int:hi], tr["Pilot_I_P"][hi - n_int:hi])
                cur[1] = 10 * np.log10(p_cno)
                res["PilotPLD"][v - 1] = p_pld
            cur[2] = 10 * np.log10(d_cno + p_cno)
        res["DataCNo"][v - 1] = cur[0] * 0.5 + prev[0] * 0.5
        res["DataPLD"][v - 1] = d_pld
        if pilot:
            res["PilotCNo"][v - 1] = cur[1] * 0.5 + prev[1] * 0.5
            res[total][v - 1] = cur[2] * 0.5 + prev[2] * 0.5
        prev = cur
    return res
