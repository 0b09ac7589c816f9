"""Tracking speed per record format (dev tool): 12 channels x 3000 ms of the same scene as int8 I/Q (bulk-copied windows), int16 I/Q,
real int8 and 2-bit packed I/Q (per-sample accessor)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
from cu_sdr_collection_b200 import Engine, init_settings, synth, preRun
fs, nms = 16.368e6, 3000
sc = synth.default_scene(fs=fs, nsat=8)
iq8 = synth.make_record(sc, 16368 * (nms + 60))
recs = {(2, "schar"): iq8, (2, "int16"): (iq8.astype(np.int16) * 90), (1, "schar"): iq8[0::2].copy(),
        (3, "schar"): synth.pack_cplx2(synth.quantize2(iq8))}
for (ft, dt), rec in recs.items():
    s = init_settings(samplingFreq=fs, fileType=ft, dataType=dt, msToProcess=nms, numberOfChannels=12)
    eng = Engine(s); eng.set_record(rec)
    acq = eng.acquire()
    ch = preRun(acq, s)
    prn = [c["PRN"] for c in ch]; af = [c["acquiredFreq"] for c in ch]; cp = [float(c["codePhase"]) for c in ch]
    live = [i for i, p in enumerate(prn) if p]
    while len(live) and len([p for p in prn if p]) < 12:                    # fill the 12 channels with repeats of the acquired ones
        j = prn.index(0); k = live[len([p for p in prn if p]) % len(live)]
        prn[j], af[j], cp[j] = prn[k], af[k], cp[k]
    for _ in range(2):
        out, vv, vi, done = eng.track(prn, af, cp, nms)
    ms = eng.stats()["track_kernel_ms"]
    print("fileType %d %-6s record %7.1f MB  acquired %d  track %.2f ms  %.2f us/epoch  acq %.3f ms" % (
        ft, dt, rec.nbytes / 1e6, eng.stats()["n_acquired"], ms, 1e3 * ms / nms, eng.stats()["acq_total_ms"]), flush=True)
    eng.close()
