"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): one acquisition on the fused
and the generic path, tracking with 1-CTA and 8-CTA-cluster channels, a record that runs out."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
from cu_sdr_collection_b200 import Engine, init_settings, synth, preRun

for fs, ncoh in ((16.368e6, 2), (2.046e6, 2)):
    sc = synth.default_scene(fs=fs, nsat=2, seed=5)
    for s_ in sc.sats:
        s_.cn0 = 48
    sv = sorted({x.prn for x in sc.sats} | {1})
    s = init_settings(samplingFreq=fs, acqSatelliteList=sv, acqNonCohTime=ncoh, msToProcess=40, numberOfChannels=3)
    N = int(round(fs / 1000))
    raw = synth.make_record(sc, N * 60)
    eng = Engine(s)
    acq = eng.acquire(sv, host_iq=raw)
    ch = preRun(acq, s)
    prn = [c["PRN"] for c in ch]; af = [c["acquiredFreq"] for c in ch]; cp = [float(c["codePhase"]) for c in ch]
    eng.set_record(raw)
    for g in ("1", "8", "4"):
        os.environ["GC_TRACK_CLUSTER"] = g
        out, vv, vi, done = eng.track(prn, af, cp, 40)
        print(fs, "cluster", g, done, eng.stats()["acq_path"])
    eng.set_record(raw[: 2 * N * 30])
    out, vv, vi, done = eng.track(prn, af, cp, 40)
    print("short", done)
    eng.close()
print("sanitize run ok")
