"""Tracking workload for ncu captures and A/B timing (dev tool): nCh channels x nMs epochs on a synthetic L1CA record."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from cu_sdr_collection_b200 import Engine, init_settings, synth, preRun
fs = 16.368e6
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 12
nms = int(sys.argv[2]) if len(sys.argv) > 2 else 300
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
sc = synth.default_scene(fs=fs, nsat=8)
for sat in sc.sats:
    sat.cn0 = max(sat.cn0, 44.0)
s = init_settings(samplingFreq=fs, msToProcess=nms)
rec = synth.make_record_torch(sc, 16368 * (nms + 50), device="cuda")
eng = Engine(s); eng.set_record(rec)
acq = eng.acquire()
ch = [c for c in preRun(acq, s) if c["PRN"]]
big = [ch[i % len(ch)] for i in range(nch)]
prn = [c["PRN"] for c in big]; af = [c["acquiredFreq"] for c in big]; cp = [float(c["codePhase"]) for c in big]
for _ in range(reps):
    out, vv, vi, done = eng.track(prn, af, cp, nms)
    st = eng.stats()
    print("track %d ch x %d ms: kernel %.3f ms  %.3f us/epoch  %.4g channel-ms/s" % (nch, nms, st["track_kernel_ms"], st["track_kernel_ms"] * 1e3 / nms, nch * nms / (st["track_kernel_ms"] * 1e-3)), flush=True)
