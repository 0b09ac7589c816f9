"""Small workload for ncu captures (dev tool): one full-grid acquisition + a short tracking run."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from cu_sdr_collection_b200 import Engine, init_settings, synth, preRun
fs = 16.368e6
nms = int(sys.argv[1]) if len(sys.argv) > 1 else 300
nsv = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sc = synth.default_scene(fs=fs, nsat=8)
s = init_settings(samplingFreq=fs, msToProcess=nms, acqSatelliteList=list(range(1, nsv + 1)))
rec = synth.make_record_torch(sc, 16368 * (nms + 50), device="cuda")
eng = Engine(s); eng.set_record(rec)
acq = eng.acquire()
ch = preRun(acq, s)
prn = [c["PRN"] for c in ch]; af = [c["acquiredFreq"] for c in ch]; cp = [float(c["codePhase"]) for c in ch]
if not any(prn):
    prn, af, cp = [7] * 12, [15025.0] * 12, [1441.0] * 12
out, vv, vi, done = eng.track(prn, af, cp, nms)
print(eng.stats(), done)
