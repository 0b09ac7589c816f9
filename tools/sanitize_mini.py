"""Smallest acquisition that runs every kernel changed in round 2's last session (row pass with rows-per-warp loop and the Rader
31-point codelet, fine search through moments, per-pair fine selection) - for compute-sanitizer memcheck."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
from cu_sdr_collection_b200 import Engine, init_settings, synth

fs = 16.368e6
sc = synth.default_scene(fs=fs, nsat=2, seed=5)
for s_ in sc.sats:
    s_.cn0 = 48
sv = sorted({x.prn for x in sc.sats} | {1})
s = init_settings(samplingFreq=fs, acqSatelliteList=sv, acqNonCohTime=2, msToProcess=40)
raw = synth.make_record(sc, 16368 * 60)
eng = Engine(s)
for _ in range(3):                      # direct enqueue, captured graph, replay
    acq = eng.acquire(sv, host_iq=raw)
print("acquired", [p + 1 for p in range(32) if acq["carrFreq"][p]], eng.stats()["acq_path"])
eng.close()
print("sanitize mini ok")
