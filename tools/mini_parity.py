"""One small acquisition on the 45 x 32 x 25 plan (18 Msps) against the C oracle: the quickest parity check of the 25- and
45-point codelets (dev tool; the full check is pytest -m gpu)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from cu_sdr_collection_b200 import Engine, init_settings, synth
from helpers import c_acquisition, scene

fs = 18e6
sc = scene(fs, nsat=2, seed=42, cn0=47)
sv = sorted({s.prn for s in sc.sats} | {1})
s = init_settings(samplingFreq=fs, acqSatelliteList=sv, acqNonCohTime=2, msToProcess=60, numberOfChannels=2)
raw = synth.make_record(sc, 18000 * 50)
t0 = time.time()
eng = Engine(s, device=0)
acq = eng.acquire(sv, host_iq=raw)
t1 = time.time()
ref = c_acquisition(raw, s, sv)
idx = np.array(sv) - 1
ok = np.array_equal(acq["carrFreq"], ref["carrFreq"]) and np.array_equal(acq["codePhase"], ref["codePhase"])
err = np.max(np.abs(acq["peakMetric"][idx] / ref["peakMetric"][idx] - 1))
print("fft", eng.stats()["fft_len"], "indices exact", ok, "peakMetric max rel err %.3g" % err, "gpu s %.2f oracle s %.2f" % (t1 - t0, time.time() - t1))
