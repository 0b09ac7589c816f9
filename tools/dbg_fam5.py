import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import np_oracle as O
from cu_sdr_collection_b200 import Engine, init_settings, preRun, synth
from cu_sdr_collection_b200.codes import standin_codes
from helpers import to_oracle_settings
signal, pilot, nE = "GPS_L5C", int(sys.argv[1]) if len(sys.argv) > 1 else 1, 240
codes = standin_codes(signal)
sc = synth.default_scene_fam5(signal, codes, fs=18e6, nsat=2, seed=5)
for x in sc.sats: x.cn0 = 50
sv = sorted(x.prn for x in sc.sats)
s = init_settings(signal, acqSatelliteList=sv, acqNonCohTime=3, msToProcess=nE, numberOfChannels=2, pilotTRKflag=pilot, CNo_VSMinterval=40)
so = to_oracle_settings(s)
N = 18000
raw = synth.make_record(sc, N * (nE + 4))
acq = dict(carrFreq=np.zeros(63), codePhase=np.zeros(63), peakMetric=np.zeros(63))
for i, sat in enumerate(sc.sats):
    start = (10230 - sat.code_phase) * (18e6 / 10.23e6)
    acq["carrFreq"][sat.prn - 1] = round((s.IF + sat.doppler) / 25.0) * 25.0
    acq["codePhase"][sat.prn - 1] = int(round(start)) % N + 1
    acq["peakMetric"][sat.prn - 1] = 10.0 - i
ch = preRun(acq, s)
eng = Engine(s, codes=codes); eng.set_record(raw)
out, vv, vi, done = eng.track([c["PRN"] for c in ch], [c["acquiredFreq"] for c in ch], [float(c["codePhase"]) for c in ch], nE, code_freq0=[c["codeFreq"] for c in ch])
ref = O.tracking_fam5(raw, O.preRun_fam5(acq, so), so, codes)
names = O.TRACK_FIELDS
for i in range(2):
    scl = np.hypot(ref[i]["I_P"], ref[i]["Q_P"])
    for e in (0, 1, 2, 5, 10, 20, 50, 100, 200, 239):
        print(i, e, "I_P rel %.2e" % (abs(out[i, 3, e] - ref[i]["I_P"][e]) / scl[e]), "carrFreq d %.3e" % (out[i, 2, e] - ref[i]["carrFreq"][e]),
              "codeFreq d %.3e" % (out[i, 1, e] - ref[i]["codeFreq"][e]), "pll d %.3e" % (out[i, 11, e] - ref[i]["pllDiscr"][e]), "dll d %.3e" % (out[i, 9, e] - ref[i]["dllDiscr"][e]),
              "pll %.4f" % ref[i]["pllDiscr"][e])
print("---- first bad epochs")
for i in range(2):
    scl = np.hypot(ref[i]["I_P"], ref[i]["Q_P"])
    fields = [("I_P", 3), ("Q_P", 7), ("I_E", 4), ("Q_E", 6), ("I_L", 5), ("Q_L", 8)] + ([("Pilot_I_P", 15), ("Pilot_Q_P", 16)] if out.shape[1] == 17 else [])
    err = np.max(np.stack([np.abs(out[i, k] - ref[i][n]) / scl for n, k in fields]), axis=0)
    bad = np.nonzero(err > 3e-7)[0]
    print("ch", i, "bad epochs", bad[:10], "of", len(bad))
    for e in bad[:2]:
        for ee in (e - 1, e):
            print("  epoch", ee, {n: "%.2e" % (abs(out[i, k, ee] - ref[i][n][ee]) / scl[ee]) for n, k in fields})
            print("     abs", out[i, 0, ee], ref[i]["absoluteSample"][ee], "rem", repr(out[i, 13, ee]), repr(ref[i]["remCodePhase"][ee]), "codeFreq", repr(out[i, 1, ee]), repr(ref[i]["codeFreq"][ee]),
                  "blk", (out[i, 0, ee + 1] - out[i, 0, ee]) if ee + 1 < nE else None, "carr", repr(out[i, 2, ee]), repr(ref[i]["carrFreq"][ee]), "remCarr", repr(out[i, 14, ee]), repr(ref[i]["remCarrPhase"][ee]))
