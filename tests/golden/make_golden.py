"""Writes tests/golden/l1ca_small.npz from the NumPy oracle (oracle/np_oracle.py).

The reference is MATLAB-only and cannot be executed in this image, so these are NOT outputs of
the reference: they freeze the line-by-line restatement so that the C oracle, and through it the
CUDA path, are checked against a fixed set of numbers.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import np_oracle as O  # noqa: E402
from cu_sdr_collection_b200 import synth  # noqa: E402

fs, IF = 2.046e6, 20e3
sc = synth.default_scene(fs=fs, IF=IF, nsat=3, seed=11)
for s_ in sc.sats:
    s_.cn0 = 47.0
sv = sorted({x.prn for x in sc.sats} | {5})
s = O.Settings(samplingFreq=fs, IF=IF, acqNonCohTime=3, acqSearchBand=6000, msToProcess=80, numberOfChannels=4,
               acqSatelliteList=sv)
N = O.samples_per_code(s)
raw = synth.make_record(sc, N * 130)
a = O.acquisition(O.read_acq_signal(raw, s), s)
ch = O.preRun(a, s)
tr = O.tracking(raw, ch, s)
track = np.stack([np.stack([t[f] for f in O.TRACK_FIELDS]) for t in tr])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "l1ca_small.npz"),
                    raw=raw, fs=fs, IF=IF, nonCoh=s.acqNonCohTime, band=s.acqSearchBand, nEpochs=s.msToProcess,
                    nCh=s.numberOfChannels, svList=np.array(sv), carrFreq=a["carrFreq"], codePhase=a["codePhase"],
                    peakMetric=a["peakMetric"], coarseBin=a["coarseBin"],
                    chPRN=np.array([c["PRN"] for c in ch], dtype=np.int32),
                    chFreq=np.array([c["acquiredFreq"] for c in ch]), chCodePhase=np.array([float(c["codePhase"]) for c in ch]),
                    track=track, vsm=np.stack([t["VSMValue"] for t in tr]))
print("wrote golden:", {k: a[k][np.array(sv) - 1] for k in ("carrFreq", "codePhase", "peakMetric")})
