"""Device-side timing of the L1CA acquisition grid for the correlation-stage variants (dev tool).
usage: acq_bench.py [signal] ; env GC_ACQ_PATH / GC_ACQ_CLUSTER select the variant."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from cu_sdr_collection_b200 import Engine, init_settings, synth
fs = 16.368e6
sc = synth.default_scene(fs=fs, nsat=8)
s = init_settings(samplingFreq=fs, msToProcess=100)
rec = synth.make_record_torch(sc, 16368 * 60, device="cuda")
eng = Engine(s); eng.set_record(rec)
res = None
for it in range(6):
    acq = eng.acquire()
    st = eng.stats()
    if it >= 2:
        print("path %d total %.3f fwd %.3f corr %.3f (dominant %.3f, cols %.3f) fine %.3f launches %d" % (
            st["acq_path"], st["acq_total_ms"], st["acq_fwd_ms"], st["acq_corr_ms"], st["corr_rows_ms"], st["corr_cols_ms"],
            st["acq_fine_ms"], st["acq_launches"]))
print("acquired:", [(p + 1, acq["carrFreq"][p], int(acq["codePhase"][p]), round(float(acq["peakMetric"][p]), 4)) for p in range(32) if acq["carrFreq"][p]])
print("metric checksum %.9g" % float(np.sum(acq["peakMetric"])), "bins", acq["coarseBin"].tolist())
