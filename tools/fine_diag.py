"""Fine-search stage timing on the headline grid (dev tool; GC_FINE_VARIANT selects the fine_sum instantiation)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from cu_sdr_collection_b200 import Engine, init_settings, synth
fs = 16.368e6
sc = synth.default_scene(fs=fs, nsat=10, seed=20260101)
for s_ in sc.sats:
    s_.cn0 = max(s_.cn0, 44.0)
s = init_settings(samplingFreq=fs)
rec = synth.make_record_torch(sc, 16368 * 60, device="cuda")
eng = Engine(s); eng.set_record(rec)
for sv in (None, [1, 9, 17, 25]):
    for _ in range(5):
        a = eng.acquire(sv)
    st = eng.stats()
    print(f"variant {os.environ.get('GC_FINE_VARIANT', '0')} sv {'all' if sv is None else sv}: total {st['acq_total_ms']:.3f} fine {st['acq_fine_ms']:.3f} acquired {int(np.count_nonzero(a['carrFreq']))} checksum {float(np.sum(a['carrFreq'])):.1f}")
eng.close()
