"""Host-side view of one acquisition call (dev tool): wall clock per call while the call goes direct -> captured -> replayed,
the library's own host laps (GC_HOST_TIMING=1) of a cold gc_create + gc_acquire_host + gc_destroy, and the same with the graph
switched off."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
import torch
from cu_sdr_collection_b200 import Engine, init_settings, synth

fs = 16.368e6
sc = synth.default_scene(fs=fs, nsat=10, seed=20260101)
for s_ in sc.sats:
    s_.cn0 = max(s_.cn0, 44.0)
s = init_settings(samplingFreq=fs)
rec = synth.make_record_torch(sc, 16368 * 60, device="cuda")
host = torch.empty(2 * 16368 * 42, dtype=torch.int8).pin_memory()
host.copy_(rec[: host.numel()])
hnp = host.numpy()
torch.cuda.synchronize()

def run(tag, n=8, sv=None):
    eng = Engine(s)
    eng.set_record(rec)
    res = None
    for i in range(n):
        t0 = time.perf_counter(); a = eng.acquire(sv); dt = (time.perf_counter() - t0) * 1e3
        st = eng.stats()
        print(f"{tag} resident call {i}: wall {dt:.3f} ms  kernels {st['acq_total_ms']:.3f} (fwd {st['acq_fwd_ms']:.3f} rows {st['corr_rows_ms']:.3f} cols {st['corr_cols_ms']:.3f} fine {st['acq_fine_ms']:.3f}) launches {st['acq_launches']}")
        if res is not None:
            assert all(np.array_equal(a[k], res[k]) for k in ("carrFreq", "codePhase", "peakMetric")), "results changed between calls"
        res = a
    for i in range(n):
        t0 = time.perf_counter(); b = eng.acquire(sv, host_iq=hnp); dt = (time.perf_counter() - t0) * 1e3
        print(f"{tag} host call {i}: wall {dt:.3f} ms")
        assert all(np.array_equal(b[k], res[k]) for k in ("carrFreq", "codePhase", "peakMetric"))
    eng.close()
    return res

r1 = run("graph")
r4 = run("graph 4 PRN", sv=[1, 9, 17, 25])
os.environ["GC_ACQ_GRAPH"] = "0"
r2 = run("direct")
del os.environ["GC_ACQ_GRAPH"]
os.environ["GC_ACQ_LEGACY"] = "1"
r3 = run("legacy", n=4)
del os.environ["GC_ACQ_LEGACY"]
for k in ("carrFreq", "codePhase", "peakMetric"):
    assert np.array_equal(r1[k], r2[k]) and np.array_equal(r1[k], r3[k]), k
print("graph == direct == legacy results: ok; acquired", int(np.count_nonzero(r1["carrFreq"])))
os.environ["GC_HOST_TIMING"] = "1"
for i in range(2):
    t0 = time.perf_counter()
    e = Engine(s)
    a = e.acquire(host_iq=hnp)
    e.close()
    print(f"cold one-shot {i}: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
