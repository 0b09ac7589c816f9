"""Quick device-side timing of the full L1CA grid and a 12-channel tracking run (dev tool)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
from cu_sdr_collection_b200 import Engine, init_settings, synth, preRun

fs = 16.368e6
nms = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
sc = synth.default_scene(fs=fs, nsat=8)
s = init_settings(samplingFreq=fs, msToProcess=nms)
import torch
t0 = time.time()
rec = synth.make_record_torch(sc, 16368 * (nms + 50), device="cuda")
torch.cuda.synchronize(); print("record gen s", time.time() - t0, rec.numel())
eng = Engine(s)
eng.set_record(rec)
for it in range(3):
    t0 = time.time(); acq = eng.acquire(); dt = time.time() - t0
    st = eng.stats()
    print("acq wall ms %.2f" % (dt * 1e3), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
print("acquired:", [(p + 1, acq["carrFreq"][p], acq["codePhase"][p], round(acq["peakMetric"][p], 2)) for p in range(32) if acq["carrFreq"][p]])
print("truth:", [(x.prn, round(20e3 + x.doppler), round(x.cn0, 1)) for x in sc.sats])
ch = preRun(acq, s)
prn = [c["PRN"] for c in ch]; af = [c["acquiredFreq"] for c in ch]; cp = [float(c["codePhase"]) for c in ch]
for it in range(2):
    t0 = time.time(); out, vv, vi, done = eng.track(prn, af, cp, nms); dt = time.time() - t0
    st = eng.stats()
    nlive = sum(1 for p in prn if p)
    print("track wall ms %.1f kernel ms %.2f  -> %.3g channel-ms/s (kernel), us/epoch %.2f" % (dt * 1e3, st["track_kernel_ms"], nlive * nms / (st["track_kernel_ms"] * 1e-3), st["track_kernel_ms"] * 1e3 / nms), done)
print("CNo:", [round(float(v[-1]), 1) for v in vv[:nlive]])
# many-channel batch: replicate channels
for mult in (8, 32):
    P = (prn[:nlive] * mult); A = (af[:nlive] * mult); Cp = (cp[:nlive] * mult)
    n2 = min(nms, 500)
    out, vv, vi, done = eng.track(P, A, Cp, n2)
    st = eng.stats()
    print("batch %d ch x %d ms: kernel ms %.2f -> %.3g channel-ms/s" % (len(P), n2, st["track_kernel_ms"], len(P) * n2 / (st["track_kernel_ms"] * 1e-3)))
