"""Where does the 12 x 60000-epoch run leave the C oracle?  The script behind profiles/r02_parity_60000.md (dev tool)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from cu_sdr_collection_b200 import Engine, init_settings, synth, preRun
from helpers import c_tracking, scene, TRACK_FIELDS
fs, nms = 16.368e6, int(sys.argv[1]) if len(sys.argv) > 1 else 60000
sc = scene(fs, nsat=12, seed=77)
for sat in sc.sats:
    sat.cn0 = max(sat.cn0, 42.0)
s = init_settings(samplingFreq=fs, msToProcess=nms, numberOfChannels=12)
N = 16368
rec = synth.make_record_torch(sc, N * (nms + 40), device="cuda")
eng = Engine(s); eng.set_record(rec)
acq = eng.acquire(); ch = preRun(acq, s)
prn = [c["PRN"] for c in ch]; af = [c["acquiredFreq"] for c in ch]; cp = [float(c["codePhase"]) for c in ch]
out, vv, vi, done = eng.track(prn, af, cp, nms)
print('kernel ms', eng.stats()['track_kernel_ms'], flush=True)
raw = rec.cpu().numpy()
ref, rvv, rvi, rdone = c_tracking(raw, s, prn, af, cp, nms, parallel=1)
np.set_printoptions(precision=17, linewidth=200)
for c in range(12):
    d = np.nonzero(out[c, 0] != ref[c, 0])[0]
    sc_ = np.hypot(ref[c, 3], ref[c, 7])
    e = np.abs(out[c, 3] - ref[c, 3]) / sc_
    big = np.nonzero(e > 1e-6)[0]
    print("ch", c, "prn", prn[c], "first absSample diff", d[:3], "n", d.size, "first I_P>1e-6", big[:3], "n", big.size, "max", e.max())
    k = None
    if big.size: k = big[0]
    elif d.size: k = d[0]
    if k is not None:
        for j in range(max(0, k - 1), min(nms, k + 1)):
            print("  epoch", j)
            for i, f in enumerate(TRACK_FIELDS):
                print("    %-14s %-26r %-26r %.3e" % (f, out[c, i, j], ref[c, i, j], out[c, i, j] - ref[c, i, j]))
