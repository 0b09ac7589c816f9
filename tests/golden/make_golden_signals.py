"""Writes tests/golden/<signal>_case.npz for the twelve signal folders from the NumPy oracle (tests/golden_cases.py).

NOT outputs of the reference (MATLAB-only, cannot run in this image): they freeze the restatement's results on seeded records so
that the oracle, the C ABI and later refactors of either are compared with one committed set of numbers per folder.
    python tests/golden/make_golden_signals.py [signal ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import golden_cases as G  # noqa: E402

for sig in (sys.argv[1:] or G.SIGNALS):
    t0 = time.time()
    case = G.build(sig)
    out = G.oracle_outputs(case)
    np.savez_compressed(G.fixture_path(sig), **out)
    acq = {k: out["acq_" + k][np.nonzero(out["acq_carrFreq"])[0]] for k in G.ACQ_KEYS}
    print(f"{sig}: {time.time() - t0:.1f} s, {os.path.getsize(G.fixture_path(sig))} B, acquired {acq['carrFreq'].size}, "
          f"tracked {sum(1 for k in out if k.endswith('_I_P'))} channel(s)", flush=True)
