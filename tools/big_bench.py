"""Device-side timing of the long-transform acquisitions (Galileo E1, GPS L2C, BDS B1C / B1I) at the reference's default
settings (dev tool).  usage: big_bench.py [signal ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from cu_sdr_collection_b200 import Engine, init_settings, synth
from cu_sdr_collection_b200.codes import standin_b1c_codes, standin_e1_codes, standin_varb_codes
from cu_sdr_collection_b200.settings import samples_per_code

want = sys.argv[1:] or ["E1C20", "E1C18", "L2C", "B1C", "B1I"]
for name in want:
    if name.startswith("E1C"):
        fs = 20e6 if name == "E1C20" else 18e6
        cd = standin_e1_codes()
        st = init_settings("GAL_E1C", samplingFreq=fs, **(dict(acqSatelliteList=list(range(1, 37)), acqSearchBand=8000.0, acqSearchStep=200.0) if fs == 20e6 else {}))
        sc, per = synth.default_scene_e1c(cd, fs=fs, nsat=4), 42
    elif name == "E5B":
        from cu_sdr_collection_b200.codes import standin_codes
        cd = standin_codes("GAL_E5b"); st = init_settings("GAL_E5b")
        sc, per = synth.default_scene_fam5("GAL_E5b", cd, fs=18e6, nsat=4), 102
    elif name == "L2C":
        cd = standin_varb_codes("GPS_L2C"); st = init_settings("GPS_L2C")
        sc, per = synth.default_scene_varb("GPS_L2C", cd, fs=8e6, nsat=3), 3
    elif name == "B1C":
        cd = standin_b1c_codes(); st = init_settings("BDS_B1C")
        sc, per = synth.default_scene_varb("BDS_B1C", cd, fs=18e6, nsat=3), 2
    else:
        cd = standin_varb_codes("BDS_B1I"); st = init_settings("BDS_B1I")
        sc, per = synth.default_scene_varb("BDS_B1I", cd, fs=18e6, nsat=4), 11
    n = samples_per_code(st)
    rec = torch.from_numpy(synth.make_record(sc, n * per + 64)).cuda()
    eng = Engine(st, codes=cd)
    eng.set_record(rec)
    for _ in range(3):
        eng.acquire()
    s = eng.stats()
    print("%-6s fft %6d path %d total %.3f ms (fwd %.3f corr %.3f) launches %d acquired %d" % (
        name, s["fft_len"], s["acq_path"], s["acq_total_ms"], s["acq_fwd_ms"], s["acq_corr_ms"], s["acq_launches"], s["n_acquired"]), flush=True)
    eng.close()
    del rec
