"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): one acquisition on the fused
and the generic path, tracking with 1-CTA and 8-CTA-cluster channels, a record that runs out."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
from cu_sdr_collection_b200 import Engine, init_settings, synth, preRun

for fs, ncoh in ((16.368e6, 2), (2.046e6, 2)):
    sc = synth.default_scene(fs=fs, nsat=2, seed=5)
    for s_ in sc.sats:
        s_.cn0 = 48
    sv = sorted({x.prn for x in sc.sats} | {1})
    s = init_settings(samplingFreq=fs, acqSatelliteList=sv, acqNonCohTime=ncoh, msToProcess=40, numberOfChannels=3)
    N = int(round(fs / 1000))
    raw = synth.make_record(sc, N * 60)
    eng = Engine(s)
    acq = eng.acquire(sv, host_iq=raw)
    ch = preRun(acq, s)
    prn = [c["PRN"] for c in ch]; af = [c["acquiredFreq"] for c in ch]; cp = [float(c["codePhase"]) for c in ch]
    eng.set_record(raw)
    for g in ("1", "8", "4"):
        os.environ["GC_TRACK_CLUSTER"] = g
        out, vv, vi, done = eng.track(prn, af, cp, 40)
        print(fs, "cluster", g, done, eng.stats()["acq_path"])
    eng.set_record(raw[: 2 * N * 30])
    out, vv, vi, done = eng.track(prn, af, cp, 40)
    print("short", done)
    eng.close()
print("sanitize run ok")

# ---- later additions of the round: variants B / C on the fused plans, the long column kernels, the pilot tracking modes,
#      int16 / real records and the nav-bit front end (small grids: sanitizers run 10-50x slower) ------------------------------
from cu_sdr_collection_b200.codes import (standin_b1c_codes, standin_e1_codes, standin_l2c_cl_codes, standin_varb_codes,
                                          boc61_from_boc11)
from cu_sdr_collection_b200 import acquisition, tracking
from cu_sdr_collection_b200.navsync import nav_sync
os.environ.pop("GC_TRACK_CLUSTER", None)

# BDS B1I, fused 72000 plan (variant B: shifted spectra, corrVec of the winning row)
cd = standin_varb_codes("BDS_B1I")
sc = synth.default_scene_varb("BDS_B1I", cd, fs=18e6, nsat=1, seed=3)
s = init_settings("BDS_B1I", acqSatelliteList=[sc.sats[0].prn, 30], acqSearchBand=1.0)
raw = synth.make_record(sc, 18000 * 11)
eng = Engine(s, codes=cd)
eng.set_record(raw)
print("B1I", eng.acquire()["carrFreq"][sc.sats[0].prn - 1], eng.stats()["acq_path"])
eng.close()

# GPS L2C at 8 Msps, fused 320000 plan + CL phase search + CL pilot tracking
sc = synth.default_scene_varb("GPS_L2C", standin_varb_codes("GPS_L2C"), fs=8e6, nsat=1, seed=3)
sat = sc.sats[0]
sat.doppler = 300.0
cd = standin_l2c_cl_codes([sat.prn])
sc.codes = cd
s = init_settings("GPS_L2C", acqSatelliteList=[sat.prn], acqSearchBand=0.5, pilotTRKflag=1, msToProcess=60, numberOfChannels=1)
raw = synth.make_record(sc, 160000 * 5)
eng = Engine(s, codes=cd)
eng.set_record(raw)
acq = eng.acquire()
print("L2C", acq["carrFreq"][sat.prn - 1], acq["CLCodePhase"][sat.prn - 1], eng.stats()["acq_path"])
ch = [dict(PRN=sat.prn, acquiredFreq=float(round(s.IF + sat.doppler)), codePhase=int(round((20460 - sat.code_phase) * 8e6 / 1.023e6)) % 160000,
           status="T", CLCodePhase=(sat.bit_offset + 1) % 75 + 1)]
tr, _ = tracking(None, ch, s, engine=eng)
print("L2C pilot track", tr[0]["epochsDone"])
eng.close()

# BDS B1C at 18 Msps, fused 360000 plan (variant C, weighted data + pilot) + full-band tracking (three int8 tables)
base = standin_b1c_codes()
sc = synth.default_scene_varb("BDS_B1C", base, fs=18e6, nsat=1, seed=3)
sat = sc.sats[0]
cd = {sat.prn: (base[sat.prn][0], base[sat.prn][1], boc61_from_boc11(base[sat.prn][1]))}
sc.codes = cd
s = init_settings("BDS_B1C", acqSatelliteList=[sat.prn], acqSearchBand=100.0, pilotTRKflag=2, msToProcess=30, numberOfChannels=1)
sat.doppler = 60.0
raw = synth.make_record(sc, 180000 * 5)
eng = Engine(s, codes=cd)
eng.set_record(raw)
acq = eng.acquire()
print("B1C", acq["carrFreq"][sat.prn - 1], eng.stats()["acq_path"])
cf = round((s.IF + sat.doppler) / 25.0) * 25.0
ch = [dict(PRN=sat.prn, acquiredFreq=cf, codePhase=int(round((20460 - sat.code_phase) * 18e6 / 2.046e6)) % 180000 + 1, status="T",
           codeFreq=s.codeFreqBasis + (cf - s.IF) / s.carrFreqBasis * s.codeFreqBasis)]
tr, _ = tracking(None, ch, s, engine=eng)
print("B1C WB track", tr[0]["epochsDone"])
eng.close()

# Galileo E1 at 20 Msps, fused 160000 plan (two-level columns)
cd = standin_e1_codes()
sc = synth.default_scene_e1c(cd, fs=20e6, nsat=1, seed=3)
s = init_settings("GAL_E1C", samplingFreq=20e6, acqSatelliteList=[sc.sats[0].prn, 30], acqSearchBand=400.0, acqSearchStep=200.0)
raw = synth.make_record(sc, 80000 * 42 + 64)
eng = Engine(s, codes=cd)
eng.set_record(raw)
print("E1C", eng.acquire()["carrFreq"][sc.sats[0].prn - 1], eng.stats()["acq_path"])
eng.close()

# int16 real record (the per-sample accessor in acquisition and tracking) and the nav-bit front end
fs = 16.368e6
sc = synth.default_scene(fs=fs, nsat=2, seed=5)
sv = sorted({x.prn for x in sc.sats})
s = init_settings(samplingFreq=fs, fileType=1, dataType="int16", acqSatelliteList=sv, acqNonCohTime=2, msToProcess=30, numberOfChannels=2)
raw16 = (synth.make_record(sc, 16368 * 80)[0::2].astype(np.int32) * 90).astype(np.int16)
eng = Engine(s)
eng.set_record(raw16)
acq = eng.acquire()
ch = preRun(acq, s)
tr, _ = tracking(None, ch, s, engine=eng)
print("int16 real", [t["epochsDone"] for t in tr])
eng.close()
s = init_settings(samplingFreq=fs, msToProcess=14000, numberOfChannels=2)
eng = Engine(s)
rng = np.random.default_rng(1)
rows = [dict(I_P=1000.0 * (1 - 2 * rng.integers(0, 2, size=700)).repeat(20).astype(np.float64)) for _ in range(2)]
print("nav", nav_sync(rows, s, eng)[0])
eng.close()
print("sanitize run (extended) ok")

# ---- round 2: the graph enqueue (direct -> captured -> replayed), the asynchronous device-result call, the fractional-step
#      shifted spectra (GAL E5b), the quadrature-pilot tracker with the C/N0 + lock-detector kernel (BDS B2a) and the on-device
#      code generators (every code below is generated by the library) ------------------------------------------------------------
import torch
fs = 16.368e6
sc = synth.default_scene(fs=fs, nsat=2, seed=5)
for s_ in sc.sats:
    s_.cn0 = 48
sv = sorted({x.prn for x in sc.sats} | {1})
s = init_settings(samplingFreq=fs, acqSatelliteList=sv, acqNonCohTime=2)
raw = synth.make_record(sc, 16368 * 46)
eng = Engine(s)
eng.set_record(raw)
first = eng.acquire(sv)
for _ in range(3):
    again = eng.acquire(sv)
    assert np.array_equal(again["peakMetric"], first["peakMetric"])
buf = torch.zeros(128, dtype=torch.float64, device="cuda")
for _ in range(3):
    eng.acquire_device_async(sv, buf)
    torch.cuda.synchronize()
assert np.array_equal(buf.cpu().numpy()[:32], first["peakMetric"])
print("graph + async", eng.stats()["acq_launches"])
eng.close()

from cu_sdr_collection_b200.codes import icd_codes
cd = icd_codes("GAL_E5b")
sc = synth.default_scene_fam5("GAL_E5b", cd, fs=18e6, nsat=1, seed=5)
sc.sats[0].cn0 = 50
sv = [sc.sats[0].prn, 25]
s = init_settings("GAL_E5b", acqSatelliteList=sv, acqNonCohTime=2, acqSearchBand=900.0, acqSearchStep=60.0)   # 31 bins, 25 spectra per block
raw = synth.make_record(sc, 18000 * 104)
eng = Engine(s)
eng.set_record(raw)
print("E5b", eng.acquire(sv)["carrFreq"][sv[0] - 1], eng.stats()["acq_path"])
eng.close()

cd = icd_codes("BDS_B2a")
sc = synth.default_scene_fam5("BDS_B2a", cd, fs=18e6, nsat=1, seed=5)
sat = sc.sats[0]
sat.cn0 = 50
s = init_settings("BDS_B2a", acqSatelliteList=[sat.prn], acqNonCohTime=2, msToProcess=80, numberOfChannels=1, pilotTRKflag=1, CNo_VSMinterval=20)
raw = synth.make_record(sc, 18000 * 90)
eng = Engine(s)
eng.set_record(raw)
acq = eng.acquire([sat.prn])
ch = preRun(acq, s)
tr, _ = tracking(None, ch, s, engine=eng)
print("B2a acquire + pilot track + CNo/PLD", acq["carrFreq"][sat.prn - 1], tr[0]["epochsDone"], tr[0]["DataCNo"][-1])
eng.close()
print("sanitize run (round 2) ok")
