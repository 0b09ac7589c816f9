"""One small seeded case per signal folder of the reference, shared by the fixture generator (tests/golden/make_golden.py), the
CPU test that re-runs the oracle against the committed fixtures and the GPU test that runs the CUDA path against them.

Each case is acquisition on a short synthetic record plus a short tracking run from hand-off values derived from the scene (the
way tests/test_gpu_parity.py builds them).  The records are regenerated from their seed (numpy Generator streams are stable), so a
fixture holds only the outputs and the SHA-256 of the record bytes it was made from.

The fixtures are frozen ORACLE outputs - the reference is MATLAB-only and cannot run here - so they do not pin the restatement
to the reference; they pin both implementations (and later refactors of either) to one committed set of numbers per folder."""
import hashlib
import os

import numpy as np

import np_oracle as O
from cu_sdr_collection_b200 import init_settings, preRun, synth
from helpers import oracle_codes, oracle_signal_codes, orc_set_e1_codes, to_oracle_settings

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ACQ_KEYS = ("carrFreq", "codePhase", "peakMetric")
TRK_KEYS = ("absoluteSample", "I_P", "Q_P", "I_E", "I_L", "Q_E", "Q_L", "carrFreq", "codeFreq", "remCodePhase")
SIGNALS = ("GPS_L1CA", "GLO_GL1", "GLO_GL2", "BDS_B3I", "GAL_E1C", "GPS_L5C", "GAL_E5a", "GAL_E5b", "BDS_B2a", "BDS_B1I", "GPS_L2C",
           "BDS_B1C")


class Case:
    """s / so: engine-side and oracle-side settings; raw_acq, raw_trk: int8 I,Q records; long_signal: what acquisition() gets;
    acq_oracle(): acqResults of the oracle; ch: channel hand-off (None = acquisition only); trk_oracle(): per-channel dicts."""
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def digest(self) -> str:
        h = hashlib.sha256(self.raw_acq.tobytes())
        if self.ch is not None:
            h.update(self.raw_trk.tobytes())
        return h.hexdigest()


def _handoff(sc, s, N, chips, chip_rate, grid, base=1, key="PRN"):
    """channel structs as acquisition + preRun would produce them: carrier on the fine-search grid, code phase from the scene."""
    ch = []
    for sat in sc.sats:
        start = (chips - sat.code_phase) * (s.samplingFreq / chip_rate)
        off = -s.freqSpacing * sat.prn if key == "K" else 0.0
        ch.append({key: sat.prn, "acquiredFreq": round((s.IF + off + sat.doppler) / grid) * grid, "codePhase": int(round(start)) % N + base,
                   "status": "T"})
    while len(ch) < s.numberOfChannels:
        ch.append({key: 0, "acquiredFreq": 0.0, "codePhase": 0, "status": "-"})
    return ch


def _cn0(sc, v):
    for x in sc.sats:
        x.cn0 = v
    return sc


def build(signal: str) -> Case:
    nw = os.cpu_count() or 1
    if signal == "GPS_L1CA":
        fs, nE = 4.092e6, 60
        sc = _cn0(synth.default_scene(fs=fs, nsat=3, seed=31), 47)
        sv = sorted({x.prn for x in sc.sats} | {4})
        s = init_settings(samplingFreq=fs, acqSatelliteList=sv, acqNonCohTime=3, msToProcess=nE, numberOfChannels=3)
        so = to_oracle_settings(s)
        N = O.samples_per_code(so)
        raw = synth.make_record(sc, N * (nE + 46))
        ch = _handoff(sc, s, N, 1023, 1.023e6, 25.0)
        return Case(signal=signal, s=s, so=so, sv=sv, raw_acq=raw[: 2 * N * 44], raw_trk=raw, long_signal=O.read_acq_signal(raw, so),
                    acq_oracle=lambda: O.acquisition(O.read_acq_signal(raw, so), so), ch=ch, trk_oracle=lambda: O.tracking(raw, ch, so), nE=nE)
    if signal in ("GLO_GL1", "GLO_GL2"):
        fs, nE = 2.4e6, 60
        spacing = 562.5e3 if signal == "GLO_GL1" else 437.5e3
        sc = _cn0(synth.default_scene_glo(fs=fs, nsat=3, seed=17, freqSpacing=spacing), 47)
        for x, k in zip(sc.sats, (-2, -1, 1)):                  # keep the channels inside the sampled band
            x.prn = k
        ks = [-2, -1, 0, 1]
        s = init_settings(signal, samplingFreq=fs, acqNonCohTime=4, acqSatelliteList=ks, msToProcess=nE, numberOfChannels=3)
        so = O.glo_settings(samplingFreq=fs, acqNonCohTime=4, acqSatelliteList=ks, freqSpacing=spacing, msToProcess=nE, numberOfChannels=3)
        N = O.samples_per_code(so)
        raw = synth.make_record(sc, N * (nE + 46))
        ch = _handoff(sc, s, N, 511, 0.511e6, 25.0, key="K")
        return Case(signal=signal, s=s, so=so, sv=ks, raw_acq=raw[: 2 * N * 44], raw_trk=raw, long_signal=O.read_acq_signal_glo(raw, so),
                    acq_oracle=lambda: O.acquisition_glo(O.read_acq_signal_glo(raw, so), so), ch=ch if signal == "GLO_GL1" else None,
                    trk_oracle=lambda: O.tracking_glo(raw, ch, so), nE=nE)
    if signal == "BDS_B3I":
        fs, nE = 18e6, 40
        sc = synth.default_scene_b3i(fs=fs, nsat=2, seed=9)
        for x, p in zip(sc.sats, (3, 41)):                      # one GEO (2 ms bits) and one NH-coded satellite
            x.prn, x.cn0 = p, 48
        sv = [3, 41, 7]
        s = init_settings("BDS_B3I", samplingFreq=fs, acqSatelliteList=sv, acqNonCohTime=3, msToProcess=nE, numberOfChannels=2)
        so = O.b3i_settings(samplingFreq=fs, acqSatelliteList=sv, acqNonCohTime=3, msToProcess=nE, numberOfChannels=2)
        N = 18000
        raw = synth.make_record(sc, N * (nE + 26))
        ch = _handoff(sc, s, N, 10230, 10.23e6, 25.0)
        for c in ch:                                             # BDS/B3I/include/preRun.m:71-73
            c["codeFreq"] = s.codeFreqBasis + (c["acquiredFreq"] - s.IF) / s.carrFreqBasis * s.codeFreqBasis
        return Case(signal=signal, s=s, so=so, sv=sv, raw_acq=raw[: 2 * N * 24], raw_trk=raw, long_signal=O.read_acq_signal_b3i(raw, so),
                    acq_oracle=lambda: O.acquisition_b3i(O.read_acq_signal_b3i(raw, so), so), ch=ch, trk_oracle=lambda: O.tracking_b3i(raw, ch, so),
                    nE=nE)
    if signal == "GAL_E1C":
        fs, nE = 4.092e6, 20
        sc = _cn0(synth.default_scene_e1c({}, fs=fs, nsat=2, seed=4), 48)
        sv = sorted({x.prn for x in sc.sats} | {7})
        codes = sc.codes = oracle_signal_codes("GAL_E1C", sv)
        s = init_settings("GAL_E1C", samplingFreq=fs, acqSatelliteList=sv, acqSearchBand=4500.0, msToProcess=4 * nE, numberOfChannels=2,
                          pilotTRKflag=1, CNo_VSMinterval=10)
        so = to_oracle_settings(s)
        so.pilotTRKflag = s.pilotTRKflag
        orc_set_e1_codes(codes)
        bits = oracle_codes(codes)                              # the NumPy oracle takes the E1-B / E1-C tables as 0/1 bits
        N = O.samples_per_code(so)
        raw = synth.make_record(sc, N * (nE + 24) + 64)
        ch = _handoff(sc, s, N, 4092, 1.023e6, 10.0)
        long_signal = (raw[0: 2 * N * 42: 2] + 1j * raw[1: 2 * N * 42: 2]).astype(np.complex128)
        return Case(signal=signal, s=s, so=so, sv=sv, raw_acq=raw[: 2 * N * 42], raw_trk=raw, long_signal=long_signal,
                    acq_oracle=lambda: O.acquisition_e1c(long_signal, so, bits, workers=nw), ch=ch,
                    trk_oracle=lambda: O.tracking_e1c(raw, ch, so, bits), nE=nE)
    if signal in ("GPS_L5C", "GAL_E5a", "GAL_E5b", "BDS_B2a"):
        nE = 40
        sc = _cn0(synth.default_scene_fam5(signal, {}, fs=18e6, nsat=2, seed=5), 50)
        sv = sorted({x.prn for x in sc.sats} | {25})
        codes = sc.codes = oracle_signal_codes(signal, sv)
        kw = dict(acqSearchBand=4200.0, acqSearchStep=300.0) if signal == "GAL_E5b" else dict(acqSearchBand=4500.0)
        s = init_settings(signal, acqSatelliteList=sv, acqNonCohTime=3, msToProcess=nE, numberOfChannels=2, pilotTRKflag=1, CNo_VSMinterval=20, **kw)
        so = to_oracle_settings(s)
        N = 18000
        nper = max(O._FAM5_MINPER[signal], 5) + 2
        raw = synth.make_record(sc, N * max(nper, nE + 4))
        raw_acq = raw[: 2 * N * nper]
        ch = None
        if signal in ("GPS_L5C", "BDS_B2a"):
            acq = dict(carrFreq=np.zeros(63), codePhase=np.zeros(63), peakMetric=np.zeros(63))
            for i, sat in enumerate(sc.sats):
                start = (10230 - sat.code_phase) * (18e6 / 10.23e6)
                acq["carrFreq"][sat.prn - 1] = round((s.IF + sat.doppler) / 25.0) * 25.0
                acq["codePhase"][sat.prn - 1] = int(round(start)) % N + 1
                acq["peakMetric"][sat.prn - 1] = 10.0 - i
            ch = preRun(acq, s)
        return Case(signal=signal, s=s, so=so, sv=sv, codes=codes, raw_acq=raw_acq, raw_trk=raw, long_signal=O.read_acq_signal_fam5(raw_acq, so),
                    acq_oracle=lambda: O.acquisition_fam5(O.read_acq_signal_fam5(raw_acq, so), so, codes, workers=nw), ch=ch,
                    trk_oracle=lambda: O.tracking_fam5(raw, O.preRun_fam5(acq, so), so, codes), nE=nE)
    if signal in ("BDS_B1I", "GPS_L2C"):
        b1i = signal == "BDS_B1I"
        fs = 4.092e6 if b1i else 2.046e6
        nE = 60 if b1i else 10
        sc = _cn0(synth.default_scene_varb(signal, {}, fs=fs, nsat=2, seed=3), 48 if b1i else 45)
        sv = sorted({x.prn for x in sc.sats} | {30})
        codes = sc.codes = oracle_signal_codes(signal, sv)
        s = init_settings(signal, samplingFreq=fs, acqSatelliteList=sv, msToProcess=nE * (1 if b1i else 20), numberOfChannels=2,
                          CNo_VSMinterval=20 if b1i else 5, **({} if b1i else dict(acqSearchBand=9.0)))
        so = to_oracle_settings(s)
        so.stepSize, so.acqStep = s.stepSize, s.acqStep
        N = O.samples_per_code(so)
        raw = synth.make_record(sc, N * (nE + (12 if b1i else 4)))
        if b1i:
            raw_acq = raw[: 2 * N * 11]
            long_signal = O.read_acq_signal_varb(raw_acq, so)
            acq_oracle = lambda: O.acquisition_b1i(long_signal, so, codes, workers=nw)
            ch = _handoff(sc, s, N, 2046, 2.046e6, 25.0)
            trk_oracle = lambda: O.tracking_b1i(raw, ch, so, codes)
        else:
            raw_acq = raw[: 2 * N * 3]
            long_signal = (raw_acq[0::2] + 1j * raw_acq[1::2]).astype(np.complex128)
            acq_oracle = lambda: O.acquisition_l2c(long_signal, so, codes, workers=nw)
            ch = _handoff(sc, s, N, 20460, 1.023e6, 12.5, base=0)
            trk_oracle = lambda: O.tracking_l2c(raw, ch, so, codes)
        return Case(signal=signal, s=s, so=so, sv=sv, codes=codes, raw_acq=raw_acq, raw_trk=raw, long_signal=long_signal, acq_oracle=acq_oracle, ch=ch,
                    trk_oracle=trk_oracle, nE=nE)
    if signal == "BDS_B1C":
        fs, nE = 4.092e6, 12
        sc = _cn0(synth.default_scene_varb("BDS_B1C", {}, fs=fs, nsat=2, seed=3), 46)
        sv = sorted({x.prn for x in sc.sats} | {30})
        codes = sc.codes = oracle_signal_codes("BDS_B1C", sv)
        s = init_settings("BDS_B1C", samplingFreq=fs, acqSatelliteList=sv, acqSearchBand=4500.0, pilotACQflag=1, msToProcess=10 * nE,
                          numberOfChannels=2, CNo_VSMinterval=4)
        so = to_oracle_settings(s)
        so.acqStep, so.pilotACQflag, so.acqCohT = s.acqStep, s.pilotACQflag, s.acqCohT
        N = O.samples_per_code(so)
        raw = synth.make_record(sc, N * (nE + 2))
        raw_acq = raw[: 2 * N * 2]
        long_signal = (raw_acq[0::2] + 1j * raw_acq[1::2]).astype(np.complex128)
        acq = dict(carrFreq=np.zeros(63), codePhase=np.zeros(63), peakMetric=np.zeros(63))
        for i, sat in enumerate(sc.sats):
            start = (20460 - sat.code_phase) * (fs / 2.046e6)
            acq["carrFreq"][sat.prn - 1] = round((s.IF + sat.doppler) / 25.0) * 25.0
            acq["codePhase"][sat.prn - 1] = int(round(start)) % N + 1
            acq["peakMetric"][sat.prn - 1] = 20.0 - i
        ch = preRun(acq, s)
        return Case(signal=signal, s=s, so=so, sv=sv, codes=codes, raw_acq=raw_acq, raw_trk=raw, long_signal=long_signal,
                    acq_oracle=lambda: O.acquisition_b1c(long_signal, so, codes, workers=nw), ch=ch,
                    trk_oracle=lambda: O.tracking_b1c_nb(raw, ch, so, codes), nE=nE)
    raise ValueError(signal)


def oracle_outputs(case: Case) -> dict:
    """Flat dict of arrays: acq_<field>, and trk<i>_<field> for every live channel."""
    out = {"sha256": np.frombuffer(bytes.fromhex(case.digest()), dtype=np.uint8).copy(), "svList": np.asarray(case.sv)}
    a = case.acq_oracle()
    for k in ACQ_KEYS:
        out["acq_" + k] = np.asarray(a[k], dtype=np.float64)
    if case.ch is not None:
        tr = case.trk_oracle()
        for i, t in enumerate(tr):
            if t["status"] != "T":
                continue
            for k in TRK_KEYS:
                out[f"trk{i}_{k}"] = np.asarray(t[k], dtype=np.float64)
    return out


def fixture_path(signal: str) -> str:
    return os.path.join(GOLDEN_DIR, f"{signal.lower()}_case.npz")
