"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
the oracle on the same seeded bytes.  Bars: indices (acquired PRN set, code phase, Doppler bin,
carrFreq grid value, absoluteSample) bit-exact; I/Q correlator sums within 1e-6 of |I_P + iQ_P|;
peakMetric within 1e-6 relative."""
import os

import numpy as np
import pytest

import np_oracle as O
from cu_sdr_collection_b200 import Engine, acquisition, init_settings, preRun, synth, tracking
from helpers import (ROOT, c_acquisition, c_tracking, first_illconditioned_epoch, oracle_signal_codes, orc_set_e1_codes, scene,
                     to_oracle_settings, track_rel_err, windowed_iq_compare)
from cu_sdr_collection_b200.engine import GC_PARAM_TRACK_EXACT_SUMS

pytestmark = pytest.mark.gpu

IQ_TOL = 1e-6        # north_star: floating-point I_P/Q_P within 1e-6 relative
METRIC_TOL = 1e-6
# Variant B (BDS B1I, GPS L2C): peakMetric = peak / SECOND peak of one fp32 correlation row.  Round 1 gated it at 1e-5 on the argument
# that the second peak is a noise-floor value; measured, the worst relative error over the variant B cases is 2.3e-7
# (gpurun_out/r02_gputests.log: B1I 18 Msps 1.6e-7, L2C 2.046 Msps 2.3e-7, L2C 8 Msps 1.1e-7, B1I 4.092 Msps 1.0e-7), so the survey's
# 1e-6 gate applies here like everywhere else.  Each test prints its measured worst case.
VARB_METRIC_TOL = 1e-6
# The 60000-epoch closed-loop comparison needs correlator sums that follow the float64 reference far below 1e-6: whenever the code
# phase of a block start passes a sample-grid alignment, ~1000 samples of that block lie within 1e-6 chips of a chip edge at once,
# and a loop state that is off by 1e-10 chips (fp32 sums) flips one of them about once per channel-minute, after which the two
# closed loops separate for good (profiles/r02_parity_60000.md).  1 = run it in the float64 checking mode.
FULL_SIZE_EXACT = 0


def _acq_case(fs, nsat, seed, sv_extra, nonCoh=20, cn0=None, band=7000.0):
    sc = scene(fs, nsat=nsat, seed=seed, cn0=cn0)
    sv = sorted({x.prn for x in sc.sats} | set(sv_extra))
    s = init_settings(samplingFreq=fs, IF=20e3, acqNonCohTime=nonCoh, acqSearchBand=band, acqSatelliteList=sv)
    N = O.samples_per_code(to_oracle_settings(s))
    raw = synth.make_record(sc, N * max(42, nonCoh + 2) + 64)
    return sc, s, N, raw


def _check_acq(got, ref, sv):
    idx = np.array(sv) - 1
    assert np.array_equal(got["carrFreq"] != 0, ref["carrFreq"] != 0), "acquired PRN set differs"
    assert np.array_equal(got["coarseBin"][idx], ref["coarseBin"][idx]), "Doppler bin index differs"
    acq = ref["carrFreq"] != 0
    assert np.array_equal(got["coarseCodePhase"][acq], ref["coarseCodePhase"][acq]), "code phase differs"
    assert np.array_equal(got["codePhase"], ref["codePhase"])
    assert np.array_equal(got["carrFreq"], ref["carrFreq"]), "fine carrier frequency differs"
    rel = np.abs(got["peakMetric"][idx] - ref["peakMetric"][idx]) / ref["peakMetric"][idx]
    assert rel.max() < METRIC_TOL, rel.max()
    return rel.max()


def test_acquisition_fused_32736_vs_oracle():
    """BASELINE config 2 geometry (16.368 Msps, 2N = 32736, 29 bins, 20 blocks), PRN subset."""
    sc, s, N, raw = _acq_case(16.368e6, nsat=5, seed=20260101, sv_extra=[1, 2, 3])
    eng = Engine(s)
    got = eng.acquire(s.acqSatelliteList, host_iq=raw)
    st = eng.stats()
    assert st["acq_path"] == 1 and st["fft_len"] == 32736
    ref = c_acquisition(raw, s, s.acqSatelliteList)
    _check_acq(got, ref, s.acqSatelliteList)
    for sat in sc.sats:                                   # closed loop: injected signals come back
        assert got["carrFreq"][sat.prn - 1] != 0
        assert abs(got["carrFreq"][sat.prn - 1] - (s.IF + sat.doppler)) <= 25
    eng.close()


def test_acquisition_fused_matches_numpy_oracle_too():
    sc, s, N, raw = _acq_case(16.368e6, nsat=2, seed=5, sv_extra=[9], nonCoh=3)
    so = to_oracle_settings(s)
    ref = O.acquisition(O.read_acq_signal(raw, so), so)
    got = acquisition(raw, s, verbose=False)
    _check_acq(got, ref, s.acqSatelliteList)


@pytest.mark.parametrize("fs,nonCoh,path", [(2.046e6, 4, 0), (18e6, 2, 0), (4.092e6, 20, 0),
                                            (18e6, 3, 1), (16e6, 2, 1), (20e6, 2, 1), (16.368e6, 3, 2),
                                            (18e6, 3, 2), (16e6, 2, 2), (20e6, 2, 2), (12e6, 2, 2),
                                            (16.368e6, 7, 3), (18e6, 3, 3), (12e6, 12, 3)])
def test_acquisition_other_lengths(fs, nonCoh, path, monkeypatch):
    """Other FFT lengths: generic mixed-radix passes (4092, 8184, and 36000 when forced) and the fused
    C x 32 x 25 plans with a Cooley-Tukey column/row link (36000 = reference default 18 Msps, 32000, 40000,
    24000), each with the two-kernel inverse rows / inverse columns correlation stage (path 1, the default)
    and with the one-kernel cluster version (path 2, GC_ACQ_PATH=cluster) and the persistent work-queue version (path 3,
    GC_ACQ_PATH=queue: row and column items of one ordered queue, the work buffer a ring in L2)."""
    if path == 0:
        monkeypatch.setenv("GC_FORCE_GENERIC", "1")
    if path == 2:
        monkeypatch.setenv("GC_ACQ_PATH", "cluster")
    if path == 3:
        monkeypatch.setenv("GC_ACQ_PATH", "queue")
        monkeypatch.setenv("GC_Q_SLOTS", "5")           # a ring shorter than the grid, so that slots are reused and waited for
        monkeypatch.setenv("GC_Q_LAG", "2")
    sc, s, N, raw = _acq_case(fs, nsat=3, seed=13, sv_extra=[4], nonCoh=nonCoh, cn0=47, band=6000.0)
    eng = Engine(s)
    got = eng.acquire(s.acqSatelliteList, host_iq=raw)
    assert eng.stats()["acq_path"] == path
    ref = c_acquisition(raw, s, s.acqSatelliteList)
    _check_acq(got, ref, s.acqSatelliteList)
    eng.close()


def test_acquisition_noise_only_and_short_record():
    fs = 16.368e6
    sc = scene(fs, nsat=0, seed=3)
    s = init_settings(samplingFreq=fs, acqSatelliteList=[3, 17], acqNonCohTime=2)
    N = 16368
    raw = synth.make_record(sc, N * 42)
    eng = Engine(s)
    got = eng.acquire(s.acqSatelliteList, host_iq=raw)
    assert np.all(got["carrFreq"] == 0) and np.all(got["codePhase"] == 0)
    ref = c_acquisition(raw, s, s.acqSatelliteList)
    assert np.allclose(got["peakMetric"], ref["peakMetric"], rtol=METRIC_TOL)
    with pytest.raises(Exception, match="shorter"):
        eng.acquire(s.acqSatelliteList, host_iq=raw[: 2 * N * 30])
    eng.close()


def _track_case(fs, nEpochs, nsat=3, seed=21, cn0=46):
    sc = scene(fs, nsat=nsat, seed=seed, cn0=cn0)
    s = init_settings(samplingFreq=fs, IF=20e3, msToProcess=nEpochs, numberOfChannels=nsat + 1)
    N = O.samples_per_code(to_oracle_settings(s))
    raw = synth.make_record(sc, N * (nEpochs + 6))
    # channel hand-off values as acquisition would give them (exact Doppler grid / code start)
    prn, af, cp = [], [], []
    for sat in sc.sats:
        prn.append(sat.prn)
        af.append(round((s.IF + sat.doppler) / 25.0) * 25.0)
        start = (1023 - sat.code_phase) * (fs / 1.023e6)
        cp.append(float(int(round(start)) % N + 1))
    prn.append(0); af.append(0.0); cp.append(0.0)       # an unused channel (PRN 0, tracking.m:136)
    return sc, s, N, raw, prn, af, cp


@pytest.mark.parametrize("fs,nEpochs", [(16.368e6, 400), (2.046e6, 1500), (18e6, 100)])
def test_tracking_vs_oracle(fs, nEpochs):
    sc, s, N, raw, prn, af, cp = _track_case(fs, nEpochs)
    eng = Engine(s)
    eng.set_record(raw)
    out, vv, vi, done = eng.track(prn, af, cp, nEpochs)
    ref, rvv, rvi, rdone = c_tracking(raw, s, prn, af, cp, nEpochs)
    assert np.array_equal(done, rdone) and np.all(done[:-1] == nEpochs) and done[-1] == 0
    live = np.array(prn) != 0
    assert np.array_equal(out[live][:, 0], ref[live][:, 0]), "absoluteSample (block boundaries) differ"
    errs = track_rel_err(out[live], ref[live])
    for f in ("I_P", "Q_P", "I_E", "Q_E", "I_L", "Q_L"):
        assert errs[f] < IQ_TOL, (f, errs[f])
    for f in ("codeFreq", "carrFreq"):
        assert errs[f] < 1e-6, (f, errs[f])
    # quantities that pass through zero: absolute agreement (chips, rad, cycles)
    for i, f in ((13, "remCodePhase"), (14, "remCarrPhase"), (11, "pllDiscr"), (9, "dllDiscr")):
        d = np.abs(out[live][:, i] - ref[live][:, i])
        d = np.minimum(d, np.abs(d - 2 * np.pi)) if f == "remCarrPhase" else d
        assert d.max() < 2e-6, (f, d.max())
    assert np.allclose(vv[live], rvv[live], rtol=1e-5) and np.array_equal(vi, rvi)
    # untouched channel keeps the reference's initial fill (tracking.m:51-77)
    assert np.all(out[-1, 0] == 0) and np.all(out[-1, 3:9] == 0) and np.all(np.isinf(out[-1, 1]))
    # lock on every injected satellite
    for c in range(len(sc.sats)):
        assert np.mean(np.abs(out[c, 3, 50:])) > 3 * np.mean(np.abs(out[c, 7, 50:]))
    eng.close()


def test_tracking_short_record_semantics():
    sc, s, N, raw, prn, af, cp = _track_case(16.368e6, 120)
    short = raw[: 2 * N * 70]
    eng = Engine(s)
    eng.set_record(short)
    out, vv, vi, done = eng.track(prn, af, cp, 120)
    ref, rvv, rvi, rdone = c_tracking(short, s, prn, af, cp, 120)
    assert np.array_equal(done, rdone) and 0 < done[0] < 120 and np.all(done[1:] == 0)
    assert np.all(np.isinf(out[1:, 2])) and np.all(out[1:, 3] == 0)
    n0 = done[0]
    assert track_rel_err(out[:1, :, :n0], ref[:1, :, :n0])["I_P"] < IQ_TOL
    assert np.all(np.isinf(out[0, 2, n0:]))
    eng.close()


def test_reference_facing_wrappers_roundtrip(tmp_path):
    """acquisition() -> preRun() -> tracking(fid, ...) exactly as postProcessing.m:100-124 chains them."""
    fs = 16.368e6
    sc = scene(fs, nsat=3, seed=99, cn0=47)
    s = init_settings(samplingFreq=fs, msToProcess=200, numberOfChannels=4,
                      acqSatelliteList=sorted({x.prn for x in sc.sats} | {6}), acqNonCohTime=5)
    N = 16368
    raw = synth.make_record(sc, N * 260)
    path = tmp_path / "if.bin"
    raw.tofile(path)
    so = to_oracle_settings(s)
    longSignal = O.read_acq_signal(raw, so)               # complex double row vector, as MATLAB passes it
    acq = acquisition(longSignal, s, verbose=True)
    ch = preRun(acq, s)
    assert sorted(c["PRN"] for c in ch if c["PRN"]) == sorted(x.prn for x in sc.sats)
    with open(path, "rb") as fid:
        tr, ch2 = tracking(fid, ch, s)
    ref_acq = O.acquisition(longSignal, so)
    ref_ch = O.preRun(ref_acq, so)
    assert [c["PRN"] for c in ch] == [c["PRN"] for c in ref_ch]
    ref_tr = O.tracking(raw, ref_ch, so)
    for i, c in enumerate(ch):
        if c["PRN"] == 0:
            assert tr[i]["status"] == "-"
            continue
        assert tr[i]["status"] == "T" and tr[i]["PRN"] == c["PRN"]
        sc_ = np.hypot(ref_tr[i]["I_P"], ref_tr[i]["Q_P"])
        assert np.max(np.abs(tr[i]["I_P"] - ref_tr[i]["I_P"]) / sc_) < IQ_TOL
        assert np.max(np.abs(tr[i]["Q_P"] - ref_tr[i]["Q_P"]) / sc_) < IQ_TOL
        assert np.array_equal(tr[i]["absoluteSample"], ref_tr[i]["absoluteSample"])
        assert np.allclose(tr[i]["CNo"]["VSMValue"], ref_tr[i]["VSMValue"], rtol=1e-5)


def test_golden_fixture_through_cuda():
    g = np.load(os.path.join(ROOT, "tests", "golden", "l1ca_small.npz"))
    s = init_settings(samplingFreq=float(g["fs"]), IF=float(g["IF"]), acqNonCohTime=int(g["nonCoh"]),
                      acqSearchBand=float(g["band"]), msToProcess=int(g["nEpochs"]), numberOfChannels=int(g["nCh"]),
                      acqSatelliteList=[int(p) for p in g["svList"]])
    eng = Engine(s)
    got = eng.acquire(s.acqSatelliteList, host_iq=g["raw"])
    assert np.array_equal(got["carrFreq"], g["carrFreq"]) and np.array_equal(got["codePhase"], g["codePhase"])
    idx = g["svList"] - 1
    assert np.allclose(got["peakMetric"][idx], g["peakMetric"][idx], rtol=METRIC_TOL)
    eng.set_record(g["raw"])
    out, vv, vi, done = eng.track(g["chPRN"], g["chFreq"], g["chCodePhase"], s.msToProcess)
    live = g["chPRN"] != 0
    errs = track_rel_err(out[live], g["track"][live])
    assert errs["I_P"] < IQ_TOL and errs["Q_P"] < IQ_TOL and errs["absoluteSample"] == 0
    eng.close()


@pytest.mark.parametrize("signal", __import__("golden_cases").SIGNALS)
def test_golden_signal_fixture_through_cuda(signal, tmp_path):
    """The committed per-folder fixtures (tests/golden/<signal>_case.npz, frozen oracle outputs) against the CUDA path through
    the reference-facing wrappers: acquisition indices exact and peakMetric to 1e-6 (variant B: 1e-5), tracking absoluteSample
    exact, correlator sums to 1e-6 of |P|, NCO rows to 1e-4 Hz / 1e-6 chips."""
    import golden_cases as G
    case = G.build(signal)
    g = np.load(G.fixture_path(signal))
    assert bytes(g["sha256"]).hex() == case.digest()
    s = case.s
    eng = Engine(s)
    got = acquisition(case.long_signal, s, engine=eng, verbose=False)
    n = min(got["carrFreq"].size, g["acq_carrFreq"].size)
    assert np.array_equal(got["carrFreq"][:n], g["acq_carrFreq"][:n]) and np.array_equal(got["codePhase"][:n], g["acq_codePhase"][:n])
    nz = g["acq_peakMetric"][:n] != 0
    rel = np.abs(got["peakMetric"][:n][nz] - g["acq_peakMetric"][:n][nz]) / g["acq_peakMetric"][:n][nz]
    assert rel.max() < (VARB_METRIC_TOL if signal in ("BDS_B1I", "GPS_L2C") else METRIC_TOL), rel.max()
    if case.ch is not None:
        path = tmp_path / "rec.bin"
        case.raw_trk.tofile(path)
        with open(path, "rb") as fid:
            tr, _ = tracking(fid, case.ch, s, engine=eng)
        live = [i for i in range(len(tr)) if f"trk{i}_I_P" in g.files]
        assert live and all(tr[i]["status"] == "T" and tr[i]["epochsDone"] == case.nE for i in live)
        for i in live:
            ref = {k: g[f"trk{i}_{k}"] for k in G.TRK_KEYS}
            if signal == "GPS_L2C":                           # fractional sample positions (GPS_L2C tracking.m:223)
                assert np.max(np.abs(tr[i]["absoluteSample"] - ref["absoluteSample"])) < 1e-5
            else:
                assert np.array_equal(tr[i]["absoluteSample"], ref["absoluteSample"])
            scale = np.hypot(ref["I_P"], ref["Q_P"])
            for name in ("I_P", "Q_P", "I_E", "I_L", "Q_E", "Q_L"):
                assert np.max(np.abs(tr[i][name] - ref[name]) / scale) < IQ_TOL, (signal, i, name)
            assert np.max(np.abs(tr[i]["carrFreq"] - ref["carrFreq"])) < 1e-4 and np.max(np.abs(tr[i]["codeFreq"] - ref["codeFreq"])) < 1e-4
            assert np.max(np.abs(tr[i]["remCodePhase"] - ref["remCodePhase"])) < 1e-6
    eng.close()


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_tracking_cluster_variants_agree(cluster, monkeypatch):
    """A channel spread over a thread-block cluster of 1/2/4/8 CTAs gives the same results."""
    monkeypatch.setenv("GC_TRACK_CLUSTER", str(cluster))
    sc, s, N, raw, prn, af, cp = _track_case(16.368e6, 150, nsat=2, seed=31)
    eng = Engine(s)
    eng.set_record(raw)
    out, vv, vi, done = eng.track(prn, af, cp, 150)
    ref, rvv, rvi, rdone = c_tracking(raw, s, prn, af, cp, 150)
    assert np.array_equal(done, rdone)
    live = np.array(prn) != 0
    assert np.array_equal(out[live][:, 0], ref[live][:, 0])
    errs = track_rel_err(out[live], ref[live])
    for f in ("I_P", "Q_P", "I_E", "Q_E", "I_L", "Q_L"):
        assert errs[f] < IQ_TOL, (cluster, f, errs[f])
    # record that ends mid-run: same early-stop semantics in every variant
    short = raw[: 2 * N * 100]
    eng.set_record(short)
    out2, _, _, done2 = eng.track(prn, af, cp, 150)
    _, _, _, rdone2 = c_tracking(short, s, prn, af, cp, 150)
    assert np.array_equal(done2, rdone2)
    eng.close()


def test_full_size_acquisition_grid_vs_oracle():
    """BASELINE configs[1] at full size: 32 PRN x 29 Doppler x 20 blocks, 2N = 32736; every index exact."""
    sc = scene(16.368e6, nsat=10, seed=20260101)
    for sat in sc.sats:
        sat.cn0 = max(sat.cn0, 43.0)
    s = init_settings(samplingFreq=16.368e6)
    raw = synth.make_record(sc, 16368 * 42 + 64)
    eng = Engine(s)
    got = eng.acquire(s.acqSatelliteList, host_iq=raw)
    ref = c_acquisition(raw, s, s.acqSatelliteList)
    worst = _check_acq(got, ref, s.acqSatelliteList)
    assert int(np.sum(got["carrFreq"] != 0)) >= 8
    # PRNs that are not acquired too: the coarse code phase of a noise-only cell is the arg max of 32736 nearly equal values
    assert np.array_equal(got["coarseCodePhase"], ref["coarseCodePhase"])
    print(f"[parity] full 32 x 29 grid: worst peakMetric error {worst:.2e}")
    eng.close()


def test_full_size_tracking_properties():
    """BASELINE configs[2] at full size (12 channels x 60000 ms on a 60 s record generated on the GPU):
    size-independent properties + the C oracle on every one of the 12 x 60000 epochs."""
    import torch
    fs, nms = 16.368e6, 60000
    sc = scene(fs, nsat=12, seed=77)
    for sat in sc.sats:
        sat.cn0 = max(sat.cn0, 42.0)
    s = init_settings(samplingFreq=fs, msToProcess=nms, numberOfChannels=12)
    N = 16368
    rec = synth.make_record_torch(sc, N * (nms + 40), device="cuda")
    eng = Engine(s)
    eng.set_record(rec)
    acq = eng.acquire()
    ch = preRun(acq, s)
    found = {c["PRN"] for c in ch if c["PRN"]}
    assert found == {x.prn for x in sc.sats}
    prn = [c["PRN"] for c in ch]; af = [c["acquiredFreq"] for c in ch]; cp = [float(c["codePhase"]) for c in ch]
    eng.set_param(GC_PARAM_TRACK_EXACT_SUMS, FULL_SIZE_EXACT)
    out, vv, vi, done = eng.track(prn, af, cp, nms)
    assert np.all(done == nms)
    sat_of = {x.prn: x for x in sc.sats}
    for c in range(12):
        sat = sat_of[prn[c]]
        a = out[c, 0]
        blk = np.diff(a)
        assert np.all((blk >= N - 2) & (blk <= N + 2)), "block sizes"               # one code period per epoch
        # code-phase bookkeeping closes: block starts follow the injected code Doppler over the whole minute
        drift = (a[-1] - a[0]) - (nms - 1) * N
        expect = -(nms - 1) * N * sat.doppler / 1575.42e6
        assert abs(drift - expect) < 3.0, (drift, expect)
        # carrier NCO sits on the injected Doppler
        assert abs(np.mean(out[c, 2, 1000:]) - (s.IF + sat.doppler)) < 1.0
        # lock: prompt I carries the power, and its sign over 20 ms bit periods is the injected data (up to polarity)
        Ip, Qp = out[c, 3, 2000:], out[c, 7, 2000:]
        assert np.mean(np.abs(Ip)) > 4 * np.mean(np.abs(Qp))
        assert np.all(np.hypot(out[c, 3], out[c, 7]) > 0)
        # C/N0 estimate within 3 dB of the injected value
        assert abs(np.median(vv[c, 10:]) - sat.cn0) < 3.0, (np.median(vv[c, 10:]), sat.cn0)
    # the same bytes through the C oracle for ALL 60000 epochs of all 12 channels (OpenMP over channels, about half a minute):
    # block boundaries exact, every correlator sum within 1e-6 of |P|, the NCO rows at their own bars
    raw = rec.cpu().numpy()
    ref, rvv, rvi, rdone = c_tracking(raw, s, prn, af, cp, nms, parallel=1)
    assert np.array_equal(done, rdone)
    assert np.array_equal(out[:, 0], ref[:, 0]), "absoluteSample differs somewhere in the minute"
    errs = track_rel_err(out, ref)
    print("[parity] configs[2] 12 x 60000 epochs vs C oracle:", {k: float("%.3g" % v) for k, v in errs.items()})
    for f in ("I_P", "Q_P", "I_E", "Q_E", "I_L", "Q_L"):
        assert errs[f] < IQ_TOL, (f, errs[f])
    assert errs["codeFreq"] < 1e-9 and errs["carrFreq"] < 1e-6, errs
    d = np.abs(out[:, 14] - ref[:, 14])
    d = np.minimum(d, np.abs(d - 2 * np.pi))
    assert d.max() < 1e-6 and np.abs(out[:, 13] - ref[:, 13]).max() < 1e-7, (d.max(), np.abs(out[:, 13] - ref[:, 13]).max())
    assert np.allclose(vv, rvv, rtol=1e-5) and np.array_equal(vi, rvi)
    from helpers import PARITY_REPORT
    PARITY_REPORT.append(("GPS_L1CA configs[2] 12 ch x 60000 ms vs C oracle (no window needed)", nms, nms, max(errs["I_P"], errs["Q_P"]), 0.0))
    eng.close()


# ------------------------------------------------------------------------------- GLONASS (GLO_GL1 / GLO_GL2)
from cu_sdr_collection_b200.engine import GC_SV_NONE  # noqa: E402


def _glo_check_acq(got, ref, ks):
    idx = np.array(ks) + 7
    assert got["carrFreq"].shape == (21,)
    assert np.array_equal(got["carrFreq"] != 0, ref["carrFreq"] != 0), "acquired channel set differs"
    assert np.array_equal(got["coarseBin"][idx], ref["coarseBin"][idx])
    assert np.array_equal(got["codePhase"], ref["codePhase"])
    assert np.array_equal(got["carrFreq"], ref["carrFreq"])
    rel = np.abs(got["peakMetric"][idx] - ref["peakMetric"][idx]) / ref["peakMetric"][idx]
    assert rel.max() < METRIC_TOL, rel.max()


@pytest.mark.parametrize("signal,fs,nonCoh", [("GLO_GL1", 12e6, 20), ("GLO_GL2", 2.4e6, 4)])
def test_glonass_acquisition_vs_oracle(signal, fs, nonCoh):
    """GLO_GL1 at the reference defaults (12 Msps, FFT length 24000, 14 frequency channels x 21 bins x 20
    blocks) and GLO_GL2 spacing at a small rate; both through the generic mixed-radix path."""
    spacing = 562.5e3 if signal == "GLO_GL1" else 437.5e3
    sc = synth.default_scene_glo(fs=fs, nsat=4, seed=17, freqSpacing=spacing)
    for x in sc.sats:
        x.cn0 = 47
    if fs < 5e6:                                           # keep the channels inside the sampled band
        for i, x in enumerate(sc.sats):
            x.prn = [-2, -1, 1, 2][i]
    ks = sorted({x.prn for x in sc.sats} | {0}) if fs < 5e6 else list(range(-7, 7))
    s = init_settings(signal, samplingFreq=fs, acqNonCohTime=nonCoh, acqSatelliteList=ks)
    so = O.glo_settings(samplingFreq=fs, acqNonCohTime=nonCoh, acqSatelliteList=ks, freqSpacing=spacing)
    N = O.samples_per_code(so)
    raw = synth.make_record(sc, N * max(42, nonCoh + 2) + 64)
    eng = Engine(s)
    got = eng.acquire(ks, host_iq=raw)
    ref = c_acquisition(raw, so, ks)
    _glo_check_acq(got, ref, ks)
    for sat in sc.sats:
        assert got["carrFreq"][sat.prn + 7] != 0
        assert abs(got["carrFreq"][sat.prn + 7] - (s.IF - spacing * sat.prn + sat.doppler)) <= 25
    eng.close()


def test_glonass_tracking_and_wrappers_vs_oracle(tmp_path):
    """acquisition() -> preRun() -> tracking() on a GLONASS record, against the GLONASS oracle."""
    fs, nms = 12e6, 300
    sc = synth.default_scene_glo(fs=fs, nsat=3, seed=23)
    for x in sc.sats:
        x.cn0 = 47
    ks = sorted({x.prn for x in sc.sats} | {5, -6})
    s = init_settings("GLO_GL1", samplingFreq=fs, acqSatelliteList=ks, acqNonCohTime=6, msToProcess=nms, numberOfChannels=4)
    so = O.glo_settings(samplingFreq=fs, acqSatelliteList=ks, acqNonCohTime=6, msToProcess=nms, numberOfChannels=4)
    N = 12000
    raw = synth.make_record(sc, N * (nms + 50))
    path = tmp_path / "glo.bin"
    raw.tofile(path)
    longSignal = O.read_acq_signal_glo(raw, so)            # Q + 1i*I, as the GLONASS postProcessing.m builds it
    acq = acquisition(longSignal, s, verbose=True)
    ref_acq = O.acquisition_glo(longSignal, so)
    _glo_check_acq(acq, ref_acq, ks)
    ch = preRun(acq, s)
    ref_ch = O.preRun_glo(ref_acq, so)
    assert [(c["K"], c["status"]) for c in ch] == [(c["K"], c["status"]) for c in ref_ch]
    assert sorted(c["K"] for c in ch if c["status"] == "T") == sorted(x.prn for x in sc.sats)
    with open(path, "rb") as fid:
        tr, _ = tracking(fid, ch, s)
    sv = [c["K"] if c["status"] != "-" else GC_SV_NONE for c in ch]
    ref, rvv, rvi, rdone = c_tracking(raw, so, sv, [c["acquiredFreq"] for c in ch], [float(c["codePhase"]) for c in ch], nms)
    for i, c in enumerate(ch):
        if c["status"] == "-":
            assert tr[i]["status"] == "-" and tr[i]["epochsDone"] == 0 and np.all(tr[i]["I_P"] == 0)
            continue
        assert tr[i]["status"] == "T" and tr[i]["PRN"] == c["K"] and tr[i]["epochsDone"] == nms
        assert np.array_equal(tr[i]["absoluteSample"], ref[i, 0])
        sc_ = np.hypot(ref[i, 3], ref[i, 7])
        for f, name in ((3, "I_P"), (7, "Q_P"), (4, "I_E"), (5, "I_L"), (6, "Q_E"), (8, "Q_L")):
            assert np.max(np.abs(tr[i][name] - ref[i, f]) / sc_) < IQ_TOL, name
        assert np.max(np.abs(tr[i]["carrFreq"] - ref[i, 2])) < 1e-4
        assert np.allclose(tr[i]["CNo"]["VSMValue"], rvv[i], rtol=1e-5)
    # one-CTA and 8-CTA-cluster variants agree for GLONASS too
    eng = Engine(s)
    eng.set_record(raw)
    out8, _, _, d8 = eng.track(sv, [c["acquiredFreq"] for c in ch], [float(c["codePhase"]) for c in ch], nms)
    os.environ["GC_TRACK_CLUSTER"] = "1"
    try:
        out1, _, _, d1 = eng.track(sv, [c["acquiredFreq"] for c in ch], [float(c["codePhase"]) for c in ch], nms)
    finally:
        del os.environ["GC_TRACK_CLUSTER"]
    assert np.array_equal(d1, d8) and np.array_equal(out1[:, 0], out8[:, 0])
    live = np.array(sv) != GC_SV_NONE
    assert track_rel_err(out1[live], out8[live])["I_P"] < IQ_TOL
    eng.close()


# ------------------------------------------------------------------------------- BeiDou B3I (BDS/B3I)
def test_b3i_acquisition_tracking_and_wrappers_vs_oracle(tmp_path):
    """BDS/B3I at the reference's sampling rate (18 Msps, FFT length 36000, fused 45 x 32 x 25 plan):
    acquisition() with the NH-code / GEO fine search, preRun() with the carrier-aided code NCO centre,
    tracking() with the 3-coefficient carrier filter, against the B3I oracle."""
    fs, nms = 18e6, 200
    sc = synth.default_scene_b3i(fs=fs, nsat=4, seed=9)
    for x, p in zip(sc.sats, (3, 20, 41, 60)):           # two GEO (2 ms bits) and two NH-coded satellites
        x.prn, x.cn0 = p, 48
    sv = [3, 20, 41, 60, 7, 33]
    s = init_settings("BDS_B3I", samplingFreq=fs, acqSatelliteList=sv, acqNonCohTime=5, msToProcess=nms, numberOfChannels=5)
    so = O.b3i_settings(samplingFreq=fs, acqSatelliteList=sv, acqNonCohTime=5, msToProcess=nms, numberOfChannels=5)
    N = 18000
    raw = synth.make_record(sc, N * (nms + 40))
    path = tmp_path / "b3i.bin"
    raw.tofile(path)
    longSignal = O.read_acq_signal_b3i(raw, so)
    acq = acquisition(longSignal, s, verbose=True)
    assert acq["carrFreq"].shape == (63,)
    ref = c_acquisition(raw, so, sv)
    _check_acq(acq, ref, sv)
    assert {p for p in sv if acq["carrFreq"][p - 1] != 0} == {3, 20, 41, 60}
    ch = preRun(acq, s)
    ref_ch = O.preRun_b3i(ref, so)
    assert [(c["PRN"], c["codeFreq"]) for c in ch] == [(c["PRN"], c["codeFreq"]) for c in ref_ch]
    with open(path, "rb") as fid:
        tr, _ = tracking(fid, ch, s)
    rout, rvv, rvi, rdone = c_tracking(raw, so, [c["PRN"] for c in ch], [c["acquiredFreq"] for c in ch],
                                       [float(c["codePhase"]) for c in ch], nms, code_freq0=[c["codeFreq"] for c in ch])
    for i, c in enumerate(ch):
        if c["PRN"] == 0:
            assert tr[i]["status"] == "-" and tr[i]["epochsDone"] == 0
            continue
        assert tr[i]["status"] == "T" and tr[i]["epochsDone"] == nms == rdone[i]
        assert np.array_equal(tr[i]["absoluteSample"], rout[i, 0])
        assert tr[i]["codeFreq"][0] == c["codeFreq"]
        sc_ = np.hypot(rout[i, 3], rout[i, 7])
        for f, name in ((3, "I_P"), (7, "Q_P"), (4, "I_E"), (5, "I_L"), (6, "Q_E"), (8, "Q_L")):
            assert np.max(np.abs(tr[i][name] - rout[i, f]) / sc_) < IQ_TOL, name
        assert np.max(np.abs(tr[i]["codeFreq"] - rout[i, 1])) < 1e-4
        assert np.allclose(tr[i]["CNo"]["VSMValue"], rvv[i], rtol=1e-5)


# ------------------------------------------------------------------------------- Galileo E1 (GAL/GAL_E1C)
def _e1c_case(fs, nsat, seed, extra, band, ms, nch, cn0=48, **kw):
    sc = synth.default_scene_e1c({}, fs=fs, nsat=nsat, seed=seed)
    for x in sc.sats:
        x.cn0 = cn0
    sv = sorted({x.prn for x in sc.sats} | set(extra))
    # the real E1-B / E1-C memory codes, from the oracle's generators; the engine is given none and decodes its own on the device
    codes = sc.codes = oracle_signal_codes("GAL_E1C", sv)
    s = init_settings("GAL_E1C", samplingFreq=fs, acqSatelliteList=sv, acqSearchBand=band, msToProcess=ms, numberOfChannels=nch, **kw)
    so = to_oracle_settings(s)
    so.pilotTRKflag = s.pilotTRKflag
    orc_set_e1_codes(codes)
    return codes, sc, s, so, sv


@pytest.mark.parametrize("fs,band,generic", [(4.092e6, 4500.0, 0), (18e6, 4200.0, 0), (18e6, 4200.0, 1), (20e6, 4200.0, 0)])
def test_e1c_acquisition_vs_oracle(fs, band, generic, monkeypatch):
    """GAL_E1C acquisition (E1B + E1C BOC(1,1) replicas summed, 10 Hz fine search over 25 periods against the
    25-chip secondary code) on caller-supplied memory codes: a small rate and the reference's 18 Msps
    (FFT length 144000: fused 180 x 32 x 25 plan with two-level columns; GC_FORCE_GENERIC covers the generic passes)."""
    if generic:
        monkeypatch.setenv("GC_FORCE_GENERIC", "1")
    codes, sc, s, so, sv = _e1c_case(fs, nsat=2, seed=4, extra=[7], band=band, ms=80, nch=2)
    N = O.samples_per_code(so)
    raw = synth.make_record(sc, N * 42 + 64)
    eng = Engine(s)
    got = eng.acquire(sv, host_iq=raw)
    assert got["carrFreq"].shape == (50,) and eng.stats()["fft_len"] == 2 * N
    assert eng.stats()["acq_path"] == (0 if generic else 1)      # 32736, 144000 and 160000 all have fused plans
    ref = c_acquisition(raw, s, sv)
    _check_acq(got, ref, sv)
    for sat in sc.sats:                                   # closed loop: injected signals come back on the 10 Hz grid
        assert got["carrFreq"][sat.prn - 1] != 0
        assert abs(got["carrFreq"][sat.prn - 1] - (s.IF + sat.doppler)) <= 10
        start = (4092 - sat.code_phase) * (fs / 1.023e6)
        assert abs((got["codePhase"][sat.prn - 1] - 1 - start + N / 2) % N - N / 2) <= 2
    assert got["carrFreq"][7 - 1] == 0
    # gc_set_code is an override: codes handed over by the caller (here the same ones, for one SV) give the same results
    eng2 = Engine(s, codes={sv[0]: codes[sv[0]]})
    got2 = eng2.acquire(sv, host_iq=raw)
    assert np.array_equal(got2["carrFreq"], got["carrFreq"]) and np.array_equal(got2["peakMetric"], got["peakMetric"])
    eng2.close()
    eng.close()


@pytest.mark.parametrize("fs,nE,pilot", [(4.092e6, 120, 1), (4.092e6, 60, 0), (18e6, 25, 1)])
def test_e1c_tracking_and_wrappers_vs_oracle(fs, nE, pilot, tmp_path):
    """acquisition() -> preRun() -> tracking() on a Galileo E1 record: 4 ms epochs, ceil(tcode*2) sub-chip
    tables, data + pilot discriminators averaged (pilotTRKflag), against the E1C oracle."""
    codes, sc, s, so, sv = _e1c_case(fs, nsat=2, seed=4, extra=[], band=4500.0, ms=4 * nE, nch=3,
                                     pilotTRKflag=pilot, CNo_VSMinterval=20)
    N = O.samples_per_code(so)
    raw = synth.make_record(sc, N * (nE + 4) + 64)
    # channel hand-off values as acquisition would give them
    ch = []
    for sat in sc.sats:
        start = (4092 - sat.code_phase) * (fs / 1.023e6)
        ch.append(dict(PRN=sat.prn, acquiredFreq=round((s.IF + sat.doppler) / 10.0) * 10.0,
                       codePhase=int(round(start)) % N + 1, status="T"))
    ch.append(dict(PRN=0, acquiredFreq=0.0, codePhase=0, status="-"))
    path = tmp_path / "e1.bin"
    raw.tofile(path)
    eng = Engine(s)
    with open(path, "rb") as fid:
        tr, _ = tracking(fid, ch, s, engine=eng)
    rout, rvv, rvi, rdone = c_tracking(raw, s, [c["PRN"] for c in ch], [c["acquiredFreq"] for c in ch],
                                       [float(c["codePhase"]) for c in ch], nE)
    assert tr[2]["status"] == "-" and tr[2]["epochsDone"] == 0
    for i in range(2):
        assert tr[i]["status"] == "T" and tr[i]["epochsDone"] == nE == rdone[i]
        assert np.array_equal(tr[i]["absoluteSample"], rout[i, 0])
        sc_ = np.hypot(rout[i, 3], rout[i, 7])
        for f, name in ((3, "I_P"), (7, "Q_P"), (4, "I_E"), (5, "I_L"), (6, "Q_E"), (8, "Q_L")):
            assert np.max(np.abs(tr[i][name] - rout[i, f]) / sc_) < IQ_TOL, name
        assert np.max(np.abs(tr[i]["carrFreq"] - rout[i, 2])) < 1e-4
        assert np.max(np.abs(tr[i]["codeFreq"] - rout[i, 1])) < 1e-4
        assert np.max(np.abs(tr[i]["remCodePhase"] - rout[i, 13])) < 1e-6      # chips; follows codeFreq (fp32 discriminators)
        assert np.allclose(tr[i]["CNo"]["VSMValue"], rvv[i], rtol=1e-5)
        if nE >= 100:                                     # the loops have pulled in: prompt energy sits in I
            assert np.mean(np.abs(tr[i]["I_P"][60:])) > 4 * np.mean(np.abs(tr[i]["Q_P"][60:]))
    eng.close()


# ------------------------------------------------------------- GPS L5C, GAL E5a, GAL E5b, BDS B2a (10230-chip data + pilot)
def _fam5_case(signal, nsat, seed, extra, nonCoh, ms, nch, cn0=50, **kw):
    sc = synth.default_scene_fam5(signal, {}, fs=18e6, nsat=nsat, seed=seed)
    for x in sc.sats:
        x.cn0 = cn0
    sv = sorted({x.prn for x in sc.sats} | set(extra))
    codes = sc.codes = oracle_signal_codes(signal, sv)          # real ICD codes (oracle generators); the engine generates its own
    s = init_settings(signal, acqSatelliteList=sv, acqNonCohTime=nonCoh, msToProcess=ms, numberOfChannels=nch, **kw)
    so = to_oracle_settings(s)
    return codes, sc, s, so, sv


@pytest.mark.parametrize("signal,path", [("GPS_L5C", "split"), ("GPS_L5C", "cluster"), ("GPS_L5C", "generic"),
                                         ("GAL_E5a", "split"), ("GAL_E5b", "split"), ("BDS_B2a", "split")])
def test_fam5_acquisition_vs_oracle(signal, path, monkeypatch):
    """Two-replica variant-A acquisition (abs(ifft(X.*C_data)) + abs(ifft(X.*C_pilot))) at the reference's 18 Msps
    (FFT length 36000, fused 45 x 32 x 25 plan) with each signal's fine search: L5C NH20 over 20 periods, E5a the
    PRN's 100-chip secondary code over 100 periods on a 5 Hz grid, E5b none, B2a data + pilot non-coherent."""
    if path == "cluster":
        monkeypatch.setenv("GC_ACQ_PATH", "cluster")
    if path == "generic":
        monkeypatch.setenv("GC_FORCE_GENERIC", "1")
    kw = dict(acqSearchBand=4500.0)
    if signal == "GAL_E5b":
        kw = dict(acqSearchBand=4200.0, acqSearchStep=300.0)
    codes, sc, s, so, sv = _fam5_case(signal, nsat=2, seed=5, extra=[25], nonCoh=3, ms=60, nch=3, **kw)
    N = 18000
    raw = synth.make_record(sc, N * (max(O._FAM5_MINPER[signal], 5) + 2))
    longSignal = O.read_acq_signal_fam5(raw, so)
    eng = Engine(s)
    got = acquisition(longSignal, s, engine=eng, verbose=False)
    st = eng.stats()
    assert st["fft_len"] == 36000 and st["acq_path"] == {"split": 1, "cluster": 2, "generic": 0}[path]
    ref = O.acquisition_fam5(longSignal, so, codes, workers=os.cpu_count() or 1)
    assert got["carrFreq"].shape == ref["carrFreq"].shape
    _check_acq(got, ref, sv)
    for sat in sc.sats:
        # B2a sums |per-period sums| non-coherently, which barely resolves frequency inside a coarse bin
        step = {"GAL_E5a": 5, "GAL_E5b": 300, "BDS_B2a": 250}.get(signal, 25)
        assert got["carrFreq"][sat.prn - 1] != 0 and abs(got["carrFreq"][sat.prn - 1] - (s.IF + sat.doppler)) <= step
    assert got["carrFreq"][25 - 1] == 0
    eng.close()


@pytest.mark.parametrize("signal,pilot,nE,exact", [("GPS_L5C", 1, 1200, 0), ("GPS_L5C", 1, 1200, 1), ("GPS_L5C", 0, 240, 0), ("GAL_E5a", 1, 240, 0),
                                                   ("GAL_E5a", 1, 240, 1), ("GAL_E5b", 1, 240, 0), ("BDS_B2a", 1, 240, 0), ("BDS_B2a", 1, 240, 1)])
def test_fam5_tracking_and_wrappers_vs_oracle(signal, pilot, nE, exact, tmp_path):
    """preRun() (carrier-aided code NCO centre) -> tracking() with the quadrature pilot (prompt rotated by -pi/2
    before the atan, discriminators averaged, Pilot_I_P / Pilot_Q_P recorded), against the NumPy oracle.
    exact = the engine's float64 checking mode: 1e-6 over the whole run, no window."""
    codes, sc, s, so, sv = _fam5_case(signal, nsat=2, seed=5, extra=[], nonCoh=3, ms=nE, nch=3, pilotTRKflag=pilot,
                                      CNo_VSMinterval=40)
    N = 18000
    raw = synth.make_record(sc, N * (nE + 4))
    acq = dict(carrFreq=np.zeros(63), codePhase=np.zeros(63), peakMetric=np.zeros(63))
    for i, sat in enumerate(sc.sats):
        start = (10230 - sat.code_phase) * (18e6 / 10.23e6)
        acq["carrFreq"][sat.prn - 1] = round((s.IF + sat.doppler) / 25.0) * 25.0
        acq["codePhase"][sat.prn - 1] = int(round(start)) % N + 1
        acq["peakMetric"][sat.prn - 1] = 10.0 - i
    ch = preRun(acq, s)
    ref_ch = O.preRun_fam5(acq, so)
    assert [(c["PRN"], c["codeFreq"], c["status"]) for c in ch] == [(c["PRN"], c["codeFreq"], c["status"]) for c in ref_ch]
    path = tmp_path / "fam5.bin"
    raw.tofile(path)
    eng = Engine(s)
    eng.set_param(GC_PARAM_TRACK_EXACT_SUMS, exact)
    with open(path, "rb") as fid:
        tr, _ = tracking(fid, ch, s, engine=eng)
    ref = O.tracking_fam5(raw, ref_ch, so, codes)
    assert tr[2]["status"] == "-" and tr[2]["epochsDone"] == 0
    for i in range(2):
        assert tr[i]["status"] == "T" == ref[i]["status"] and tr[i]["epochsDone"] == nE
        assert np.array_equal(tr[i]["absoluteSample"], ref[i]["absoluteSample"])
        names = ["I_P", "Q_P", "I_E", "I_L", "Q_E", "Q_L"] + (["Pilot_I_P", "Pilot_Q_P"] if pilot else [])
        # 18000 distinct code-phase fractions per epoch: now and then a sample sits within 1e-9 chips of a chip edge
        # and its replica chip hangs on the 10th decimal of remCodePhase (helpers.first_illconditioned_epoch); the
        # 1e-6 comparison runs up to the first such epoch, after it the trajectories may differ by one sample's worth
        ok, _, _ = windowed_iq_compare(f"{signal} pilot={pilot} ch{i} 18 Msps", tr[i], ref[i], names, 18e6, s.dllCorrelatorSpacing, exact=bool(exact))
        assert ("Pilot_I_P" in tr[i]) == bool(pilot)
        assert np.max(np.abs(tr[i]["carrFreq"][:ok] - ref[i]["carrFreq"][:ok])) < 1e-4
        assert np.max(np.abs(tr[i]["codeFreq"][:ok] - ref[i]["codeFreq"][:ok])) < 1e-4
        assert np.max(np.abs(tr[i]["carrFreq"] - ref[i]["carrFreq"])) < 0.05
        if exact:                                          # the loop state follows the float64 reference closely
            assert np.max(np.abs(tr[i]["remCodePhase"] - ref[i]["remCodePhase"])) < 1e-11
            assert np.max(np.abs(tr[i]["codeFreq"] - ref[i]["codeFreq"])) < 1e-8 and np.max(np.abs(tr[i]["carrFreq"] - ref[i]["carrFreq"])) < 1e-8
            d = np.abs(tr[i]["remCarrPhase"] - ref[i]["remCarrPhase"])
            assert np.max(d) < 1e-9, np.max(d)
        if signal == "BDS_B2a":                           # Calc_CNo_PLD.m on the device vs its restatement on the oracle's rows
            want = O.cno_pld_rows(ref[i], so, nE, "BDS_B2a")
            assert tr[i]["DataCNo"].shape == (nE // 40,) and np.all(np.isfinite(tr[i]["DataCNo"]))
            for got_k, want_k in (("DataCNo", "DataCNo"), ("DataPLD", "DataPLD"), ("PilotCNo", "PilotCNo"), ("PilotPLD", "PilotPLD"), ("B2a_CNo", "TotalCNo")):
                assert np.allclose(tr[i][got_k], want[want_k], rtol=1e-7, atol=1e-9), (got_k, tr[i][got_k], want[want_k])
            assert np.all(tr[i]["DataPLD"][-2:] > 0.8) and np.all(tr[i]["DataCNo"][1:] > 35)     # locked by the end of the run
        else:
            nv = ok // 40
            assert np.allclose(tr[i]["CNo"]["VSMValue"][:nv], ref[i]["VSMValue"][:nv], rtol=1e-5)
        assert np.mean(np.abs(tr[i]["I_P"][150:])) > 2 * np.mean(np.abs(tr[i]["Q_P"][150:]))     # pulling in: data in phase,
        if pilot:
            assert np.mean(np.abs(tr[i]["Pilot_Q_P"][150:])) > 2 * np.mean(np.abs(tr[i]["Pilot_I_P"][150:]))   # pilot in quadrature
    eng.close()


# ------------------------------------------------------------- sample formats: int16 and real records
def test_run_shorter_than_one_cno_interval_gives_empty_cno_rows(tmp_path):
    """BDS B2a with msToProcess < CNoInterval: the reference's DataCNo / DataPLD ... are zeros(1, 0) (BDS/B2a/include/tracking.m:79-83);
    the wrapper must return empty rows instead of failing on a zero-interval gc_get_cno_pld."""
    codes, sc, s, so, sv = _fam5_case("BDS_B2a", nsat=1, seed=5, extra=[], nonCoh=3, ms=10, nch=1, pilotTRKflag=1, CNo_VSMinterval=20)
    raw = synth.make_record(sc, 18000 * 14)
    sat = sc.sats[0]
    acq = dict(carrFreq=np.zeros(63), codePhase=np.zeros(63), peakMetric=np.zeros(63))
    acq["carrFreq"][sat.prn - 1] = round((s.IF + sat.doppler) / 25.0) * 25.0
    acq["codePhase"][sat.prn - 1] = int(round((10230 - sat.code_phase) * (18e6 / 10.23e6))) % 18000 + 1
    acq["peakMetric"][sat.prn - 1] = 10.0
    ch = preRun(acq, s)
    path = tmp_path / "b2a.bin"
    raw.tofile(path)
    with open(path, "rb") as fid:
        tr, _ = tracking(fid, ch, s)
    assert tr[0]["status"] == "T" and tr[0]["epochsDone"] == 10 and tr[0]["DataCNo"].shape == (0,) and tr[0]["DataPLD"].shape == (0,)


@pytest.mark.parametrize("fileType,dataType", [(2, "int16"), (1, "schar"), (1, "int16")])
def test_sample_formats_vs_oracle(fileType, dataType, tmp_path):
    """GPS L1CA acquisition + tracking on int16 and on real (fileType 1) records - the dataAdaptCoeff / int16 branches of
    postProcessing.m:66-96 and tracking.m:141-153, 229-240.  The int16 samples exceed the int8 range (x 90), so nothing can
    pass through an 8-bit path unnoticed."""
    fs = 16.368e6
    sc = scene(fs, nsat=3, seed=11, cn0=47)
    nE = 120
    s = init_settings(samplingFreq=fs, fileType=fileType, dataType=dataType, msToProcess=nE, numberOfChannels=4, acqSatelliteList=sorted({x.prn for x in sc.sats} | {30}))
    so = to_oracle_settings(s)
    so.fileType = fileType
    N = O.samples_per_code(so)
    iq8 = synth.make_record(sc, N * (nE + 44))
    scale = 90 if dataType == "int16" else 1
    dt = np.int16 if dataType == "int16" else np.int8
    if fileType == 2:
        raw = (iq8.astype(np.int32) * scale).astype(dt)
    else:
        raw = (iq8[0::2].astype(np.int32) * scale).astype(dt)    # real record: the in-phase samples only
    longSignal = O.read_acq_signal(raw, so)
    assert np.iscomplexobj(longSignal) == (fileType == 2)
    ref = O.acquisition(longSignal, so, workers=os.cpu_count() or 1)
    eng = Engine(s)
    got = acquisition(longSignal, s, engine=eng, verbose=False)
    _check_acq(got, ref, s.acqSatelliteList)
    for sat in sc.sats:
        assert got["carrFreq"][sat.prn - 1] != 0
    assert got["carrFreq"][30 - 1] == 0
    ch = preRun(got, s)
    path = tmp_path / "rec.bin"
    raw.tofile(path)
    with open(path, "rb") as fid:
        tr, _ = tracking(fid, ch, s, engine=eng)
    ref_tr = O.tracking(raw, ch, so)
    for i in range(3):
        assert tr[i]["status"] == "T" == ref_tr[i]["status"] and tr[i]["epochsDone"] == nE
        assert np.array_equal(tr[i]["absoluteSample"], ref_tr[i]["absoluteSample"])
        sc_ = np.hypot(ref_tr[i]["I_P"], ref_tr[i]["Q_P"])
        for name in ("I_P", "Q_P", "I_E", "I_L", "Q_E", "Q_L"):
            assert np.max(np.abs(tr[i][name] - ref_tr[i][name]) / sc_) < IQ_TOL, name
        assert np.max(np.abs(tr[i]["carrFreq"] - ref_tr[i]["carrFreq"])) < 1e-4
        assert np.max(np.abs(tr[i]["remCodePhase"] - ref_tr[i]["remCodePhase"])) < 1e-7
    assert tr[3]["status"] == "-"
    eng.close()


# ------------------------------------------------------------- acquisition variant B: BDS B1I, GPS L2C
def _varb_case(signal, fs, nsat, seed, extra, cn0, **kw):
    sc = synth.default_scene_varb(signal, {}, fs=fs, nsat=nsat, seed=seed)
    for x in sc.sats:
        x.cn0 = cn0
    sv = sorted({x.prn for x in sc.sats} | set(extra))
    codes = sc.codes = oracle_signal_codes(signal, sv)          # real ICD codes (oracle generators); the engine generates its own
    s = init_settings(signal, samplingFreq=fs, acqSatelliteList=sv, **kw)
    so = to_oracle_settings(s)
    so.stepSize, so.acqStep = s.stepSize, s.acqStep
    return codes, sc, s, so, sv


@pytest.mark.parametrize("signal,fs", [("BDS_B1I", 18e6), ("GPS_L2C", 2.046e6), ("GPS_L2C", 8e6), ("BDS_B1I", 4.092e6)])
def test_varb_acquisition_vs_oracle(signal, fs):
    """Variant B: one wipe-off + FFT per sub-bin shift, Doppler bins by circshift of the spectrum, abs(ifft) per row,
    the row with the largest peak kept (two 4 ms blocks for B1I), metric = peak / second peak outside +-1 chip."""
    codes, sc, s, so, sv = _varb_case(signal, fs, nsat=2, seed=3, extra=[30], cn0=48 if signal == "BDS_B1I" else 45,
                                      **({} if signal == "BDS_B1I" else dict(acqSearchBand=9.0 if fs < 8e6 else 4.0)))
    N = O.samples_per_code(so)
    if fs == 8e6:                               # keep the Dopplers inside the reduced +-2 kHz band of this case
        for x in sc.sats:
            x.doppler /= 2.5
    if signal == "BDS_B1I":
        raw = synth.make_record(sc, N * 11)
        longSignal = O.read_acq_signal_varb(raw, so)
        ref = O.acquisition_b1i(longSignal, so, codes, workers=os.cpu_count() or 1)
    else:
        raw = synth.make_record(sc, N * 3)
        longSignal = (raw[0::2] + 1j * raw[1::2]).astype(np.complex128)
        ref = O.acquisition_l2c(longSignal, so, codes, workers=os.cpu_count() or 1)
    eng = Engine(s)
    got = acquisition(longSignal, s, engine=eng, verbose=False)
    # 72000 = 90 x 800 (B1I at 18 Msps) and 320000 = 400 x 800 (L2C at 8 Msps) have fused plans (spectrum shift = row + residue shift)
    assert got["carrFreq"].shape == ref["carrFreq"].shape and eng.stats()["acq_path"] == (1 if fs in (18e6, 8e6) else 0)
    idx = np.array(sv) - 1
    assert np.array_equal(got["carrFreq"], ref["carrFreq"]), "carrier frequency differs"
    assert np.array_equal(got["codePhase"], ref["codePhase"]), "code phase differs"
    assert np.array_equal(got["coarseBin"][idx], ref["coarseBin"][idx]) and np.array_equal(got["coarseCodePhase"][idx], ref["coarseCodePhase"][idx])
    rel = np.abs(got["peakMetric"][idx] - ref["peakMetric"][idx]) / ref["peakMetric"][idx]
    print(f"[parity] {signal} {fs / 1e6:g} Msps variant B peakMetric: worst relative error {rel.max():.2e}")
    assert rel.max() < VARB_METRIC_TOL, rel.max()
    for sat in sc.sats:
        assert got["carrFreq"][sat.prn - 1] != 0
    assert got["carrFreq"][30 - 1] == 0
    eng.close()


@pytest.mark.parametrize("nE,exact", [(1200, 0), (1200, 1)])
def test_b1i_tracking_and_wrappers_vs_oracle(nE, exact, tmp_path):
    """BDS B1I tracking() (1 ms epochs, 2046-chip code from the caller, three-coefficient carrier filter) vs the oracle."""
    codes, sc, s, so, sv = _varb_case("BDS_B1I", 18e6, nsat=2, seed=3, extra=[], cn0=48, msToProcess=nE, numberOfChannels=3,
                                      CNo_VSMinterval=40)
    N = 18000
    raw = synth.make_record(sc, N * (nE + 4))
    ch = []
    for sat in sc.sats:
        start = (2046 - sat.code_phase) * (18e6 / 2.046e6)
        ch.append(dict(PRN=sat.prn, acquiredFreq=round((s.IF + sat.doppler) / 25.0) * 25.0, codePhase=int(round(start)) % N + 1, status="T"))
    ch.append(dict(PRN=0, acquiredFreq=0.0, codePhase=0, status="-"))
    path = tmp_path / "b1i.bin"
    raw.tofile(path)
    eng = Engine(s)
    eng.set_param(GC_PARAM_TRACK_EXACT_SUMS, exact)
    with open(path, "rb") as fid:
        tr, _ = tracking(fid, ch, s, engine=eng)
    ref = O.tracking_b1i(raw, ch, so, codes)
    for i in range(2):
        assert tr[i]["status"] == "T" == ref[i]["status"] and tr[i]["epochsDone"] == nE
        assert np.array_equal(tr[i]["absoluteSample"], ref[i]["absoluteSample"])
        ok, _, _ = windowed_iq_compare(f"BDS_B1I ch{i} 18 Msps", tr[i], ref[i], ("I_P", "Q_P", "I_E", "I_L", "Q_E", "Q_L"), 18e6,
                                       s.dllCorrelatorSpacing, exact=bool(exact))
        assert np.max(np.abs(tr[i]["carrFreq"][:ok] - ref[i]["carrFreq"][:ok])) < 1e-4
        assert np.max(np.abs(tr[i]["carrFreq"] - ref[i]["carrFreq"])) < 0.05
        assert "Pilot_I_P" not in tr[i]
    assert tr[2]["status"] == "-"
    eng.close()


# ------------------------------------------------------------- acquisition variant C: BDS B1C
@pytest.mark.parametrize("fs,pilot", [(4.092e6, 1), (4.092e6, 0), (18e6, 1)])
def test_b1c_acquisition_vs_oracle(fs, pilot):
    """Variant C: one wipe-off + FFT of 20 ms, Doppler bins by circshift, (|data|*sqrt(11) + |pilot|*sqrt(29))/sqrt(40),
    2-D maximum over bins x code phases, 25 Hz fine search over one 10 ms period (FFT length 360000 at 18 Msps)."""
    sc = synth.default_scene_varb("BDS_B1C", {}, fs=fs, nsat=2, seed=3)
    for x in sc.sats:
        x.cn0 = 46
    sv = sorted({x.prn for x in sc.sats} | {30})
    codes = sc.codes = oracle_signal_codes("BDS_B1C", sv)       # the real Weil codes (oracle generators)
    s = init_settings("BDS_B1C", samplingFreq=fs, acqSatelliteList=sv, acqSearchBand=4500.0 if fs < 10e6 else 4000.0, pilotACQflag=pilot)
    so = to_oracle_settings(s)
    so.acqStep, so.pilotACQflag, so.acqCohT = s.acqStep, s.pilotACQflag, s.acqCohT
    N = O.samples_per_code(so)
    raw = synth.make_record(sc, N * 2)
    longSignal = (raw[0::2] + 1j * raw[1::2]).astype(np.complex128)
    ref = O.acquisition_b1c(longSignal, so, codes, workers=os.cpu_count() or 1)
    eng = Engine(s)
    got = acquisition(longSignal, s, engine=eng, verbose=False)
    assert got["carrFreq"].shape == ref["carrFreq"].shape == (max(sv),) and eng.stats()["fft_len"] == 2 * N
    assert eng.stats()["acq_path"] == (1 if fs == 18e6 else 0)          # 360000 = 450 x 800 has a fused plan
    _check_acq(got, ref, sv)
    for sat in sc.sats:
        if pilot:           # the data component alone carries 11/40 of the power and stays under the threshold of 10 here
            assert got["carrFreq"][sat.prn - 1] != 0 and abs(got["carrFreq"][sat.prn - 1] - (s.IF + sat.doppler)) <= 25
    assert got["carrFreq"][30 - 1] == 0
    eng.close()


@pytest.mark.parametrize("fs,nE,exact", [(2.046e6, 60, 0), (8e6, 12, 0), (8e6, 12, 1)])
def test_l2c_tracking_and_wrappers_vs_oracle(fs, nE, exact, tmp_path):
    """GPS L2C tracking() with pilotTRKflag == 0: 20 ms epochs (160000 samples at 8 Msps) in half-chip units on the
    return-to-zero CM table, fseek to codePhase, fractional absoluteSample, halved recorded code quantities."""
    codes, sc, s, so, sv = _varb_case("GPS_L2C", fs, nsat=2, seed=3, extra=[], cn0=45, msToProcess=20 * nE, numberOfChannels=3,
                                      CNo_VSMinterval=10)
    N = O.samples_per_code(so)
    raw = synth.make_record(sc, N * (nE + 2))
    ch = []
    for sat in sc.sats:
        start = (20460 - sat.code_phase) * (fs / 1.023e6)
        ch.append(dict(PRN=sat.prn, acquiredFreq=round((s.IF + sat.doppler) / 12.5) * 12.5, codePhase=int(round(start)) % N, status="T"))
    ch.append(dict(PRN=0, acquiredFreq=0.0, codePhase=0, status="-"))
    path = tmp_path / "l2c.bin"
    raw.tofile(path)
    eng = Engine(s)
    eng.set_param(GC_PARAM_TRACK_EXACT_SUMS, exact)
    with open(path, "rb") as fid:
        tr, _ = tracking(fid, ch, s, engine=eng)
    ref = O.tracking_l2c(raw, ch, so, codes)
    for i in range(2):
        assert tr[i]["status"] == "T" == ref[i]["status"] and tr[i]["epochsDone"] == nE
        ok, _, _ = windowed_iq_compare(f"GPS_L2C ch{i} {fs / 1e6:g} Msps", tr[i], ref[i], ("I_P", "Q_P", "I_E", "I_L", "Q_E", "Q_L"), fs,
                                       s.dllCorrelatorSpacing, rem_scale=2.0, abs_sample=np.floor(ref[i]["absoluteSample"]), exact=bool(exact))
        assert np.max(np.abs(tr[i]["absoluteSample"][:ok] - ref[i]["absoluteSample"][:ok])) < 1e-5      # fractional samples
        assert np.max(np.abs(tr[i]["carrFreq"][:ok] - ref[i]["carrFreq"][:ok])) < 1e-4
        assert np.max(np.abs(tr[i]["codeFreq"][:ok] - ref[i]["codeFreq"][:ok])) < 1e-4
        assert np.max(np.abs(tr[i]["remCodePhase"][:ok] - ref[i]["remCodePhase"][:ok])) < 1e-6
        assert np.allclose(tr[i]["CNo"]["VSMValue"][: ok // 10], ref[i]["VSMValue"][: ok // 10], rtol=1e-5)
    assert tr[2]["status"] == "-"
    eng.close()


def _l2c_pilot_case(fs, nE, cn0=45):
    """GPS L2C scene with the CL pilot time-multiplexed into the CM signal; settings with pilotTRKflag = 1."""
    sc = synth.default_scene_varb("GPS_L2C", {}, fs=fs, nsat=2, seed=3)
    codes = sc.codes = oracle_signal_codes("GPS_L2C", [x.prn for x in sc.sats] + [30], cl=True)   # real CM and CL sequences
    for x in sc.sats:
        x.cn0 = cn0
    sv = sorted({x.prn for x in sc.sats} | {30})
    s = init_settings("GPS_L2C", samplingFreq=fs, acqSatelliteList=sv, acqSearchBand=9.0, pilotTRKflag=1, msToProcess=20 * nE,
                      numberOfChannels=3, CNo_VSMinterval=10)
    so = to_oracle_settings(s)
    so.stepSize, so.acqStep, so.acqCohT = s.stepSize, s.acqStep, s.acqCohT
    return codes, sc, s, so, sv


def test_l2c_cl_phase_search_vs_oracle():
    """GPS L2C acquisition with pilotTRKflag == 1: the 75-way CL code phase search on the acquired PRNs
    (GPS_L2C acquisition.m:100-137) returns the oracle's CLCodePhase, which is the segment the scene put there."""
    fs = 2.046e6
    codes, sc, s, so, sv = _l2c_pilot_case(fs, 4)
    N = O.samples_per_code(so)
    raw = synth.make_record(sc, N * 3)
    longSignal = (raw[0::2] + 1j * raw[1::2]).astype(np.complex128)
    ref = O.acquisition_l2c(longSignal, so, codes, workers=os.cpu_count() or 1)
    eng = Engine(s)
    got = acquisition(longSignal, s, engine=eng, verbose=False)
    assert np.array_equal(got["carrFreq"], ref["carrFreq"]) and np.array_equal(got["codePhase"], ref["codePhase"])
    assert np.array_equal(got["CLCodePhase"], ref["CLCodePhase"]), (got["CLCodePhase"], ref["CLCodePhase"])
    for sat in sc.sats:
        assert got["carrFreq"][sat.prn - 1] != 0
        assert got["CLCodePhase"][sat.prn - 1] == (sat.bit_offset + 1) % 75 + 1       # the CM period that starts inside the record
    assert got["CLCodePhase"][30 - 1] == 0
    ch = preRun(got, s)
    assert ch[0]["CLCodePhase"] == got["CLCodePhase"][ch[0]["PRN"] - 1]
    eng.close()


@pytest.mark.parametrize("fs,nE,exact", [(2.046e6, 80, 0), (8e6, 10, 0), (8e6, 10, 1)])
def test_l2c_cl_pilot_tracking_vs_oracle(fs, nE, exact, tmp_path):
    """GPS L2C tracking() with the CL pilot (pilotTRKflag == 1): the pilot table is the CL segment CLCodePhase points at,
    reloaded every 20 ms epoch and stepping 1..75 (past the wrap here); both discriminator pairs averaged; six Pilot rows."""
    codes, sc, s, so, sv = _l2c_pilot_case(fs, nE)
    N = O.samples_per_code(so)
    raw = synth.make_record(sc, N * (nE + 2))
    ch = []
    for sat in sc.sats:
        start = (20460 - sat.code_phase) * (fs / 1.023e6)
        ch.append(dict(PRN=sat.prn, acquiredFreq=float(round(s.IF + sat.doppler)), codePhase=int(round(start)) % N, status="T",
                       CLCodePhase=(sat.bit_offset + 1) % 75 + 1))
    ch.append(dict(PRN=0, acquiredFreq=0.0, codePhase=0, status="-", CLCodePhase=0))
    path = tmp_path / "l2c.bin"
    raw.tofile(path)
    eng = Engine(s)
    eng.set_param(GC_PARAM_TRACK_EXACT_SUMS, exact)
    with open(path, "rb") as fid:
        tr, _ = tracking(fid, ch, s, engine=eng)
    ref = O.tracking_l2c(raw, ch, so, codes)
    for i in range(2):
        assert tr[i]["status"] == "T" == ref[i]["status"] and tr[i]["epochsDone"] == nE
        ok, _, _ = windowed_iq_compare(f"GPS_L2C CL pilot ch{i} {fs / 1e6:g} Msps", tr[i], ref[i],
                                       ("I_P", "Q_P", "I_E", "I_L", "Q_E", "Q_L", "Pilot_I_P", "Pilot_Q_P", "Pilot_I_E", "Pilot_I_L", "Pilot_Q_E", "Pilot_Q_L"),
                                       fs, s.dllCorrelatorSpacing, rem_scale=2.0, abs_sample=np.floor(ref[i]["absoluteSample"]), exact=bool(exact))
        assert np.max(np.abs(tr[i]["carrFreq"][:ok] - ref[i]["carrFreq"][:ok])) < 1e-4
        assert np.max(np.abs(tr[i]["codeFreq"][:ok] - ref[i]["codeFreq"][:ok])) < 1e-4
        # the pilot carries the CL chips: its prompt correlator is as strong as the data one and in phase with it
        # (8 samples per half chip; at 2 the return-to-zero triangle is too coarse for the loops to settle)
        pw = np.hypot(tr[i]["Pilot_I_P"], tr[i]["Pilot_Q_P"]), np.hypot(tr[i]["I_P"], tr[i]["Q_P"])
        assert np.mean(pw[0]) > 0.5 * np.mean(pw[1])
        if fs == 8e6:
            assert np.mean(np.abs(tr[i]["Pilot_I_P"][nE // 2:])) > 2 * np.mean(np.abs(tr[i]["Pilot_Q_P"][nE // 2:]))
    assert tr[2]["status"] == "-"
    eng.close()


@pytest.mark.parametrize("fs,nE,exact", [(4.092e6, 40, 0), (18e6, 6, 0), (18e6, 6, 1)])
def test_b1c_wb_tracking_vs_oracle(fs, nE, exact, tmp_path):
    """BDS B1C WB_tracking (pilotTRKflag 2): data BOC(1,1), pilot BOC(1,1) and pilot BOC(6,1) tables (int8, 18 sums), the
    BOC(6,1) index ceil(tcode*6)+1, composite pilot correlations, carrier (data + 3 pilot)/4, code error weighted by
    CalcWeighingFactor's factor, six composite Pilot rows."""
    from cu_sdr_collection_b200.tracking import calc_weighing_factor
    sc = synth.default_scene_varb("BDS_B1C", {}, fs=fs, nsat=2, seed=3)
    codes = sc.codes = oracle_signal_codes("BDS_B1C", [x.prn for x in sc.sats], boc61=True)
    for x in sc.sats:
        x.cn0 = 46
    s = init_settings("BDS_B1C", samplingFreq=fs, acqSatelliteList=sorted(x.prn for x in sc.sats), msToProcess=10 * nE,
                      numberOfChannels=3, CNo_VSMinterval=2, pilotTRKflag=2)
    so = to_oracle_settings(s)
    so.FEBW = s.FEBW
    factor = calc_weighing_factor(s)
    assert abs(factor - O.CalcWeighingFactor(so)) < 1e-12 and 0.1 < factor < 0.25
    N = O.samples_per_code(so)
    raw = synth.make_record(sc, N * (nE + 2))
    acq = dict(carrFreq=np.zeros(63), codePhase=np.zeros(63), peakMetric=np.zeros(63))
    for i, sat in enumerate(sc.sats):
        start = (20460 - sat.code_phase) * (fs / 2.046e6)
        acq["carrFreq"][sat.prn - 1] = round((s.IF + sat.doppler) / 25.0) * 25.0
        acq["codePhase"][sat.prn - 1] = int(round(start)) % N + 1
        acq["peakMetric"][sat.prn - 1] = 20.0 - i
    ch = preRun(acq, s)
    path = tmp_path / "b1c.bin"
    raw.tofile(path)
    eng = Engine(s)
    eng.set_param(GC_PARAM_TRACK_EXACT_SUMS, exact)
    with open(path, "rb") as fid:
        tr, _ = tracking(fid, ch, s, engine=eng)
    ref = O.tracking_b1c_wb(raw, ch, so, codes, factor)
    names = ("I_P", "Q_P", "I_E", "I_L", "Q_E", "Q_L", "Pilot_I_P", "Pilot_Q_P", "Pilot_I_E", "Pilot_I_L", "Pilot_Q_E", "Pilot_Q_L")
    for i in range(2):
        assert tr[i]["status"] == "T" == ref[i]["status"] and tr[i]["epochsDone"] == nE
        assert np.array_equal(tr[i]["absoluteSample"], ref[i]["absoluteSample"])
        ok, _, _ = windowed_iq_compare(f"BDS_B1C WB ch{i} {fs / 1e6:g} Msps", tr[i], ref[i], names, fs, s.dllCorrelatorSpacing, sub=12.0,
                                       scale_keys=("Pilot_I_P", "Pilot_Q_P"), exact=bool(exact))
        assert np.max(np.abs(tr[i]["carrFreq"][:ok] - ref[i]["carrFreq"][:ok])) < 1e-4
        assert np.max(np.abs(tr[i]["codeFreq"][:ok] - ref[i]["codeFreq"][:ok])) < 1e-4
        assert np.max(np.abs(tr[i]["dllDiscr"][:ok] - ref[i]["dllDiscr"][:ok])) < 1e-5
        # the composite pilot is in phase (atan(p_Q_P / p_I_P), :353) and carries 3/4 of the power
        if nE >= 40:                                   # (6 epochs are not enough to pull in from a 25 Hz grid)
            assert np.mean(np.abs(tr[i]["Pilot_I_P"][nE // 2:])) > 2 * np.mean(np.abs(tr[i]["Pilot_Q_P"][nE // 2:]))
        assert np.mean(np.hypot(tr[i]["Pilot_I_P"], tr[i]["Pilot_Q_P"])) > 1.2 * np.mean(np.hypot(tr[i]["I_P"], tr[i]["Q_P"]))
        want = O.cno_pld_rows(ref[i], so, nE, "BDS_B1C")     # pilotTRKflag 2: the composite pilot rows as recorded
        assert tr[i]["DataCNo"].shape == (nE // 2,)
        for got_k, want_k in (("DataCNo", "DataCNo"), ("DataPLD", "DataPLD"), ("PilotCNo", "PilotCNo"), ("PilotPLD", "PilotPLD"), ("B1C_CNo", "TotalCNo")):
            assert np.allclose(tr[i][got_k], want[want_k], rtol=1e-6, atol=1e-9, equal_nan=True), (got_k, tr[i][got_k], want[want_k])
    assert tr[2]["status"] == "-"
    eng.close()


@pytest.mark.parametrize("fs,nE,exact", [(4.092e6, 40, 0), (18e6, 8, 0), (18e6, 8, 1)])
def test_b1c_nb_tracking_and_wrappers_vs_oracle(fs, nE, exact, tmp_path):
    """BDS B1C NB_tracking (pilotTRKflag 1): 10 ms epochs (180000 samples at 18 Msps, one sample window in shared memory
    next to the two BOC(1,1) tables), carrier-aided code NCO, quadrature pilot atan(-I/Q), 11/40 : 29/40 weights,
    (1 - spacing)-scaled code discriminators, Pilot rows, DataCNo / PLD block on the host."""
    sc = synth.default_scene_varb("BDS_B1C", {}, fs=fs, nsat=2, seed=3)
    codes = sc.codes = oracle_signal_codes("BDS_B1C", [x.prn for x in sc.sats])
    for x in sc.sats:
        x.cn0 = 46
    s = init_settings("BDS_B1C", samplingFreq=fs, acqSatelliteList=sorted(x.prn for x in sc.sats), msToProcess=10 * nE,
                      numberOfChannels=3, CNo_VSMinterval=4)
    so = to_oracle_settings(s)
    N = O.samples_per_code(so)
    raw = synth.make_record(sc, N * (nE + 2))
    acq = dict(carrFreq=np.zeros(63), codePhase=np.zeros(63), peakMetric=np.zeros(63))
    for i, sat in enumerate(sc.sats):
        start = (20460 - sat.code_phase) * (fs / 2.046e6)
        acq["carrFreq"][sat.prn - 1] = round((s.IF + sat.doppler) / 25.0) * 25.0
        acq["codePhase"][sat.prn - 1] = int(round(start)) % N + 1
        acq["peakMetric"][sat.prn - 1] = 20.0 - i
    ch = preRun(acq, s)
    assert ch[0]["codeFreq"] == s.codeFreqBasis + (ch[0]["acquiredFreq"] - s.IF) / s.carrFreqBasis * s.codeFreqBasis
    path = tmp_path / "b1c.bin"
    raw.tofile(path)
    eng = Engine(s)
    eng.set_param(GC_PARAM_TRACK_EXACT_SUMS, exact)
    with open(path, "rb") as fid:
        tr, _ = tracking(fid, ch, s, engine=eng)
    ref = O.tracking_b1c_nb(raw, ch, so, codes)
    for i in range(2):
        assert tr[i]["status"] == "T" == ref[i]["status"] and tr[i]["epochsDone"] == nE
        assert np.array_equal(tr[i]["absoluteSample"], ref[i]["absoluteSample"])
        ok, _, _ = windowed_iq_compare(f"BDS_B1C NB ch{i} {fs / 1e6:g} Msps", tr[i], ref[i], ("I_P", "Q_P", "I_E", "I_L", "Q_E", "Q_L", "Pilot_I_P", "Pilot_Q_P"),
                                       fs, s.dllCorrelatorSpacing, sub=2.0, scale_keys=("Pilot_I_P", "Pilot_Q_P"), exact=bool(exact))
        assert np.max(np.abs(tr[i]["carrFreq"][:ok] - ref[i]["carrFreq"][:ok])) < 1e-4
        assert np.max(np.abs(tr[i]["codeFreq"][:ok] - ref[i]["codeFreq"][:ok])) < 1e-4
        want = O.cno_pld_rows(ref[i], so, nE, "BDS_B1C")     # pilotTRKflag 1: the pilot rows swap roles
        assert tr[i]["DataCNo"].shape == (nE // 4,) and "B1C_CNo" in tr[i] and np.all(np.isfinite(tr[i]["PilotCNo"]))
        for got_k, want_k in (("DataCNo", "DataCNo"), ("DataPLD", "DataPLD"), ("PilotCNo", "PilotCNo"), ("PilotPLD", "PilotPLD"), ("B1C_CNo", "TotalCNo")):
            assert np.allclose(tr[i][got_k], want[want_k], rtol=1e-6, atol=1e-9, equal_nan=True), (got_k, tr[i][got_k], want[want_k])
    assert tr[2]["status"] == "-"
    eng.close()


# ------------------------------------------------------------- navigation-bit front end
def test_nav_front_end_vs_oracle():
    """gc_nav_sync against the NAVdecoding.m:69-170 restatement: clean, noisy / inverted, late (bits do not fit), weak
    (bit errors break the parity of the first candidates) and random channels - subFrameStart and all 1501 bits exact."""
    from helpers import nav_message_bits, nav_prompt_row
    from cu_sdr_collection_b200.navsync import nav_sync
    n = 60000
    s = init_settings(samplingFreq=16.368e6, msToProcess=n, numberOfChannels=6)
    rows = [nav_prompt_row(nav_message_bits(9, seed=5), 1234, n, 2000.0, 300.0, seed=1),
            nav_prompt_row(nav_message_bits(9, seed=6), 4321, n, 1500.0, 400.0, seed=2, polarity=-1),
            nav_prompt_row(nav_message_bits(9, seed=7), 40000, n, 2000.0, 100.0, seed=3),
            nav_prompt_row(nav_message_bits(9, seed=8), 777, n, 1000.0, 4000.0, seed=4),
            2000.0 * (1 - 2 * np.random.default_rng(3).integers(0, 2, size=3000)).repeat(20).astype(np.float64),
            np.zeros(n)]
    tr = [dict(I_P=r) for r in rows]
    eng = Engine(s)
    sfs, bits = nav_sync(tr, s, eng)
    for ch, r in enumerate(rows):
        want_sfs, want_bits = O.nav_sync(r, n)
        assert sfs[ch] == want_sfs, (ch, sfs[ch], want_sfs)
        assert (bits[ch] is None) == (want_bits is None)
        if want_bits is not None:
            assert np.array_equal(bits[ch], want_bits), ch
    assert sfs[0] == 1234 and sfs[1] == 4321 and sfs[2] == 40000 and bits[2] is None and sfs[4] == 0 and sfs[5] == 0
    eng.close()


def test_acquire_track_one_call_matches_two_calls():
    """gc_acquire_track (acquisition -> preRun -> tracking inside the library) returns exactly what
    acquisition(), preRun() and tracking() return one after the other; C/N0 (now computed on the device) matches the oracle."""
    fs = 16.368e6
    sc = scene(fs, nsat=3, seed=11, cn0=47)
    nE = 120
    s = init_settings(samplingFreq=fs, msToProcess=nE, numberOfChannels=5, acqSatelliteList=sorted({x.prn for x in sc.sats} | {30}), CNo_VSMinterval=40)
    N = 16368
    raw = synth.make_record(sc, N * (nE + 44))
    eng = Engine(s)
    eng.set_record(raw)
    acq2 = eng.acquire()
    ch2 = preRun(acq2, s)
    tr2, _ = tracking(None, ch2, s, engine=eng)
    acq1, ch1, out, vv, vi, done = eng.acquire_track(s.numberOfChannels, nE)
    for k in ("carrFreq", "codePhase", "peakMetric"):
        assert np.array_equal(acq1[k], acq2[k]), k
    assert [c["PRN"] for c in ch1] == [c["PRN"] for c in ch2] and [c["codePhase"] for c in ch1] == [c["codePhase"] for c in ch2]
    assert [c["PRN"] for c in ch1][3:] == [0, 0]
    from cu_sdr_collection_b200.tracking import TRACK_FIELDS
    for i in range(5):
        for j, f in enumerate(TRACK_FIELDS):
            assert np.array_equal(out[i, j], tr2[i][f]), (i, f)
        assert np.array_equal(vv[i], tr2[i]["CNo"]["VSMValue"]) and np.array_equal(vi[i], tr2[i]["CNo"]["VSMIndex"])
    ref = O.tracking(raw, ch2, to_oracle_settings(s))
    for i in range(3):
        assert np.allclose(vv[i], ref[i]["VSMValue"], rtol=1e-5) and np.array_equal(vi[i], ref[i]["VSMIndex"]) and vi[i][-1] == 120
    assert not np.any(vv[3:]) and not np.any(vi[3:])
    eng.close()


def test_packed_2bit_record_equals_unpacked_schar_record(tmp_path):
    """fileType GC_FILE_PACKED2: the 2-bit packed I/Q record (the input of unpack_cplx.m) decoded on the fly gives exactly what
    the engine gives on the 'schar' file unpack_cplx.m would write - acquisition results and every tracking row - and that
    matches the oracle on the unpacked record."""
    fs = 16.368e6
    sc = scene(fs, nsat=3, seed=11, cn0=50)
    nE = 80
    sv = sorted({x.prn for x in sc.sats} | {30})
    N = 16368
    x = synth.quantize2(synth.make_record(sc, N * (nE + 44)))
    packed = synth.pack_cplx2(x)
    unpacked = O.unpack_cplx(packed)
    res = {}
    for ft, rec in ((3, packed), (2, unpacked)):
        s = init_settings(samplingFreq=fs, fileType=ft, msToProcess=nE, numberOfChannels=4, acqSatelliteList=sv)
        eng = Engine(s)
        eng.set_record(rec)
        acq = eng.acquire()
        ch = preRun(acq, s)
        path = tmp_path / ("rec%d.bin" % ft)
        rec.tofile(path)
        with open(path, "rb") as fid:
            tr, _ = tracking(fid, ch, s, engine=eng)
        res[ft] = (acq, ch, tr)
        eng.close()
    for k in ("carrFreq", "codePhase", "coarseBin"):
        assert np.array_equal(res[3][0][k], res[2][0][k]), k
    assert np.allclose(res[3][0]["peakMetric"], res[2][0]["peakMetric"], rtol=1e-12)
    from cu_sdr_collection_b200.tracking import TRACK_FIELDS
    for i in range(4):
        for f in TRACK_FIELDS:
            a, b = res[3][2][i][f], res[2][2][i][f]
            assert np.allclose(a, b, rtol=1e-9, atol=1e-9, equal_nan=True), (i, f)
    so = to_oracle_settings(init_settings(samplingFreq=fs, msToProcess=nE, numberOfChannels=4, acqSatelliteList=sv))
    ref = O.acquisition(O.read_acq_signal(unpacked, so), so, workers=os.cpu_count() or 1)
    _check_acq(res[3][0], ref, sv)
    for sat in sc.sats:
        assert res[3][0]["carrFreq"][sat.prn - 1] != 0
    ref_tr = O.tracking(unpacked, res[3][1], so)
    for i in range(3):
        sc_ = np.hypot(ref_tr[i]["I_P"], ref_tr[i]["Q_P"])
        for name in ("I_P", "Q_P", "I_E", "I_L", "Q_E", "Q_L"):
            assert np.max(np.abs(res[3][2][i][name] - ref_tr[i][name]) / sc_) < IQ_TOL, name


@pytest.mark.parametrize("signal", ["BDS_B3I", "GLO_GL1", "GPS_L5C", "BDS_B2a"])
def test_acquire_track_one_call_other_signals(signal):
    """gc_acquire_track for the carrier-aided signals (channel.codeFreq from cfg.carr_freq_basis, GPS_L5C/include/preRun.m:69-71)
    and for GLONASS (channels carry K): the one call returns exactly what acquisition(), preRun() and tracking() return in turn."""
    from cu_sdr_collection_b200.engine import GC_SV_NONE
    nE, codes = 60, None
    if signal == "BDS_B3I":
        sc = synth.default_scene_b3i(fs=18e6, nsat=3, seed=9)
        for x, p in zip(sc.sats, (3, 20, 41)):
            x.prn, x.cn0 = p, 48
        sv, N, per = [3, 20, 41, 7], 18000, 40
        s = init_settings(signal, samplingFreq=18e6, acqSatelliteList=sv, acqNonCohTime=5, msToProcess=nE, numberOfChannels=4)
    elif signal == "GLO_GL1":
        sc = synth.default_scene_glo(fs=12e6, nsat=3, seed=23)
        for x in sc.sats:
            x.cn0 = 47
        sv, N, per = sorted({x.prn for x in sc.sats} | {5, -6}), 12000, 50
        s = init_settings(signal, samplingFreq=12e6, acqSatelliteList=sv, acqNonCohTime=6, msToProcess=nE, numberOfChannels=4)
    else:
        sc = synth.default_scene_fam5(signal, {}, fs=18e6, nsat=2, seed=5)
        for x in sc.sats:
            x.cn0 = 50
        sv, N, per = sorted({x.prn for x in sc.sats} | {25}), 18000, 44
        sc.codes = oracle_signal_codes(signal, sv)
        codes = None                                          # the engine generates its own
        s = init_settings(signal, acqSatelliteList=sv, acqNonCohTime=3, msToProcess=nE, numberOfChannels=3, pilotTRKflag=1, CNo_VSMinterval=20)
    raw = synth.make_record(sc, N * (nE + per))
    eng = Engine(s) if codes is not None else Engine(s)
    eng.set_record(raw)
    acq2 = eng.acquire(sv)
    ch2 = preRun(acq2, s)
    tr2, _ = tracking(None, ch2, s, engine=eng)
    acq1, ch1, out, vv, vi, done = eng.acquire_track(s.numberOfChannels, nE, sv_list=sv)
    for k in ("carrFreq", "codePhase", "peakMetric"):
        assert np.array_equal(acq1[k], acq2[k]), k
    key = "K" if signal == "GLO_GL1" else "PRN"
    assert [(c[key], c["status"], c["codePhase"]) for c in ch1] == [(c[key], c["status"], c["codePhase"]) for c in ch2]
    assert sum(c["status"] == "T" for c in ch1) == len(sc.sats)
    from cu_sdr_collection_b200.tracking import TRACK_FIELDS
    for i in range(s.numberOfChannels):
        for j, f in enumerate(TRACK_FIELDS):
            assert np.array_equal(out[i, j], tr2[i][f]), (i, f)
        assert done[i] == tr2[i]["epochsDone"]
    if signal == "BDS_B2a":                                  # the device-side Calc_CNo_PLD block of the same call
        pld = eng.cno_pld(s.numberOfChannels, nE)
        assert np.array_equal(pld[0, 0], tr2[0]["DataCNo"]) and np.array_equal(pld[0, 4], tr2[0]["B2a_CNo"]) and not pld[2].any()
    eng.close()
