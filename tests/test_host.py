"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the
MEX gateway type-checks against the header, the MATLAB-mirroring host logic, and the
world_size-2 sharding/gather path over gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import cu_sdr_collection_b200 as pkg
from cu_sdr_collection_b200 import engine, shard
from cu_sdr_collection_b200.settings import Settings, init_settings, samples_per_code
from helpers import ROOT


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gnsscorr.h")).read()
    declared = set(re.findall(r"\b(gc_[a-z_]+)\s*\(", hdr))
    assert {"gc_create", "gc_destroy", "gc_acquire", "gc_acquire_host", "gc_track", "gc_track_file",
            "gc_set_record_host", "gc_set_record_device", "gc_get_stats", "gc_last_error"} <= declared
    lib = engine.load_lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert set(engine.EXPORTS) == declared
    assert lib.gc_abi_version() == 5 and lib.gc_build_arch() == b"sm_100a"
    assert lib.gc_acq_result_len(0) == 32 and lib.gc_acq_result_len(1) == 21 and lib.gc_acq_result_len(3) == 50


def test_config_struct_layout_matches_header():
    hdr = open(os.path.join(ROOT, "include", "gnsscorr.h")).read()
    body = re.search(r"typedef struct gc_config \{(.*?)\} gc_config;", hdr, re.S).group(1)
    names = re.findall(r"^\s*(?:int32_t|int64_t|double)\s+([A-Za-z_0-9]+);", body, re.M)
    assert names == [f for f, _ in engine.gc_config._fields_]
    body = re.search(r"typedef struct gc_stats \{(.*?)\} gc_stats;", hdr, re.S).group(1)
    names = re.findall(r"^\s*(?:int32_t|float)\s+([A-Za-z_]+);", body, re.M)
    assert names == [f for f, _ in engine.gc_stats._fields_]


def test_generated_codelets_match_direct_dft(tmp_path):
    """The packed-fp32x2 register DFT codelets (tools/gen_codelets.py -> csrc/fft_codelets.cuh) compiled for
    the host with stub intrinsics: every length, forward and inverse, against a float64 direct DFT."""
    exe = str(tmp_path / "codelet_host_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "cu-sdr-collection_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "host_src", "codelet_host_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert out.stdout.count("max err") == 15
    gen = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_codelets.py")], capture_output=True, text=True, check=True)
    assert gen.stdout == open(os.path.join(ROOT, "cu-sdr-collection_b200", "csrc", "fft_codelets.cuh")).read(), \
        "fft_codelets.cuh is stale: regenerate with tools/gen_codelets.py"


def test_codelet_operation_counts():
    """The packed-operation counts DESIGN.md quotes for the row pass are properties of the generated header: 194 for the
    32-point codelet (three-operation butterflies), Rader's 31-point form below the direct symmetric one (510), 186 for the
    30-point prime-factor codelet inside it, and no length above the generator's model."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_codelets", os.path.join(ROOT, "tools", "gen_codelets.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    src = open(os.path.join(ROOT, "cu-sdr-collection_b200", "csrc", "fft_codelets.cuh")).read()

    def emitted(n):
        body = re.search(r"void dft%d_inv\(.*?\n\{\n(.*?)\n\}\n" % n, src, re.S).group(1)
        ops = 0
        for line in body.split("\n"):
            m = re.match(r"\s*const float2 v\d+ = (f2\w+)\(", line)
            if m:
                ops += 2 if m.group(1) == "f2cmulc" else 1
        return ops

    assert emitted(32) == 194 and emitted(30) == 186
    assert emitted(31) == 417 < 432 < gen.packed_ops_direct(31) == 510      # Rader, products fused into the first butterflies
    assert emitted(25) == 184                                                 # 5 x 5 with the twiddles fused into the pair sums
    for n in (4, 8, 9, 10, 16, 18, 20, 25, 30, 31, 32, 33, 40, 45, 50):      # the model counts every literal product as two
        assert emitted(n) <= gen.packed_ops(n), n


def test_fine_search_moment_expansion_bound():
    """fine_sum_moments_kernel (csrc/acq_common.cu) replaces the 21 per-bin carrier wipe-offs of acquisition.m:230-238 by one
    wipe-off at the centre bin and four moments per run of 256 samples.  The same arithmetic in NumPy (float32 moments, the
    kernel's operation order) against the direct float64 sums, at the reference's defaults (16.368 Msps, fine bins 25 Hz apart
    over +-250 Hz): the difference stays below 1e-6 of the sum's scale - the fine-bin arg-max cannot move."""
    rng = np.random.default_rng(7)
    fs, N, run = 16.368e6, 16368, 256
    c = 3                                                        # fourth code period: finePhasePoints index c*N + n
    z = (rng.integers(-128, 128, N) + 1j * rng.integers(-128, 128, N))
    f_c = 4.092e6 + 1234.0
    z = z * np.exp(2j * np.pi * (f_c + 60.0) / fs * (c * N + np.arange(N)))     # a tone near the centre bin, as after x.*code
    g = c * N + np.arange(N, dtype=np.float64)
    worst = 0.0
    scale = np.abs(np.sum(z * np.exp(-2j * np.pi * (f_c + 60.0) / fs * g)))
    zc = (z * np.exp(-2j * np.pi * f_c / fs * g)).astype(np.complex64)            # the centre-bin wipe-off, once
    for j in range(-10, 11):
        d = 25.0 * j / fs                                        # turns per sample between bin j and the centre bin
        direct = np.sum(z * np.exp(-2j * np.pi * (f_c / fs + d) * g))
        t = np.float32(2 * np.pi * d)
        acc = 0j
        for n0 in range(0, N, run):
            seg = zc[n0:n0 + run]
            r = (np.arange(len(seg)) - run // 2).astype(np.float32)
            m0 = seg.sum(dtype=np.complex64); m1 = (seg * r).sum(dtype=np.complex64)
            m2 = (seg * r * r).sum(dtype=np.complex64); m3 = (seg * r * r * r).sum(dtype=np.complex64)
            poly = m0 - 1j * t * m1 - np.float32(0.5) * t * t * m2 + 1j * (t * t * t / np.float32(6)) * m3
            acc += complex(poly) * np.exp(-2j * np.pi * d * (c * N + n0 + run // 2))
        worst = max(worst, abs(acc - direct) / scale)
    assert worst < 1e-6, worst


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.GnssCorrError, match="no CPU fallback"):
        pkg.Engine(init_settings(samplingFreq=16.368e6))


def test_product_never_imports_the_oracle():
    pkgdir = os.path.join(ROOT, "cu-sdr-collection_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "np_oracle" not in src and "gnss_oracle" not in src and "oracle/" not in src, f


def test_mex_gateway_typechecks_against_header():
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "matlab", "stub"),
                           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "matlab", "gnsscorr_mex.c")])


def test_settings_defaults_match_reference_initsettings():
    s = Settings()
    assert (s.msToProcess, s.numberOfChannels, s.samplingFreq, s.IF) == (60000, 12, 18e6, 20e3)
    assert (s.acqSearchBand, s.acqSearchStep, s.acqNonCohTime, s.acqThreshold) == (7000, 500, 20, 3.5)
    assert s.acqSatelliteList == list(range(1, 33))
    assert samples_per_code(s) == 18000
    assert samples_per_code(init_settings(samplingFreq=16.368e6)) == 16368
    with pytest.raises(AttributeError):
        init_settings(noSuchField=1)
    with pytest.raises(engine.GnssCorrError):
        engine.config_from_settings(init_settings(resamplingflag=1))


def test_prerun_orders_by_peak_metric():
    acq = dict(peakMetric=np.zeros(32), carrFreq=np.zeros(32), codePhase=np.zeros(32))
    for prn, pm in ((3, 5.0), (9, 9.0), (20, 7.0), (31, 2.0)):
        acq["peakMetric"][prn - 1] = pm
    for prn in (3, 9, 20):
        acq["carrFreq"][prn - 1] = 20e3 + prn
        acq["codePhase"][prn - 1] = 100 * prn
    ch = pkg.preRun(acq, init_settings(numberOfChannels=4))
    assert [c["PRN"] for c in ch] == [9, 20, 3, 0]
    assert ch[0]["acquiredFreq"] == 20009 and ch[0]["codePhase"] == 900 and ch[3]["status"] == "-"
    ch = pkg.preRun(acq, init_settings(numberOfChannels=2))
    assert [c["PRN"] for c in ch] == [9, 20]


def test_acquisition_sample_conversion():
    """longSignal -> the file's own samples for the four fileType / dataType combinations; non-integer or out-of-range
    input is refused (the accelerated path needs the raw samples)."""
    from cu_sdr_collection_b200.acquisition import _to_file_samples
    s8 = init_settings()
    x = (np.arange(8) - 4) + 1j * (np.arange(8) - 3)
    iq = _to_file_samples(x.astype(np.complex128), s8)
    assert iq.dtype == np.int8 and list(iq[:4]) == [-4, -3, -3, -2]
    with pytest.raises(engine.GnssCorrError):
        _to_file_samples(x + 0.5, s8)
    with pytest.raises(engine.GnssCorrError):
        _to_file_samples(np.arange(8, dtype=np.float64), s8)
    with pytest.raises(engine.GnssCorrError):
        _to_file_samples(300.0 * x, s8)
    s16 = init_settings(dataType="int16")
    iq = _to_file_samples(300.0 * x, s16)
    assert iq.dtype == np.int16 and list(iq[:4]) == [-1200, -900, -900, -600]
    sr = init_settings(fileType=1)
    r = _to_file_samples(np.arange(8, dtype=np.float64) - 4, sr)
    assert r.dtype == np.int8 and r.size == 8 and r[0] == -4
    with pytest.raises(engine.GnssCorrError):
        _to_file_samples(x, sr)
    cfg = engine.config_from_settings(init_settings(fileType=1, dataType="int16"))
    assert cfg.file_type == 1 and cfg.sample_bytes == 2


def test_shard_units_partition():
    sv = list(range(1, 33))
    parts = [shard.shard_units(sv, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == sv and all(len(p) == 8 for p in parts)
    assert shard.shard_units(sv, 0, 1) == sv
    assert [len(shard.shard_units(list(range(14)), r, 4)) for r in range(4)] == [4, 4, 3, 3]   # GLONASS K list


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch.distributed as dist
from cu_sdr_collection_b200 import shard
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
sv = shard.shard_units(list(range(1, 33)), rank, 2)
local = dict(peakMetric=np.zeros(32), codePhase=np.zeros(32), carrFreq=np.zeros(32), coarseBin=np.zeros(32, dtype=np.int32))
for p in sv:                                    # stand-in for the per-rank GPU search
    local["peakMetric"][p - 1] = 1.0 + p
    local["codePhase"][p - 1] = 10 * p
    local["carrFreq"][p - 1] = 20e3 + p if p % 3 == 0 else 0.0
    local["coarseBin"][p - 1] = p % 29 + 1
m = shard.gather_acq_results(local, sv)
# the device-buffer path of a one-process-per-GPU run (Engine.acquire_device -> all_gather_into_tensor -> sum), on CPU tensors here
import torch
buf = torch.from_numpy(np.concatenate([local["peakMetric"], local["codePhase"], local["carrFreq"], local["coarseBin"].astype(np.float64)]))
md = shard.merge_device_results(shard.all_gather_device(buf), 32)
for k in ("peakMetric", "codePhase", "carrFreq", "coarseBin"):
    assert np.array_equal(md[k], m[k]), k
# GLONASS: the unit list holds frequency numbers K = -7..6 stored at index K + 7 (MATLAB's K + 8)
ks = shard.shard_units(list(range(-7, 7)), rank, 2)
gl = dict(peakMetric=np.zeros(21), codePhase=np.zeros(21), carrFreq=np.zeros(21), coarseBin=np.zeros(21, dtype=np.int32))
for k in ks:
    gl["peakMetric"][k + 7] = 100.0 + k
    gl["carrFreq"][k + 7] = 1e6 - 562.5e3 * k
g = shard.gather_acq_results(gl, ks, glonass=True)
assert np.array_equal(g["peakMetric"][:14], 100.0 + np.arange(-7, 7)) and not g["peakMetric"][14:].any(), g["peakMetric"]
assert np.array_equal(g["carrFreq"][:14], 1e6 - 562.5e3 * np.arange(-7, 7))
assert np.array_equal(m["peakMetric"], 1.0 + np.arange(1, 33)), m["peakMetric"]
assert np.array_equal(m["codePhase"], 10.0 * np.arange(1, 33))
assert np.array_equal(m["carrFreq"] != 0, np.arange(1, 33) % 3 == 0)
assert np.array_equal(m["coarseBin"], np.arange(1, 33) % 29 + 1)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_plan_pairs_balances_by_cost_and_covers_every_pair():
    """The (signal, SV) pairs of the all-constellation search (BASELINE configs[4]) dealt over 8 ranks by measured cost."""
    svs = {"GPS_L1CA": range(1, 33), "GLO_GL1": range(-7, 7), "GLO_GL2": range(-7, 7), "BDS_B3I": range(1, 64), "GAL_E1C": range(1, 37),
           "GPS_L5C": range(1, 33), "GAL_E5a": range(1, 37), "GAL_E5b": range(1, 37), "BDS_B2a": range(1, 30), "BDS_B1I": range(1, 54),
           "GPS_L2C": range(1, 33), "BDS_B1C": range(1, 63)}
    pairs = [(sig, sv) for sig, r in svs.items() for sv in r]
    assert len(pairs) == 439
    for world in (1, 2, 4, 8):
        plan, load = shard.plan_pairs(pairs, world)
        got = sorted((sig, sv) for r in plan for sig, l in r.items() for sv in l)
        assert got == sorted(pairs)                                   # every pair on exactly one rank
        assert max(load) <= 1.08 * (sum(load) / world) + 1e-9, (world, load)
        assert plan == shard.plan_pairs(pairs, world)[0]              # deterministic: every rank computes the same plan
    plan, load = shard.plan_pairs(pairs, 8)
    total = sum(shard.COST_MS_PER_SV[s] * len(r) for s, r in svs.items())
    assert max(load) < total / 8 + 4.0                                # the 40+ ms signals (L2C, B1C) are spread over all ranks
    assert all("GPS_L2C" in r and "BDS_B1C" in r for r in plan)
    assert shard.shard_channels(592, 7, 8) == list(range(518, 592)) and shard.shard_channels(12, 6, 8) == [] and shard.shard_channels(12, 1, 8) == [2, 3]
    assert shard.result_index(-7, True) == 0 and shard.result_index(13, True) == 20 and shard.result_index(32) == 31


def test_two_rank_gloo_gather(tmp_path):
    """world_size 2 over gloo: PRNs sharded round-robin, one all-gather rebuilds acqResults (host arrays, device-buffer layout,
    GLONASS frequency numbers)."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, port=29000 + os.getpid() % 2000))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)
