"""GPU tests of the multi-GPU paths: results left on the device for the all-gather (gc_acquire_device), the
in-library fan-out (gc_multi_*) and the one-process-per-GPU NCCL run - each must give bit-identical acqResults / trackResults to
one GPU.  The two-GPU cases skip on a one-GPU box; the one-GPU cases of the same code paths always run."""
import os
import subprocess
import sys

import numpy as np
import pytest

import np_oracle as O
from cu_sdr_collection_b200 import Engine, MultiEngine, init_settings, preRun, synth
from cu_sdr_collection_b200.engine import GC_SV_NONE
from helpers import ROOT, scene, to_oracle_settings

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _l1ca_case(nonCoh=4, nE=150, nsat=5):
    fs = 16.368e6
    sc = scene(fs, nsat=nsat, seed=20260101, cn0=46)
    sv = sorted({x.prn for x in sc.sats} | {1, 2, 3, 30, 31})
    s = init_settings(samplingFreq=fs, acqNonCohTime=nonCoh, acqSatelliteList=sv, msToProcess=nE, numberOfChannels=nsat + 2)
    raw = synth.make_record(sc, 16368 * (nE + 46))
    return sc, s, sv, raw


def _device_results(eng, sv):
    import torch
    n = eng.lib.gc_acq_result_len(__import__("cu_sdr_collection_b200").engine.signal_id(eng.settings))
    buf = torch.full((4 * n,), -1.0, dtype=torch.float64, device="cuda")
    eng.acquire_device(sv, buf)
    a = buf.cpu().numpy()
    return dict(peakMetric=a[:n], codePhase=a[n:2 * n], carrFreq=a[2 * n:3 * n], coarseBin=a[3 * n:].astype(np.int32))


def test_acquire_device_equals_host_results_variant_a():
    """gc_acquire_device (acqResults assembled by pack_results_kernel) == gc_acquire, bit for bit: GPS L1CA and GLONASS."""
    sc, s, sv, raw = _l1ca_case()
    eng = Engine(s)
    eng.set_record(raw)
    host = eng.acquire(sv)
    dev = _device_results(eng, sv)
    for k in ("peakMetric", "codePhase", "carrFreq", "coarseBin"):
        assert np.array_equal(dev[k], host[k]), k
    assert np.count_nonzero(host["carrFreq"]) >= 5
    part = _device_results(eng, sv[:3])                      # entries of SVs not searched are zero
    idx = np.array(sv[:3]) - 1
    assert np.array_equal(part["peakMetric"][idx], host["peakMetric"][idx]) and np.count_nonzero(part["peakMetric"]) == 3
    eng.close()
    gsc = synth.default_scene_glo(fs=12e6, nsat=3, seed=17)
    for x in gsc.sats:
        x.cn0 = 47
    ks = list(range(-7, 7))
    gs = init_settings("GLO_GL1", samplingFreq=12e6, acqNonCohTime=4, acqSatelliteList=ks)
    graw = synth.make_record(gsc, 12000 * 44)
    geng = Engine(gs)
    geng.set_record(graw)
    host = geng.acquire(ks)
    dev = _device_results(geng, ks)
    for k in ("peakMetric", "codePhase", "carrFreq", "coarseBin"):
        assert np.array_equal(dev[k], host[k]), k
    assert np.count_nonzero(host["carrFreq"]) >= 3
    geng.close()


def test_graph_replay_and_async_results_equal_the_first_call():
    """The variant A enqueue runs directly on the first call, is captured into a CUDA graph on the second identical call and
    replayed from the third; gc_acquire_device_async returns before the search has finished (results in stream order).  Every
    form must return the first call's numbers bit for bit, also after a different SV list was searched in between, and the stats
    must keep reporting kernel times."""
    import torch
    sc, s, sv, raw = _l1ca_case()
    eng = Engine(s)
    eng.set_record(raw)
    first = eng.acquire(sv)
    for _ in range(4):                                        # capture, then replays
        again = eng.acquire(sv)
        for k in ("peakMetric", "codePhase", "carrFreq", "coarseBin", "coarseCodePhase"):
            assert np.array_equal(again[k], first[k]), k
        st = eng.stats()
        assert st["acq_total_ms"] > 0 and st["corr_rows_ms"] > 0 and st["acq_launches"] >= 9
    other = eng.acquire(sv[:4])                               # another list resets the cached graph ...
    idx = np.array(sv[:4]) - 1
    assert np.array_equal(other["peakMetric"][idx], first["peakMetric"][idx]) and np.count_nonzero(other["peakMetric"]) == 4
    n = 32
    buf = torch.full((4 * n,), -1.0, dtype=torch.float64, device="cuda")
    ext = torch.cuda.ExternalStream(eng.stream_ptr)
    for _ in range(4):                                        # ... and the asynchronous form goes direct -> captured -> replayed too
        buf.fill_(-1.0)
        torch.cuda.synchronize()
        eng.acquire_device_async(sv, buf)
        with torch.cuda.stream(ext):
            a = buf.clone()                                   # ordered behind the search on the engine's stream
        ext.synchronize()
        a = a.cpu().numpy()
        assert np.array_equal(a[:n], first["peakMetric"]) and np.array_equal(a[n:2 * n], first["codePhase"])
        assert np.array_equal(a[2 * n:3 * n], first["carrFreq"]) and np.array_equal(a[3 * n:].astype(np.int32), first["coarseBin"])
        assert eng.stats()["acq_total_ms"] > 0
    host = eng.acquire(sv, host_iq=raw[: 2 * 16368 * 46])     # gc_acquire_host: copy and search without a synchronisation between
    host2 = eng.acquire(sv, host_iq=raw[: 2 * 16368 * 46])
    for k in ("peakMetric", "codePhase", "carrFreq"):
        assert np.array_equal(host[k], first[k]) and np.array_equal(host2[k], first[k]), k
    eng.close()


def test_acquire_device_equals_host_results_other_variants():
    """No fine stage (GAL E5b) and the variants that finish on the host (BDS B1I)."""
    from cu_sdr_collection_b200.codes import icd_codes
    codes = icd_codes("GAL_E5b")
    sc = synth.default_scene_fam5("GAL_E5b", codes, fs=18e6, nsat=2, seed=5)
    for x in sc.sats:
        x.cn0 = 50
    sv = sorted({x.prn for x in sc.sats} | {25})
    s = init_settings("GAL_E5b", acqSatelliteList=sv, acqNonCohTime=3, acqSearchBand=4200.0, acqSearchStep=300.0)
    raw = synth.make_record(sc, 18000 * 104)
    eng = Engine(s, codes=codes)
    eng.set_record(raw)
    host = eng.acquire(sv)
    dev = _device_results(eng, sv)
    for k in ("peakMetric", "codePhase", "carrFreq", "coarseBin"):
        assert np.array_equal(dev[k], host[k]), k
    assert np.count_nonzero(host["carrFreq"]) == 2
    eng.close()
    codes = icd_codes("BDS_B1I")
    sc = synth.default_scene_varb("BDS_B1I", codes, fs=18e6, nsat=2, seed=3)
    for x in sc.sats:
        x.cn0 = 48
    sv = sorted({x.prn for x in sc.sats} | {30})
    s = init_settings("BDS_B1I", samplingFreq=18e6, acqSatelliteList=sv)
    raw = synth.make_record(sc, 18000 * 11)
    eng = Engine(s, codes=codes)
    eng.set_record(raw)
    host = eng.acquire(sv)
    dev = _device_results(eng, sv)
    for k in ("peakMetric", "codePhase", "carrFreq", "coarseBin"):
        assert np.array_equal(dev[k], host[k]), k
    eng.close()


@pytest.mark.parametrize("n_gpus", [1, 2, 4])
def test_multi_engine_equals_one_gpu(n_gpus, tmp_path):
    """gc_multi_*: SV list dealt round-robin, channels in blocks, merged results bit-identical to one GPU; a record that
    ends mid-run stops every later channel whichever GPU it ran on (tracking.m:241-245)."""
    if _n_gpus() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    sc, s, sv, raw = _l1ca_case()
    one = Engine(s)
    one.set_record(raw)
    ref = one.acquire(sv)
    multi = MultiEngine(s, n_gpus=n_gpus)
    assert multi.n_gpus == n_gpus
    got = multi.acquire(sv, host_iq=raw[: 2 * 16368 * 44])          # the reference-facing call: longSignal from host memory
    for k in ("carrFreq", "codePhase", "peakMetric", "coarseBin", "coarseCodePhase"):
        assert np.array_equal(got[k], ref[k]), k
    multi.set_record(raw)
    got = multi.acquire(sv)
    for k in ("carrFreq", "codePhase", "peakMetric", "coarseBin", "coarseCodePhase"):
        assert np.array_equal(got[k], ref[k]), k
    ch = preRun(ref, s)
    prn = [c["PRN"] for c in ch]; af = [c["acquiredFreq"] for c in ch]; cp = [float(c["codePhase"]) for c in ch]
    assert sum(1 for p in prn if p) == 5 and prn[-1] == 0
    nE = s.msToProcess
    out1, vv1, vi1, d1 = one.track(prn, af, cp, nE)
    outm, vvm, vim, dm = multi.track(prn, af, cp, nE)
    assert np.array_equal(dm, d1) and np.array_equal(outm, out1) and np.array_equal(vvm, vv1) and np.array_equal(vim, vi1)
    assert multi.times()["track_ms"] > 0
    # tracking(fid, ...) through the file entry point
    path = tmp_path / "rec.bin"
    raw.tofile(path)
    outf, _, _, df = multi.track(prn, af, cp, nE, path=str(path))
    assert np.array_equal(df, d1) and np.array_equal(outf, out1)
    # short record: channel 0 runs out first, every later channel stays untouched on every GPU
    short = raw[: 2 * 16368 * 100]
    one.set_record(short)
    multi.set_record(short)
    out1, vv1, vi1, d1 = one.track(prn, af, cp, nE)
    outm, vvm, vim, dm = multi.track(prn, af, cp, nE)
    assert 0 < d1[0] < nE and np.all(d1[1:] == 0)
    assert np.array_equal(dm, d1) and np.array_equal(outm, out1) and np.array_equal(vvm, vv1)
    one.close()
    multi.close()
    # GLONASS frequency numbers (result index K + 7) through the same fan-out
    gsc = synth.default_scene_glo(fs=12e6, nsat=3, seed=17)
    for x in gsc.sats:
        x.cn0 = 47
    ks = list(range(-7, 7))
    gs = init_settings("GLO_GL1", samplingFreq=12e6, acqNonCohTime=4, acqSatelliteList=ks, msToProcess=60, numberOfChannels=4)
    graw = synth.make_record(gsc, 12000 * 110)
    g1 = Engine(gs)
    g1.set_record(graw)
    gref = g1.acquire(ks)
    gm = MultiEngine(gs, n_gpus=n_gpus)
    gm.set_record(graw)
    ggot = gm.acquire(ks)
    for k in ("carrFreq", "codePhase", "peakMetric", "coarseBin"):
        assert np.array_equal(ggot[k], gref[k]), k
    gch = preRun(gref, gs)
    gsv = [c["K"] if c["status"] != "-" else GC_SV_NONE for c in gch]
    o1 = g1.track(gsv, [c["acquiredFreq"] for c in gch], [float(c["codePhase"]) for c in gch], 60)
    om = gm.track(gsv, [c["acquiredFreq"] for c in gch], [float(c["codePhase"]) for c in gch], 60)
    assert np.array_equal(o1[0], om[0]) and np.array_equal(o1[3], om[3])
    g1.close()
    gm.close()


_WORKER = r"""
import os, sys
sys.path[:0] = [{root!r}, os.path.join({root!r}, "tests"), os.path.join({root!r}, "oracle")]
import numpy as np, torch, torch.distributed as dist
from cu_sdr_collection_b200 import Engine, init_settings, shard, synth
from helpers import scene
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
fs = 16.368e6
sc = scene(fs, nsat=6, seed=20260101, cn0=46)
s = init_settings(samplingFreq=fs, acqNonCohTime=5)
raw = synth.make_record(sc, 16368 * 44)                   # the same bytes on every rank (seeded)
eng = Engine(s, device=local)
eng.set_record(raw)
sv = shard.shard_units(s.acqSatelliteList, rank, world)   # ONE grid: 32 PRNs dealt round-robin
buf = torch.zeros(4 * 32, dtype=torch.float64, device=dev)
eng.acquire_device(sv, buf)
merged = shard.merge_device_results(shard.all_gather_device(buf), 32)   # one NCCL all-gather from device memory
if rank == 0:
    full = eng.acquire(s.acqSatelliteList)                # the whole grid on one GPU
    for k in ("peakMetric", "codePhase", "carrFreq", "coarseBin"):
        assert np.array_equal(merged[k], full[k]), k
    assert np.count_nonzero(full["carrFreq"]) >= 6
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", flush=True)
"""


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_grid_over_nccl_is_bit_identical(world, tmp_path):
    """One process per GPU: the 32-PRN grid dealt round-robin, results left on the device, ONE ncclAllGather, merged acqResults
    bit-identical to the one-GPU search."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", str(29500 + os.getpid() % 400), str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == world


def test_device_code_generators_equal_host_and_oracle():
    """codegen_kernel (one thread per SV, the reference's generate*code.m as bit-packed registers) == the same generators on the
    host == the oracle's statement-by-statement restatement."""
    from cu_sdr_collection_b200.engine import generate_code
    for signal, comp, fn, svs in (("GPS_L5C", 0, O.generateL5Icode, (1, 32)), ("GPS_L5C", 1, O.generateL5Qcode, (7,)),
                                  ("GAL_E5a", 0, O.generateE5aIcode, (1, 50)), ("GAL_E5a", 2, O.generateE5aQ_secondary, (3,)),
                                  ("GAL_E5b", 1, O.generateE5bQcode, (2,)), ("BDS_B2a", 0, O.generateB2aDataCode, (1, 63)),
                                  ("BDS_B2a", 1, O.generateB2aPilotCode, (20,)), ("BDS_B1I", 0, O.generateCAcode53, (1, 40, 58)),
                                  ("GPS_L2C", 0, O.generateCMcode, (1, 32)), ("GPS_L2C", 1, O.generateCLcode, (5,)),
                                  ("BDS_B1C", 0, O.generateDataBOC11, (1,)), ("BDS_B1C", 1, O.generatePilotBOC11, (63,)),
                                  ("BDS_B1C", 2, O.generatePilotBOC61, (19,))):
        for sv in svs:
            dev = generate_code(signal, sv, comp, device=0)
            assert np.array_equal(dev, generate_code(signal, sv, comp)), (signal, comp, sv)
            assert np.array_equal(dev, fn(sv)), (signal, comp, sv)
    assert np.array_equal(generate_code("GAL_E1C", 1, 0, device=0), O.generateE1Bcode(1)[0::2])
