#!/usr/bin/env python3
"""Benchmark of the GNSS correlator hot path on B200 (contract: see the task's bench rules).

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Headline metric (BASELINE.json): acquisition PRN x Doppler cells/s on the GPS L1 C/A grid
32 PRN x 29 Doppler bins, FFT length 2N = 32736 @ 16.368 Msps, 20 non-coherent blocks
(configs[1]); the same JSON line carries the tracking leg (configs[2]: 12 channels x 60000 ms
correlate-and-dump) under "tracking".  A "step" is one full pass of the acquisition grid over one
record.  N > 1: every rank runs the full grid on its own record (weak scaling: the SV list grows
with N), followed by one all-gather of the per-PRN peak metrics.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 16.368e6
N_CODE = 16368
N_PRN, N_BINS, N_NONCOH = 32, 29, 20
CELLS = N_PRN * N_BINS
# algorithmic bytes per (PRN, Doppler) cell = 20 blocks x 2N samples x 2 B int8-IQ
# + the replica spectrum 2N x 8 B; per channel-ms = blksize x 2 B read + 15 doubles written.
BYTES_PER_CELL = N_NONCOH * 2 * N_CODE * 2 + 2 * N_CODE * 8          # 1,571,328
BYTES_PER_CHANNEL_MS = N_CODE * 2 + 15 * 8                            # 32,856


_OUT = sys.stdout


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # under load = upper half of the samples (idle gaps between phases pull the clock down)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's CPU implementation of the path.  The reference is MATLAB (no MATLAB/Octave in
    this image, nothing to compile into oracle/_ref), so this is the oracle port: the NumPy/SciPy
    restatement of acquisition.m with pocketfft on all host threads.  Each step is a bounded sample
    of the grid (SAMPLE_PRNS PRNs x 29 bins x 20 blocks)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import np_oracle as O
    from cu_sdr_collection_b200 import synth
    cores = os.cpu_count() or 1
    sample_prns = 2
    sc = synth.default_scene(fs=FS, nsat=8)
    raw = synth.make_record(sc, N_CODE * 42)
    s = O.Settings(samplingFreq=FS, acqSatelliteList=[sc.sats[0].prn, 1 if sc.sats[0].prn != 1 else 2])
    sig = O.read_acq_signal(raw, s)
    for _ in range(max(1, min(args.warmup, 1))):
        O.acquisition(sig, s, workers=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.acquisition(sig, s, workers=cores)
    dt = (time.perf_counter() - t0) / args.steps
    cells = sample_prns * N_BINS
    value = cells / dt
    line = {
        "impl": "reference", "metric": "acquisition PRNxDoppler cells/s", "value": value, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "GPS_L1CA acquisition grid: 32 PRN x 29 Doppler x 20 non-coherent blocks, "
                               "FFT length 32736 @ 16.368 Msps", "sample_per_step": f"{sample_prns} PRN x 29 bins"},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_prns} PRN x 29 Doppler bins x 20 blocks per step (NumPy/SciPy pocketfft oracle, "
                                   f"workers={cores}); work is linear in PRNs"},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_OUT, flush=True)


# --------------------------------------------------------------------------------- CPU baselines
def cpu_baselines(raw_acq, raw_trk_host, settings, chans):
    """Oracle timed on the host cores (rank 0, N = 1): bounded samples of both workloads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import np_oracle as O
    from helpers import c_tracking, to_oracle_settings
    cores = os.cpu_count() or 1
    so = to_oracle_settings(settings)
    so.acqSatelliteList = [chans[0][0], chans[1][0]]
    sig = O.read_acq_signal(raw_acq, so)
    O.acquisition(sig, O.Settings(samplingFreq=FS, acqSatelliteList=[1], acqNonCohTime=1), workers=cores)   # warm pocketfft plans
    t0 = time.perf_counter()
    O.acquisition(sig, so, workers=cores)
    dt = time.perf_counter() - t0
    acq = {"value": 2 * N_BINS / dt, "unit": "cells/s", "cores": cores, "kind": "port",
           "sample": f"2 PRN x 29 bins x 20 blocks of the same record in {dt:.2f} s "
                     f"(NumPy/SciPy pocketfft oracle, workers={cores})"}
    n_ms = 1500
    prn = [c[0] for c in chans]; af = [c[1] for c in chans]; cp = [c[2] for c in chans]
    settings_trk = to_oracle_settings(settings)
    t0 = time.perf_counter()
    out, vv, vi, done = c_tracking(raw_trk_host, settings_trk, prn, af, cp, n_ms, parallel=1)
    dt = time.perf_counter() - t0
    import ctypes
    from helpers import orc
    thr = orc().orc_num_threads()
    trk = {"value": int(done.sum()) / dt, "unit": "channel-ms/s", "cores": min(thr, len(prn)), "kind": "port",
           "sample": f"{len(prn)} channels x {n_ms} ms in {dt:.2f} s (C oracle, OpenMP over channels, {thr} threads)"}
    return acq, trk


# --------------------------------------------------------------------------------- this engine
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--track-ms", type=int, default=60000)
    ap.add_argument("--track-channels", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tracking", action="store_true")
    args = ap.parse_args()
    # ONE JSON line on stdout: keep the real stdout for it and send everything else that writes to fd 1 (NCCL's version banner,
    # library chatter) to stderr
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from cu_sdr_collection_b200 import Engine, init_settings, preRun, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's version banner off stdout: ONE JSON line there
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- synthetic 60 s record, one per rank (seed + rank), generated on the GPU (untimed setup)
    settings = init_settings(samplingFreq=FS, msToProcess=args.track_ms, numberOfChannels=args.track_channels)
    scene = synth.default_scene(fs=FS, nsat=10, seed=20260101 + rank)
    for sat in scene.sats:                      # strong enough that >= 10 SVs clear acqThreshold
        sat.cn0 = max(sat.cn0, 44.0)
    n_samples = N_CODE * (args.track_ms + 60)
    t_gen = time.perf_counter()
    rec = synth.make_record_torch(scene, n_samples, device=dev)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    eng = Engine(settings, device=local)
    eng.set_record(rec)
    ext = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        with torch.cuda.stream(ext):
            flush_buf.zero_()

    def gather_metrics(acq):
        if world == 1:
            return
        t = torch.from_numpy(np.stack([acq["peakMetric"], acq["codePhase"], acq["carrFreq"],
                                       acq["coarseBin"].astype(np.float64)])).to(dev)
        outl = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outl, t)                # the one NCCL all-gather of per-PRN peak metrics

    sampler = ClockSampler(local)
    # ---- acquisition, device-timed with inputs resident in HBM ------------------------------
    for _ in range(args.warmup):
        acq = eng.acquire()
        gather_metrics(acq)
    barrier()
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    rows_ms = cols_ms = 0.0
    launches = 0
    t_wall = time.perf_counter()
    for i in range(args.steps):
        flush_l2()
        ev[i][0].record(ext)
        acq = eng.acquire()
        ev[i][1].record(ext)
        gather_metrics(acq)
        st = eng.stats()
        rows_ms += st["corr_rows_ms"]; cols_ms += st["corr_cols_ms"]; launches += st["acq_launches"]
    barrier()
    t_wall = time.perf_counter() - t_wall
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    step_ms = max_over_ranks(dev_ms / args.steps)
    value = CELLS * world / (step_ms * 1e-3)
    st = eng.stats()
    n_chunks = max(1, st["corr_row_launches"])
    rows_launch_ms = rows_ms / args.steps / n_chunks
    cells_per_launch = CELLS / n_chunks

    # ---- acquisition end to end through the reference-facing call with HOST buffers ---------
    n_acq_samples = N_CODE * 42
    host_acq = torch.empty(2 * n_acq_samples, dtype=torch.int8).pin_memory()
    host_acq.copy_(rec[: 2 * n_acq_samples])
    host_acq_np = host_acq.numpy()
    eng_e2e = Engine(settings, device=local)
    for _ in range(2):
        eng_e2e.acquire(host_iq=host_acq_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a2 = eng_e2e.acquire(host_iq=host_acq_np)      # H2D of longSignal + search + D2H of acqResults
        gather_metrics(a2)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) / args.steps * 1e3)
    e2e_value = CELLS * world / (e2e_ms * 1e-3)
    d2h_acq = 32 * (3 * 8 + 2 * 4) + 32 * 16 + 8
    assert np.array_equal(a2["carrFreq"], acq["carrFreq"]) and np.array_equal(a2["codePhase"], acq["codePhase"])
    eng_e2e.close()

    # ---- tracking: 12 channels x 60000 ms (configs[2]) ---------------------------------------
    tracking = None
    chans = []
    ch = preRun(acq, settings)
    for c in ch:
        if c["PRN"]:
            chans.append((c["PRN"], c["acquiredFreq"], float(c["codePhase"])))
    n_found = len(chans)
    while len(chans) < args.track_channels and n_found:       # fill the 12 channels (repeat SVs if fewer were found)
        chans.append(chans[len(chans) % n_found])
    if not args.no_tracking and chans:
        prn = [c[0] for c in chans]; af = [c[1] for c in chans]; cp = [c[2] for c in chans]
        eng.track(prn, af, cp, min(2000, args.track_ms))        # warm-up
        barrier()
        reps, kms, wall = 2, [], []
        for _ in range(reps):
            flush_l2()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out, vv, vi, done = eng.track(prn, af, cp, args.track_ms)
            wall.append(time.perf_counter() - t0)
            kms.append(eng.stats()["track_kernel_ms"])
        barrier()
        k_ms = max_over_ranks(sum(kms) / reps)
        units = int(done.sum())
        lock = float(np.mean(np.abs(out[:, 3, 100:])) / max(1e-9, np.mean(np.abs(out[:, 7, 100:]))))
        # end to end: H2D of the whole record from pinned host memory + kernel + D2H of trackResults
        host_rec = torch.empty(rec.numel(), dtype=torch.int8).pin_memory()
        host_rec.copy_(rec)
        eng_t = Engine(settings, device=local)
        eng_t.set_record_host_ptr(host_rec.data_ptr(), host_rec.numel())      # untimed warm-up: the same call once, so that the
        eng_t.track(prn, af, cp, args.track_ms)                               # timed one measures copies + kernel, not cudaMalloc
        barrier()
        t0 = time.perf_counter()
        eng_t.set_record_host_ptr(host_rec.data_ptr(), host_rec.numel())
        out2, _, _, done2 = eng_t.track(prn, af, cp, args.track_ms)
        barrier()
        e2e_t = max_over_ranks(time.perf_counter() - t0)
        assert np.array_equal(done2, done)
        eng_t.close()
        pk, pk_src = peaks()
        t_ach = units * BYTES_PER_CHANNEL_MS / (k_ms * 1e-3) / 1e9
        tracking = {
            "metric": "tracking channel-ms/s", "value": units * world / (k_ms * 1e-3), "unit": "channel-ms/s",
            "channels": len(prn), "ms": args.track_ms, "kernel_ms": k_ms, "us_per_epoch": k_ms * 1e3 / args.track_ms,
            "prompt_I_over_Q": lock,
            "e2e": {"value": units * world / e2e_t, "unit": "channel-ms/s", "h2d_bytes_per_step": int(host_rec.numel()),
                    "d2h_bytes_per_step": int(out2.nbytes)},
            "roofline": {"bound": "hbm", "achieved": t_ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": t_ach / pk["hbm_gbs"], "traffic": None,
                         "note": "latency-bound: 60000-epoch loop-carried dependency per channel, 12 CTAs on 148 SMs"},
        }
        # bandwidth regime: many independent channels (same kernel), 1000 ms
        big = [chans[i % len(chans)] for i in range(592)]
        eng.track([c[0] for c in big], [c[1] for c in big], [c[2] for c in big], 1000)
        bk = max_over_ranks(eng.stats()["track_kernel_ms"])
        b_ach = 592 * 1000 * BYTES_PER_CHANNEL_MS / (bk * 1e-3) / 1e9
        tracking["batch_592ch_1000ms"] = {"value": 592 * 1000 * world / (bk * 1e-3), "unit": "channel-ms/s", "kernel_ms": bk,
                                          "roofline_frac": b_ach / pk["hbm_gbs"], "achieved_gbs": b_ach}
    # ---- widened rows: GLONASS L1 at the reference's default settings ----------
    glonass = None
    if not args.no_tracking:
        gs = init_settings("GLO_GL1", msToProcess=5000, numberOfChannels=12)
        gscene = synth.default_scene_glo(fs=gs.samplingFreq, nsat=8, seed=20260101 + rank)
        for sat in gscene.sats:
            sat.cn0 = max(sat.cn0, 44.0)
        grec = synth.make_record_torch(gscene, 12000 * 5060, device=dev)
        geng = Engine(gs, device=local)
        geng.set_record(grec)
        for _ in range(2):
            gacq = geng.acquire()
        gst = geng.stats()
        gcells = len(gs.acqSatelliteList) * 21
        gch = preRun(gacq, gs)
        gsv = [c["K"] for c in gch if c["status"] == "T"]
        glonass = {"acquisition": {"value": gcells * world / (max_over_ranks(gst["acq_total_ms"]) * 1e-3), "unit": "cells/s",
                                   "workload": "GLO_GL1 defaults: 14 frequency channels x 21 Doppler x 20 blocks, FFT length 24000 "
                                               "@ 12 Msps (fused 30 x 32 x 25 plan)", "ms": gst["acq_total_ms"],
                                   "n_acquired": int(gst["n_acquired"])}}
        if gsv:
            while len(gsv) < 12:
                gsv.append(gsv[len(gsv) % len(set(gsv))])
            byK = {c["K"]: c for c in gch if c["status"] == "T"}
            geng.track(gsv, [byK[k]["acquiredFreq"] for k in gsv], [float(byK[k]["codePhase"]) for k in gsv], 5000)
            gk = max_over_ranks(geng.stats()["track_kernel_ms"])
            glonass["tracking"] = {"value": 12 * 5000 * world / (gk * 1e-3), "unit": "channel-ms/s", "channels": 12, "ms": 5000,
                                   "us_per_epoch": gk * 1e3 / 5000}
        geng.close()
        del grec
    # ---- widened rows (BASELINE configs[3]): Galileo E1 36 PRN x 81 Doppler x 4 ms @ 20 Msps
    #      (FFT length 160000, two replicas, fused 200 x 32 x 25 plan) and GPS L5C at the reference defaults
    #      (32 PRN x 21 Doppler x 25 blocks x 2 replicas, FFT length 36000, fused plan) ----------------------
    widened = None
    if not args.no_tracking:
        from cu_sdr_collection_b200.codes import standin_codes, standin_e1_codes
        widened = {}
        e1codes = standin_e1_codes()
        es = init_settings("GAL_E1C", samplingFreq=20e6, acqSearchBand=6000.0, acqSearchStep=150.0)     # 81 bins
        escene = synth.default_scene_e1c(e1codes, fs=20e6, nsat=6, seed=20260101 + rank)
        for sat in escene.sats:
            sat.cn0 = max(sat.cn0, 46.0)
        erec = synth.make_record_torch(escene, 80000 * 43, device=dev)
        eeng = Engine(es, device=local, codes=e1codes)
        eeng.set_record(erec)
        for _ in range(2):
            eacq = eeng.acquire()
        est = eeng.stats()
        ecells = len(es.acqSatelliteList) * 81
        widened["gal_e1c_acquisition"] = {
            "value": ecells * world / (max_over_ranks(est["acq_total_ms"]) * 1e-3), "unit": "cells/s", "ms": est["acq_total_ms"],
            "workload": "GAL_E1C 36 PRN x 81 Doppler x 1 block x 2 replicas (E1B + E1C), FFT length 160000 @ 20 Msps "
                        "(BASELINE.json configs[3] grid; fused 200 x 32 x 25 plan with two-level 20 x 10 columns, stand-in memory codes)",
            "n_acquired": int(est["n_acquired"]), "acq_path": int(est["acq_path"])}
        eeng.close()
        del erec
        lcodes = standin_codes("GPS_L5C")
        ls = init_settings("GPS_L5C")
        lscene = synth.default_scene_fam5("GPS_L5C", lcodes, fs=18e6, nsat=8, seed=20260101 + rank)
        for sat in lscene.sats:
            sat.cn0 = max(sat.cn0, 44.0)
        lrec = torch.from_numpy(synth.make_record(lscene, 18000 * 44)).to(dev)
        leng = Engine(ls, device=local, codes=lcodes)
        leng.set_record(lrec)
        for _ in range(2):
            lacq = leng.acquire()
        lst = leng.stats()
        widened["gps_l5c_acquisition"] = {
            "value": 32 * 21 * world / (max_over_ranks(lst["acq_total_ms"]) * 1e-3), "unit": "cells/s", "ms": lst["acq_total_ms"],
            "workload": "GPS_L5C defaults: 32 PRN x 21 Doppler x 25 blocks x 2 replicas (I5 + Q5), FFT length 36000 @ 18 Msps "
                        "(fused 45 x 32 x 25 plan, stand-in codes)",
            "n_acquired": int(lst["n_acquired"]), "acq_path": int(lst["acq_path"])}
        leng.close()
        del lrec
    # ---- BASELINE configs[4]: every constellation's acquisition at the reference's default initSettings.m, one after the
    #      other on this GPU (per rank; N ranks = N independent receivers).  Cells = SVs x Doppler rows searched. ---------
    if not args.no_tracking and widened is not None:
        from cu_sdr_collection_b200.codes import standin_b1c_codes, standin_varb_codes
        from cu_sdr_collection_b200.settings import samples_per_code
        allc = {}
        tot_cells, tot_ms = 0, 0.0

        def run_sig(name, st_, codes_, scene_, periods, cells):
            nonlocal tot_cells, tot_ms
            n_ = samples_per_code(st_)
            r_ = torch.from_numpy(synth.make_record(scene_, n_ * periods + 64)).to(dev)
            e_ = Engine(st_, device=local, codes=codes_) if codes_ is not None else Engine(st_, device=local)
            e_.set_record(r_)
            for _ in range(2):
                e_.acquire()
            ms_ = max_over_ranks(e_.stats()["acq_total_ms"])
            allc[name] = {"cells": cells, "ms": ms_, "fft_len": int(e_.stats()["fft_len"]), "n_acquired": int(e_.stats()["n_acquired"])}
            tot_cells += cells; tot_ms += ms_
            e_.close()
            del r_

        sd = 20260101 + rank
        st_ = init_settings("GPS_L1CA")
        run_sig("GPS_L1CA", st_, None, synth.default_scene(fs=18e6, nsat=6, seed=sd), 42, 32 * 29)
        st_ = init_settings("GLO_GL1")
        run_sig("GLO_GL1", st_, None, synth.default_scene_glo(fs=12e6, nsat=5, seed=sd), 42, 14 * 21)
        st_ = init_settings("GLO_GL2")
        run_sig("GLO_GL2", st_, None, synth.default_scene_glo(fs=12e6, nsat=5, seed=sd + 1, freqSpacing=437.5e3), 42, 14 * 21)
        st_ = init_settings("BDS_B3I")
        run_sig("BDS_B3I", st_, None, synth.default_scene_b3i(fs=18e6, nsat=5, seed=sd), 22, 63 * 21)
        st_ = init_settings("GAL_E1C")
        run_sig("GAL_E1C", st_, e1codes, synth.default_scene_e1c(e1codes, fs=18e6, nsat=4, seed=sd), 42, 36 * 94)
        for sg, per, bins in (("GPS_L5C", 42, 21), ("GAL_E5a", 102, 21), ("GAL_E5b", 102, 168), ("BDS_B2a", 17, 21)):
            cd = standin_codes(sg)
            st_ = init_settings(sg)
            run_sig(sg, st_, cd, synth.default_scene_fam5(sg, cd, fs=18e6, nsat=4, seed=sd), per, len(st_.acqSatelliteList) * bins)
        cd = standin_varb_codes("BDS_B1I")
        st_ = init_settings("BDS_B1I")
        run_sig("BDS_B1I", st_, cd, synth.default_scene_varb("BDS_B1I", cd, fs=18e6, nsat=4, seed=sd), 11, 53 * 81)
        cd = standin_varb_codes("GPS_L2C")
        st_ = init_settings("GPS_L2C")
        run_sig("GPS_L2C", st_, cd, synth.default_scene_varb("GPS_L2C", cd, fs=8e6, nsat=3, seed=sd), 3, 32 * 801)
        cd = standin_b1c_codes()
        st_ = init_settings("BDS_B1C")
        run_sig("BDS_B1C", st_, cd, synth.default_scene_varb("BDS_B1C", cd, fs=18e6, nsat=3, seed=sd), 2, 62 * 201)
        widened["all_constellation_acquisition"] = {
            "value": tot_cells * world / (tot_ms * 1e-3), "unit": "cells/s", "cells": tot_cells, "ms": tot_ms,
            "sv_signal_pairs": 32 + 14 + 14 + 63 + 36 + 32 + 36 + 36 + 29 + 53 + 32 + 62,
            "workload": "BASELINE.json configs[4]: the twelve signal folders' acquisitions at their default initSettings.m, run "
                        "back to back on one GPU per rank (cells = SVs x Doppler rows; every length has a fused plan: 36000, 24000, 144000, "
                        "72000, 320000, 360000; stand-in codes where the reference's are data)",
            "per_signal": allc}
    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline (oracle on the host cores; rank 0, N = 1 only) -------------------------
    cpu_acq = cpu_trk = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and chans:
        raw_trk = rec[: 2 * N_CODE * 1600].cpu().numpy()
        cpu_acq, cpu_trk = cpu_baselines(host_acq_np, raw_trk, settings, chans)
        if tracking is not None:
            tracking["cpu_baseline"] = cpu_trk

    if rank == 0:
        pk, pk_src = peaks()
        achieved = cells_per_launch * BYTES_PER_CELL / (rows_launch_ms * 1e-3) / 1e9
        step_ach = CELLS * BYTES_PER_CELL / (step_ms * 1e-3) / 1e9
        prof = os.path.join(ROOT, "profiles", "r01_traffic.json")   # ncu --set full capture of the dominant kernel
        traffic = None
        if os.path.exists(prof):                                    # per launch like `achieved`: bytes per cell x cells of one launch
            tj = json.load(open(prof))
            traffic = tj["inv_rows_dram_bytes_per_cell"] * cells_per_launch if "inv_rows_dram_bytes_per_cell" in tj else tj.get("inv_rows_dram_bytes_per_launch")
        line = {
            "metric": "acquisition PRNxDoppler cells/s", "value": value, "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "GPS_L1CA acquisition grid: 32 PRN x 29 Doppler x 20 non-coherent blocks, FFT length "
                                   "32736 @ 16.368 Msps, 8-bit complex IF (BASELINE.json configs[1]); per GPU",
                       "record": f"synthetic 60 s IF record per rank, {rec.numel()} B resident in HBM, seed 20260101+rank "
                                 f"(generated in {t_gen:.1f} s, untimed)",
                       "l2": "flushed between timed steps (256 MiB memset); per-step working set 5 GB > 126 MB L2",
                       "timing": "CUDA events on the engine's stream per step, max over ranks",
                       "parallelism": f"{world} x (full grid on own record) + 1 NCCL all-gather of per-PRN metrics"},
            "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": 2 * n_acq_samples,
                    "d2h_bytes_per_step": d2h_acq, "ms_per_step": e2e_ms,
                    "call": "gc_acquire_host (host longSignal -> acqResults on host)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / pk["hbm_gbs"], "traffic": traffic,
                         "kernel": "inv_rows_kernel (spectrum multiply + inverse 31x32 row DFTs of the 33x32x31 prime-factor "
                                   "transform, packed fp32x2 codelets)",
                         "launch_ms": rows_launch_ms, "cells_per_launch": cells_per_launch, "peak_source": pk_src,
                         "whole_step_achieved": step_ach, "whole_step_frac": step_ach / pk["hbm_gbs"],
                         "kernel_share_of_step": {"inv_rows": rows_ms / args.steps / step_ms, "inv_cols": cols_ms / args.steps / step_ms},
                         "note": "FFT stages carry about 69 flop per algorithmic byte; the rows kernel is bound by FP32 issue and "
                                 "L2 throughput, the column kernel by the HBM read of the work buffer (6.5 TB/s)"},
            "cpu_baseline": cpu_acq,
            "clocks": clocks,
            "n_acquired": int(st["n_acquired"]),
            "wall_ms_per_step_incl_flush_and_gather": t_wall / args.steps * 1e3,
            "tracking": tracking,
            "glonass": glonass,
            "widened": widened,
        }
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
