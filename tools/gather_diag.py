"""Where a sharded acquisition step spends its time (dev tool; torchrun --nproc-per-node N tools/gather_diag.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
import torch
import torch.distributed as dist
from cu_sdr_collection_b200 import Engine, init_settings, shard, synth

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
fs = 16.368e6
sc = synth.default_scene(fs=fs, nsat=10, seed=20260101)
for s_ in sc.sats:
    s_.cn0 = max(s_.cn0, 44.0)
s = init_settings(samplingFreq=fs)
rec = synth.make_record_torch(sc, 16368 * 60, device=dev)
eng = Engine(s, device=local)
eng.set_record(rec)
ext = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
sv = shard.shard_units(list(s.acqSatelliteList), rank, world)
buf = torch.zeros(128, dtype=torch.float64, device=dev)

def step(mode):
    t = [time.perf_counter()]
    if mode == "sync":
        eng.acquire_device(sv, buf); t.append(time.perf_counter())
        g = shard.all_gather_device(buf); torch.cuda.synchronize(); t.append(time.perf_counter())
        r = shard.merge_device_results(g, 32); t.append(time.perf_counter())
    else:
        with torch.cuda.stream(ext):
            eng.acquire_device_async(sv, buf); t.append(time.perf_counter())
            g = shard.all_gather_device(buf); t.append(time.perf_counter())
            r = shard.merge_device_results(g, 32); t.append(time.perf_counter())
    return r, [(b - a) * 1e3 for a, b in zip(t, t[1:])]

for mode in ("sync", "async"):
    for _ in range(4):
        step(mode)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    acc = np.zeros(3); n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        r, dt = step(mode)
        acc += dt
    tot = (time.perf_counter() - t0) / n * 1e3
    st = eng.stats()
    print(f"rank {rank}/{world} {mode}: step {tot:.3f} ms = acquire {acc[0]/n:.3f} + gather {acc[1]/n:.3f} + merge {acc[2]/n:.3f}; kernels {st['acq_total_ms']:.3f} (fine {st['acq_fine_ms']:.3f}) acquired {int(np.count_nonzero(r['carrFreq']))}", flush=True)
if world > 1:
    dist.destroy_process_group()
