"""Multi-GPU sharding of the hot path for one-process-per-GPU runs (``torch.distributed``; NCCL on GPUs, gloo in the CPU tests).

Acquisition units (PRNs, GLONASS frequency numbers, or (signal, SV) pairs of an all-constellation search) and tracking units
(channels) are independent - the PRN loop of acquisition.m:155, GLO_GL1/include/acquisition.m:172-183, the channel loop of
tracking.m:133 - so every rank takes a share of ONE unit list, works on it on its own B200 with its own copy of the IF window,
and the only exchange is one all-gather of the per-SV results (4 doubles per SV: peakMetric, codePhase, carrFreq, coarse bin)
straight from device memory (``Engine.acquire_device`` leaves them there).  The library does the same fan-out behind the C ABI
for callers without ``torch.distributed`` (``gc_multi_*``, ``MultiEngine``).
"""
from __future__ import annotations

import numpy as np

# Measured cost of one SV of each signal's acquisition at the reference's default initSettings.m (ms per SV on one B200,
# profiles/r01_bench_final.json: per-signal grid time / SVs searched) and the part of a signal's acquisition every rank that
# takes any of its SVs pays again (wipe-off + forward spectra, replica spectra are per SV).  Only ratios matter.
COST_MS_PER_SV = {
    "GPS_L1CA": 2.73 / 32, "GLO_GL1": 1.45 / 14, "GLO_GL2": 1.44 / 14, "BDS_B3I": 1.92 / 63, "GAL_E1C": 5.50 / 36,
    "GPS_L5C": 4.68 / 32, "GAL_E5a": 5.01 / 36, "GAL_E5b": 24.57 / 36, "BDS_B2a": 2.61 / 29, "BDS_B1I": 3.50 / 53,
    "GPS_L2C": 40.41 / 32, "BDS_B1C": 46.72 / 62,
}
FIXED_MS_PER_SIGNAL = 0.25


def shard_units(units, rank: int, world: int):
    """Round-robin share of a unit list (PRNs, FDMA channels, tracking channels) for one rank."""
    return list(units)[rank::world]


def result_index(sv: int, glonass: bool = False) -> int:
    """Index of an SV in the acqResults vectors: PRN - 1, or K + 7 for a GLONASS frequency number (MATLAB's K + 8,
    GLO_GL1/include/acquisition.m:212)."""
    return int(sv) + 7 if glonass else int(sv) - 1


def plan_pairs(pairs, world: int, cost=None, fixed=FIXED_MS_PER_SIGNAL):
    """Deal (signal, SV) pairs over ``world`` ranks by measured cost: longest-processing-time-first greedy on the per-pair cost,
    with the per-signal fixed part charged to a rank the first time it receives an SV of that signal (a rank that already pays
    it is preferred on ties).  ``pairs``: iterable of (signal, sv); ``cost``: {signal: ms per SV} (default: the measured table).
    Returns ``[{signal: [sv, ...]}, ...]`` per rank (SVs in the order given) and the predicted load per rank.  Deterministic."""
    cost = dict(COST_MS_PER_SV if cost is None else cost)
    pairs = list(pairs)
    order = sorted(range(len(pairs)), key=lambda i: (-cost.get(pairs[i][0], 1.0), i))
    load = [0.0] * world
    plan = [dict() for _ in range(world)]
    for i in order:
        sig, sv = pairs[i]
        c = cost.get(sig, 1.0)
        best = min(range(world), key=lambda r: (load[r] + c + (0.0 if sig in plan[r] else fixed), r))
        if sig not in plan[best]:
            plan[best][sig] = []
            load[best] += fixed
        plan[best][sig].append(sv)
        load[best] += c
    pos = {p: i for i, p in enumerate(pairs)}
    for r in range(world):
        for sig in plan[r]:
            plan[r][sig].sort(key=lambda sv: pos[(sig, sv)])
    return plan, load


def merge_device_results(gathered, n: int) -> dict:
    """acqResults from the all-gathered ``[world, 4 * n]`` (or already summed ``[4 * n]``) device buffers of
    ``Engine.acquire_device``: every SV was searched by exactly one rank and the other ranks hold zeros, so the merge is a sum
    (exact: x + 0 + ... + 0).  One D2H copy of 4 * n doubles."""
    t = gathered.reshape(-1, 4 * n).sum(dim=0) if gathered.dim() > 1 or gathered.numel() != 4 * n else gathered
    a = t.cpu().numpy()
    return dict(peakMetric=a[:n].copy(), codePhase=a[n:2 * n].copy(), carrFreq=a[2 * n:3 * n].copy(),
                coarseBin=a[3 * n:].astype(np.int32))


def all_gather_device(local, group=None):
    """One all-gather of the per-rank result buffers (device tensors; NCCL) -> ``[world, numel]``."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local.reshape(1, -1)
    world = dist.get_world_size(group)
    out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.reshape(-1), group=group)
    return out.reshape(world, -1)


def gather_acq_results(local: dict, sv_local, group=None, device=None, glonass: bool = False) -> dict:
    """All-gather the per-SV acquisition results of every rank (host arrays, e.g. from ``Engine.acquire``) and merge them into
    one ``acqResults``; each SV is searched by exactly one rank, unsearched entries stay zero.  ``glonass``: the unit list holds
    frequency numbers K = -7..13 stored at index K + 7."""
    import torch
    import torch.distributed as dist

    n = local["peakMetric"].shape[0]
    searched = np.zeros(n)
    idx = np.asarray([result_index(sv, glonass) for sv in sv_local], dtype=np.int64)
    if idx.size:
        assert idx.min() >= 0 and idx.max() < n, "SV id outside the result vectors"
        searched[idx] = 1.0
    rows = [local["peakMetric"], local["codePhase"], local["carrFreq"],
            np.asarray(local.get("coarseBin", np.zeros(n)), dtype=np.float64), searched]
    t = torch.from_numpy(np.stack(rows))
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        parts = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, t, group=group)
    else:
        parts = [t]
    merged = dict(peakMetric=np.zeros(n), codePhase=np.zeros(n), carrFreq=np.zeros(n), coarseBin=np.zeros(n, dtype=np.int32))
    for ptn in parts:
        a = ptn.cpu().numpy()
        m = a[4] != 0
        merged["peakMetric"][m] = a[0][m]
        merged["codePhase"][m] = a[1][m]
        merged["carrFreq"][m] = a[2][m]
        merged["coarseBin"][m] = a[3][m].astype(np.int32)
    return merged


def shard_channels(n_channels: int, rank: int, world: int):
    """Contiguous block of channel indices for one rank (the library's gc_multi_track uses the same split)."""
    per = (n_channels + world - 1) // world
    return list(range(min(rank * per, n_channels), min((rank + 1) * per, n_channels)))
