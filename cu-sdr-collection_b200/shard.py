"""Multi-GPU sharding of the hot path: one process per GPU, no collective on the data path.

Acquisition units (PRNs) and tracking units (channels) are independent, so each rank takes a
round-robin slice of them and processes it on its own B200 with its own copy of the IF window;
the only exchange is one all-gather of the per-PRN results (4 doubles per PRN: peakMetric,
codePhase, carrFreq, coarse bin) after acquisition and of the per-channel rows after tracking.
``torch.distributed`` is the plumbing (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_units(units, rank: int, world: int):
    """Round-robin slice of a unit list (PRNs, FDMA channels, tracking channels) for one rank."""
    return list(units)[rank::world]


def gather_acq_results(local: dict, sv_local, group=None, device=None) -> dict:
    """All-gather the per-PRN acquisition results of every rank and merge them into one
    ``acqResults`` (each PRN is searched by exactly one rank; unsearched entries are zero)."""
    import torch
    import torch.distributed as dist

    n = local["peakMetric"].shape[0]
    searched = np.zeros(n)
    searched[np.asarray(list(sv_local), dtype=np.int64) - 1] = 1.0
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
