"""Multi-GPU plumbing: lncRNAs are independent (even the multiple-testing scope is one alignment,
`pipeline.py:763-776`), so the path shards by lncRNA with NO exchange during compute. The only
collective is the gather of results — per-rank counts and the variable-length ortholog rows —
over `torch.distributed` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from .pipeline import lpt_shards


def my_shard(batch, rank, world):
    """Deterministic LPT shard of `batch` for `rank` (every rank computes the same partition).
    Returns (lncRNA indices of this rank, sub-batch)."""
    if world == 1:
        return list(range(batch.n_lnc)), batch
    a0 = batch.aln_off[:-1]
    a1 = batch.aln_off[1:]
    weights = (batch.hsp_off[a1] - batch.hsp_off[a0]) + 1
    idx = lpt_shards(weights, world)[rank]
    return idx, batch.take_lnc(idx)


def gather_counts(counts, device):
    """all_gather of the five result counters -> int64 array [world, 5] on every rank."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(counts), dtype=torch.int64, device=device)
    world = dist.get_world_size()
    out = torch.zeros(world * t.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t)
    return out.view(world, -1).cpu().numpy()


def gather_rows(local, device):
    """Variable-length gather: every rank contributes a 1-D int64/float64 numpy array; every rank
    receives the list of all ranks' arrays (counts first, then one padded all_gather)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(local)).to(device)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=device)
    ns = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(ns, n)
    ns = ns.cpu().numpy()
    cap = int(ns.max()) if world else 0
    pad = torch.zeros(max(cap, 1), dtype=t.dtype, device=device)
    pad[:t.numel()] = t
    out = torch.zeros(world * max(cap, 1), dtype=t.dtype, device=device)
    dist.all_gather_into_tensor(out, pad)
    out = out.view(world, -1).cpu().numpy()
    return [out[r, :int(ns[r])].copy() for r in range(world)]


def gather_best_orthologs(res, lnc_idx, n_lnc_total, device):
    """Every rank ends with the global per-lncRNA view: has_ortholog[n_lnc_total] and the number of
    ortholog records per lncRNA — the part of the result `get_best_orthologs` consumes."""
    has = (res.best_aln >= 0).astype(np.int64)
    rows = gather_rows(np.stack([np.asarray(lnc_idx, dtype=np.int64), has], axis=1).reshape(-1), device)
    out = np.zeros(n_lnc_total, dtype=np.int64)
    seen = np.zeros(n_lnc_total, dtype=np.int64)
    for r in rows:
        r = r.reshape(-1, 2)
        out[r[:, 0]] = r[:, 1]
        seen[r[:, 0]] += 1
    assert (seen == 1).all(), "every lncRNA must be owned by exactly one rank"
    return out
