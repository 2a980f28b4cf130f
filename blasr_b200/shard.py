"""Read sharding over the GPUs of one box: no collective on the data path.

Every (read, candidate) job is independent, so GPU g of G takes jobs g, g+G, ... exactly like the reference's own
`-start S -stride N` partial runs (Blasr.cpp:4057-4058); results come back to rank 0 in read order.  torch.distributed
is plumbing only (rendezvous, barrier, the max-over-ranks timing reduction and the ordered gather of small result
records); the alignment kernels never touch it.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def shard_indices(n_jobs: int, rank: int, world: int) -> np.ndarray:
    """Job indices of this rank: start=rank, stride=world."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return np.arange(rank, n_jobs, world, dtype=np.int64)


def merge_in_read_order(n_jobs: int, per_rank: Sequence[np.ndarray]) -> np.ndarray:
    """Inverse of shard_indices: interleave the per-rank result arrays back into read order."""
    world = len(per_rank)
    first = next((a for a in per_rank if len(a)), None)
    out = np.zeros(n_jobs, dtype=first.dtype if first is not None else np.int64)
    for r, a in enumerate(per_rank):
        idx = shard_indices(n_jobs, r, world)
        if len(a) != len(idx):
            raise ValueError(f"rank {r} returned {len(a)} results for {len(idx)} jobs")
        out[idx] = a
    return out


def all_reduce_scalar(x: float, op: str, device=None) -> float:
    """max / sum of a python scalar over the process group (identity without one)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def gather_records(local: np.ndarray, n_jobs: int, device=None):
    """Gathers fixed-size per-job records (e.g. score, qPos, tPos) to every rank, in read order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    width = local.shape[1] if local.ndim == 2 else 1
    per = (n_jobs + world - 1) // world
    buf = torch.zeros((per, width), dtype=torch.int64, device=device or "cpu")
    buf[:len(local)] = torch.from_numpy(np.ascontiguousarray(local.reshape(len(local), width).astype(np.int64)))
    outs: List = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    parts = [o.cpu().numpy()[:len(shard_indices(n_jobs, r, world))] for r, o in enumerate(outs)]
    merged = np.zeros((n_jobs, width), dtype=np.int64)
    for r, a in enumerate(parts):
        merged[shard_indices(n_jobs, r, world)] = a
    return merged


def run_on_devices(batch, run, devices: Sequence[int], threads_per_device: int = 1):
    """One process, many devices -- how a pthread blasr drives a multi-GPU box: worker thread w of W = len(devices) *
    threads_per_device owns a context on device devices[w % len(devices)], takes jobs w, w+W, ... (the reference's
    -start / -stride, Blasr.cpp:4057-4058) and calls run(aligner, sub_batch) -> per-job records (a structured / 2-D
    numpy array, one row per job).  Returns the records merged back into read order.  No collective is involved."""
    import threading
    from .align import Aligner
    W = len(devices) * threads_per_device
    parts: List = [None] * W
    errs: List = []

    def work(w):
        try:
            idx = shard_indices(batch.n, w, W)
            al = Aligner(devices[w % len(devices)])
            try:
                parts[w] = run(al, batch.slice(idx)) if len(idx) else None
            finally:
                al.close()
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(w,)) for w in range(W)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    if errs:
        raise errs[0]
    first = next(p for p in parts if p is not None)
    out = np.zeros((batch.n,) + first.shape[1:], dtype=first.dtype)
    for w, p in enumerate(parts):
        if p is not None:
            out[shard_indices(batch.n, w, W)] = p
    return out
