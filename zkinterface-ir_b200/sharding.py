"""Witness-batch sharding across the GPUs of one box (one process per GPU).

The path shards over independent witnesses: rank r evaluates a contiguous
block of the batch against its own replica of the levelized program; there is no data-path
collective.  The only exchange is one MIN all-reduce of the per-witness `first_fail` vector
(TRUE = NO_FAIL), equivalent to an AND of the verdict bits — a few KB over NVLink via NCCL.
"""
from __future__ import annotations

import numpy as np

NO_FAIL = np.int64(1) << 40   # larger than any assertion index


def shard_range(total: int, rank: int, world: int):
    """contiguous block [lo, hi) of a batch of `total` witnesses owned by `rank`"""
    return total * rank // world, total * (rank + 1) // world


def first_fail_vector(verdicts: np.ndarray) -> np.ndarray:
    """zkb_verdict array -> int64 first-fail indices, NO_FAIL where the statement holds"""
    return np.where(verdicts["ok"] == 1, NO_FAIL, verdicts["first_fail_seq"].astype(np.int64))


def allreduce_first_fail(local_ff: np.ndarray, lo: int, hi: int, total: int, device=None):
    """every rank contributes its block; returns the whole batch's first_fail vector on every rank"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(local_ff, dtype=np.int64)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    full = torch.full((total,), int(NO_FAIL), dtype=torch.int64, device=dev)
    full[lo:hi] = torch.from_numpy(np.ascontiguousarray(local_ff, dtype=np.int64)).to(dev)
    dist.all_reduce(full, op=dist.ReduceOp.MIN)
    return full.cpu().numpy()
