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


# ---- R1CS, one assignment: shard the ROWS -------------------------------------------------
# A batch of assignment vectors is sharded like a batch of witnesses (shard_range over the batch).  For a single z
# every rank takes a contiguous block of rows of A, B, C with z replicated; a rank's first violated row is local to
# its block, so it is shifted by the block's first row before the same MIN all-reduce.
def shard_csr_rows(m, lo: int, hi: int):
    """rows [lo, hi) of a CSR matrix (row_ptr uint64[n+1], col uint32[nnz], coef_idx uint32[nnz])"""
    rp, col, ci = m
    rp = np.asarray(rp, dtype=np.uint64)
    e0, e1 = int(rp[lo]), int(rp[hi])
    return (rp[lo:hi + 1] - rp[lo]).astype(np.uint64), np.asarray(col)[e0:e1], np.asarray(ci)[e0:e1]


def shard_r1cs_rows(A, B, Cm, rank: int, world: int):
    """(A_r, B_r, C_r, first_row) of this rank's block of constraints"""
    n_rows = len(A[0]) - 1
    lo, hi = shard_range(n_rows, rank, world)
    return shard_csr_rows(A, lo, hi), shard_csr_rows(B, lo, hi), shard_csr_rows(Cm, lo, hi), lo


def global_first_row(verdicts: np.ndarray, first_row: int) -> np.ndarray:
    """local first violated rows -> global row indices (NO_FAIL where every row of the block holds)"""
    ff = first_fail_vector(verdicts)
    return np.where(ff == NO_FAIL, NO_FAIL, ff + np.int64(first_row))


def allreduce_min(local_ff: np.ndarray, device=None):
    """MIN all-reduce of equally shaped first-fail vectors (row-sharded R1CS: every rank holds every assignment)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(local_ff, dtype=np.int64)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.from_numpy(np.ascontiguousarray(local_ff, dtype=np.int64)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return t.cpu().numpy()
