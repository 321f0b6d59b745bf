"""N > 1 host logic on CPU: witness-block sharding + the verdict MIN all-reduce, world_size 2, gloo."""
import os
import socket
import subprocess
import sys

import numpy as np

from tests.util import ROOT

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import importlib, zkb_loader
zkb_loader.load()
shard = importlib.import_module("zkir_b200.sharding")
import torch.distributed as dist
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
total = 37
lo, hi = shard.shard_range(total, rank, world)
# every rank "evaluates" its block: witness j fails at assertion 1000 + j when j % 5 == 0
v = np.zeros(hi - lo, dtype=np.dtype([("ok", "u1"), ("pad", "u1", (7,)), ("first_fail_seq", "<u8")]))
for k, j in enumerate(range(lo, hi)):
    v[k]["ok"] = 0 if j % 5 == 0 else 1
    v[k]["first_fail_seq"] = 1000 + j if j % 5 == 0 else (1 << 64) - 1
full = shard.allreduce_first_fail(shard.first_fail_vector(v), lo, hi, total)
exp = np.array([1000 + j if j % 5 == 0 else int(shard.NO_FAIL) for j in range(total)])
assert (full == exp).all(), (rank, full, exp)
# row-sharded R1CS: every rank holds a first violated row per assignment, already shifted to global row numbers
m = shard.allreduce_min(np.array([500 + rank, int(shard.NO_FAIL) if rank == 0 else 3, int(shard.NO_FAIL)], dtype=np.int64))
assert m.tolist() == [500, 3, int(shard.NO_FAIL)], m
print("rank", rank, "ok", lo, hi)
dist.destroy_process_group()
'''


def test_shard_ranges_cover_the_batch():
    import importlib
    import zkb_loader
    zkb_loader.load()
    shard = importlib.import_module("zkir_b200.sharding")
    for total in (1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            blocks = [shard.shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))


def test_verdict_allreduce_world2_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "rank 0 ok" in outs[0] and "rank 1 ok" in outs[1]


def test_r1cs_row_blocks_cover_the_system():
    import importlib
    import zkb_loader
    zkb_loader.load()
    shard = importlib.import_module("zkir_b200.sharding")
    c = importlib.import_module("zkir_b200.circuits")
    r = c.random_r1cs(1001, 50, 101, seed=2)
    for world in (1, 2, 3, 8):
        rows = 0
        for rank in range(world):
            A, B, C, row0 = shard.shard_r1cs_rows(r.A, r.B, r.C, rank, world)
            assert row0 == rows
            n = len(A[0]) - 1
            for full, part in ((r.A, A), (r.B, B), (r.C, C)):
                assert int(part[0][0]) == 0 and len(part[1]) == int(part[0][-1]) == len(part[2])
                e0 = int(full[0][row0])
                assert (np.asarray(full[1])[e0:e0 + len(part[1])] == part[1]).all()
                assert (np.diff(part[0].astype(np.int64)) == np.diff(np.asarray(full[0], dtype=np.int64)[row0:row0 + n + 1])).all()
            rows += n
        assert rows == r.n_rows
    v = np.zeros(3, dtype=np.dtype([("ok", "u1"), ("pad", "u1", (7,)), ("first_fail_seq", "<u8")]))
    v["ok"] = [1, 0, 0]
    v["first_fail_seq"] = [(1 << 64) - 1, 0, 17]
    assert shard.global_first_row(v, 500).tolist() == [int(shard.NO_FAIL), 500, 517]
