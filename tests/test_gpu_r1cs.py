"""GPU parity of the R1CS check (CSR mod-p SpMV + Hadamard) against the oracle's evaluation of the
gate expansion `zkif-to-ir` produces for the same system (producers/from_r1cs.rs:27-141)."""
import numpy as np
import pytest

from oracle import evaluator as ev
from oracle import fixtures as fx
from tests.util import FIELDS, circuits, zkb

pytestmark = pytest.mark.gpu


def to_zkif(c, r, z, n_public=8):
    """R1cs + assignment -> the arguments of fixtures.r1cs_to_gates"""
    eb = c.elem_bytes(r.p)
    le = lambda v: int(v).to_bytes(eb, "little")
    instance_vars = [(i, le(z[i])) for i in range(1, 1 + n_public)]
    witness_ids = list(range(1 + n_public, r.n_vars))
    witness_values = [le(z[i]) for i in witness_ids]
    cons = []
    for row in range(r.n_rows):
        lcs = []
        for rp, col, ci in (r.A, r.B, r.C):
            lcs.append([(int(col[e]), le(r.coefs[int(ci[e])])) for e in range(int(rp[row]), int(rp[row + 1]))])
        cons.append(tuple(lcs))
    return le(r.p - 1), instance_vars, witness_ids, cons, witness_values


def oracle_first_failing_row(c, r, z):
    msgs, _ = fx.r1cs_to_gates(*to_zkif(c, r, z))
    tb = ev.TracingBackend()
    e = ev.Evaluator.from_messages(msgs, tb)
    v = e.get_violations()
    if not v:
        return -1
    assert v[0].startswith("Wire_")
    return len(tb.asserts) - 1      # one AssertZero per row, in row order: the failing one is the last reached


@pytest.mark.parametrize("name", ["p101", "goldilocks", "bn254"])
def test_r1cs_matches_gate_expansion(name):
    c = circuits()
    z_ = zkb()
    p = FIELDS[name]
    r = c.random_r1cs(400, 60, p, seed=3)
    good = c.r1cs_assignment(r, seed=1)
    bad1 = list(good)
    bad1[r.n_free + 1 + 37] = (bad1[r.n_free + 1 + 37] + 1) % p          # a slack: row 37 (and readers) fail
    bad2 = list(good)
    bad2[5] = (bad2[5] + 7) % p                                          # a free variable
    zs = [good, bad1, bad2, good]
    b = z_.GpuBackend(0)
    b.set_field(p)
    b.r1cs_load(r.A, r.B, r.C, r.coef_table, r.n_vars)
    zb = np.stack([c.assignment_bytes(z, p) for z in zs])
    v = b.r1cs_check(zb)
    exp = [oracle_first_failing_row(c, r, z) for z in zs]
    got = [(-1 if x["ok"] else int(x["first_fail_seq"])) for x in v]
    assert got == exp
    assert exp[0] == -1 and exp[1] == min(37, c.r1cs_first_row_reading(r, r.n_free + 1 + 37))
    assert exp[2] == c.r1cs_first_row_reading(r, 5)
    # single assignment, device-resident re-run
    b.r1cs_upload(zb[1:2])
    assert int(b.r1cs_run()[0]["first_fail_seq"]) == exp[1]
    assert b.timing()["levels_ms"] > 0


def test_r1cs_zkif_example():
    # producers/from_r1cs.rs:178-221: x*x = xx ; y*y = yy ; 1*(xx+yy) = zz over p = 101
    z_ = zkb()
    c = circuits()
    b = z_.GpuBackend(0)
    b.set_field(101)
    coef = np.array([[1, 0, 0, 0]], dtype=np.uint8)
    A = (np.array([0, 1, 2, 3], np.uint64), np.array([1, 2, 0], np.uint32), np.zeros(3, np.uint32))
    B = (np.array([0, 1, 2, 4], np.uint64), np.array([1, 2, 4, 5], np.uint32), np.zeros(4, np.uint32))
    C = (np.array([0, 1, 2, 3], np.uint64), np.array([4, 5, 3], np.uint32), np.zeros(3, np.uint32))
    b.r1cs_load(A, B, C, coef, 6)
    z = np.zeros((2, 6, 4), dtype=np.uint8)
    z[0, :, 0] = [1, 3, 4, 25, 9, 16]
    z[1, :, 0] = [1, 3, 4, 26, 9, 16]
    v = b.r1cs_check(z)
    assert [int(x["ok"]) for x in v] == [1, 0] and int(v[1]["first_fail_seq"]) == 2
    z[0, 0, 0] = 2
    with pytest.raises(z_.ZkbError) as e:
        b.r1cs_check(z)
    assert str(e.value) == "value for instance id:0 should be a constant 1"


@pytest.mark.parametrize("tile", ["0", "3"])
def test_r1cs_batch_tiles(tile, monkeypatch):
    monkeypatch.setenv("ZKB_TILE_LOG2", tile)
    c = circuits()
    z_ = zkb()
    p = FIELDS["bls381"]
    r = c.random_r1cs(300, 50, p, seed=9)
    good = c.r1cs_assignment(r, seed=2)
    n = 21
    zs = []
    exp = []
    for j in range(n):
        z = list(good)
        if j % 4 == 1:
            row = (j * 13) % r.n_rows
            var = r.n_free + 1 + row
            z[var] = (z[var] + 1) % p
            exp.append(min(row, c.r1cs_first_row_reading(r, var)))
        else:
            exp.append(-1)
        zs.append(c.assignment_bytes(z, p))
    b = z_.GpuBackend(0)
    b.set_field(p)
    b.r1cs_load(r.A, r.B, r.C, r.coef_table, r.n_vars)
    v = b.r1cs_check(np.stack(zs))
    assert [(-1 if x["ok"] else int(x["first_fail_seq"])) for x in v] == exp


@pytest.mark.parametrize("name", ["goldilocks", "bn254"])
def test_r1cs_wide_tiles_use_the_unsplit_layout(name):
    """tiles of >= 32 assignments take layout kind 1 (one class per matrix, ones tagged inside the general class):
    70 assignments, ragged last warp, a third of them violated at known rows; coefficient table with 0 and p + 1"""
    c = circuits()
    z_ = zkb()
    p = FIELDS[name]
    eb = c.elem_bytes(p)
    r = c.random_r1cs(700, 90, p, seed=17)
    table = np.concatenate([r.coef_table, c.le_bytes(0, eb)[None]])
    zero_idx = len(r.coefs)
    good = c.r1cs_assignment(r, seed=4)
    # zero out some coefficients of B and fix the slack values accordingly through a fresh assignment
    B = (r.B[0], r.B[1], r.B[2].copy())
    B[2][::9] = zero_idx
    coefs = r.coefs + [0]
    r.B = B
    r.coefs = coefs
    good = c.r1cs_assignment(r, seed=4)
    n = 70
    zs, exp = [], []
    for j in range(n):
        zv = list(good)
        if j % 3 == 1:
            row = (j * 29) % r.n_rows
            var = r.n_free + 1 + row
            zv[var] = (zv[var] + 1 + j) % p
            exp.append(min(row, c.r1cs_first_row_reading(r, var)))
        else:
            exp.append(-1)
        zs.append(c.assignment_bytes(zv, p))
    b = z_.GpuBackend(0)
    b.set_field(p)
    b.r1cs_load(r.A, r.B, r.C, table, r.n_vars)
    v = b.r1cs_check(np.stack(zs))
    assert [(-1 if x["ok"] else int(x["first_fail_seq"])) for x in v] == exp
    # the same system, one assignment at a time (layout kind 0), agrees
    for j in (0, 1, 4):
        b.r1cs_upload(np.stack(zs[j:j + 1]))
        x = b.r1cs_run()[0]
        assert (-1 if x["ok"] else int(x["first_fail_seq"])) == exp[j]


def test_r1cs_row_sharding_over_devices():
    """one assignment, constraints sharded in row blocks over the GPUs (z replicated), first violated row =
    MIN over the blocks of (block's first row + local row).  Uses every visible device (1 on the test box: the blocks
    then run one after the other on device 0); the MIN all-reduce itself is covered by tests/test_sharding_gloo.py."""
    import importlib
    import torch
    c = circuits()
    z_ = zkb()
    shard = importlib.import_module("zkir_b200.sharding")
    p = FIELDS["bn254"]
    r = c.random_r1cs(3000, 200, p, seed=23)
    good = c.r1cs_assignment(r, seed=6)
    bad = list(good)
    row = 1777
    bad[r.n_free + 1 + row] = (bad[r.n_free + 1 + row] + 5) % p
    want = min(row, c.r1cs_first_row_reading(r, r.n_free + 1 + row))
    n_dev = max(1, torch.cuda.device_count())
    world = 4
    for zv, exp in ((good, -1), (bad, want)):
        zb = c.assignment_bytes(zv, p)[None]
        parts = []
        for rank in range(world):
            A, B, Cm, row0 = shard.shard_r1cs_rows(r.A, r.B, r.C, rank, world)
            b = z_.GpuBackend(rank % n_dev)
            b.set_field(p)
            b.r1cs_load(A, B, Cm, r.coef_table, r.n_vars)
            parts.append(shard.global_first_row(b.r1cs_check(zb), row0))
            b.close()
        ff = np.minimum.reduce(parts)
        assert (-1 if ff[0] == shard.NO_FAIL else int(ff[0])) == exp
