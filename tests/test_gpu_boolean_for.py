"""Config C5: Boolean-profile relation whose leaf gates only exist after nested For loops are unrolled on
the host; evaluated bit-sliced on the device.  Parity vs the Python oracle at small sizes, vs an
independent numpy evaluation of the same function at larger sizes, single witness and 64-witness batch."""
import numpy as np
import pytest

from oracle import evaluator as ev
from oracle import ir
from oracle import sieve_fbs as F
from oracle import workloads as wl
from tests.util import zkb

pytestmark = pytest.mark.gpu


def record(rel, w):
    z = zkb()
    b = z.GpuBackend(0)
    e = z.Evaluator(b)
    msgs = [ir.Witness(rel.header, [bytes([int(x)]) for x in w]), rel]
    e.ingest_source(z.Source.from_buffers([F.write_messages(msgs)]))
    return z, b, e, msgs


def chain_value(out_last):
    v = 0
    for k in range(8):
        v ^= int(out_last[k])
    return v


def test_small_against_python_oracle():
    rel, n_leaf = wl.boolean_for_relation(3, 4, 64)
    rng = np.random.default_rng(2)
    for trial in range(6):
        w = np.zeros(64, np.uint8) if trial == 0 else rng.integers(0, 2, 64).astype(np.uint8)
        z, b, e, msgs = record(rel, w)
        expected = ev.evaluate(msgs)
        assert e.get_violations() == expected
        assert b.stats()["ir_gates"] == n_leaf + 7 + 1
        outs = wl.boolean_for_expected_outputs(w, 3, 4).reshape(-1)
        got = [e.get(64 + k) for k in range(len(outs))]
        assert got == [int(x) for x in outs]
        assert (chain_value(outs[-32:]) == 0) == (expected == [])


def test_medium_single_and_bitsliced_batch():
    lo, li, n_wit = 7, 9, 4096
    rel, n_leaf = wl.boolean_for_relation(lo, li, n_wit)
    z, b, e, msgs = record(rel, np.zeros(n_wit, np.uint8))
    assert e.get_violations() == []                      # all-zero witness: TRUE
    st = b.stats()
    assert st["binary"] == 1 and st["ir_gates"] == n_leaf + 8
    # 64 independent witnesses, bit-sliced (one bit per witness in every word)
    rng = np.random.default_rng(3)
    n_batch = 64
    W = rng.integers(0, 2, size=(n_batch, n_wit, 1)).astype(np.uint8)
    W[5] = 0
    v = b.evaluate(None, W, n_batch)
    block = 2 << li
    base = n_wit + ((1 << lo) - 1) * block
    for j in range(n_batch):
        outs = wl.boolean_for_expected_outputs(W[j, :, 0], lo, li)
        assert bool(v[j]["ok"]) == (chain_value(outs[-1]) == 0), j
        if j in (0, 5, 31, 32, 63):
            probe = [n_wit + k for k in (0, 1, 2, 777, block, 5 * block + 3)] + [base + k for k in range(8)]
            vals = b.read_values(j, [e.value_handle(w_) for w_ in probe], 4)
            flat = outs.reshape(-1)
            assert vals == [int(flat[w_ - n_wit]) for w_ in probe]

