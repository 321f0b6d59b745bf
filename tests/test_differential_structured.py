"""Differential tests on random STRUCTURED relations (tests/gen_programs.py).
CPU part (no GPU): the C++ flattener records, op for op, what the oracle's Evaluator asks of its backend.
GPU part: verdict / violation text / every value / live wires equal to the oracle."""
import numpy as np
import pytest

from oracle import evaluator as ev
from oracle import ir
from oracle import sieve_fbs as F
from tests.gen_programs import Gen
from tests.util import zkb

KIND = {0: "constant", 1: "instance", 2: "witness", 3: "add", 4: "mul", 5: "addc", 6: "mulc", 7: "and", 8: "xor", 9: "not"}
BLS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def oracle_run(msgs):
    tb = ev.TracingBackend()
    o = ev.Evaluator.from_messages(msgs, tb)
    return tb, o, o.get_violations()


def compare_prefix(b, tb, complete):
    """our recorded program must start with the oracle's trace (the oracle stops at its first error)"""
    kinds, a, bb = b.program()
    alias, nxt = {}, 0
    p = tb.m
    for i, (k, ops, val) in enumerate(tb.trace):
        if k == "copy":
            alias[i] = alias[ops[0]]
            continue
        h = nxt
        nxt += 1
        alias[i] = h
        assert h < len(kinds), "oracle recorded more values than the flattener"
        assert KIND[int(kinds[h])] == k, (i, k, KIND[int(kinds[h])])
        if k in ("add", "mul", "and", "xor"):
            assert (int(a[h]), int(bb[h])) == (alias[ops[0]], alias[ops[1]]), (i, k)
        elif k in ("addc", "mulc", "not"):
            assert int(a[h]) == alias[ops[0]], (i, k)
        if k == "constant":
            assert b.const_value(int(bb[h])) == val % p
    for s, (_, ssa, _) in enumerate(tb.asserts):
        assert b.assert_value(s) == alias[ssa]
    if complete:
        st = b.stats()
        assert nxt == st["n_values"] and st["n_asserts"] == len(tb.asserts)
        oc = tb.counts()
        for k, v in st["callbacks"].items():
            assert v == oc.get(k, 0), (k, v, oc.get(k, 0))


CASES = [(seed, 101, False) for seed in range(40)] + [(seed, 2, True) for seed in range(100, 125)] + \
        [(seed, (1 << 61) - 1, False) for seed in range(200, 206)] + [(300, BLS, False), (301, BLS, False)]


@pytest.mark.parametrize("seed,p,boolean", CASES)
def test_flattener_matches_oracle_on_random_structured_programs(seed, p, boolean):
    msgs = Gen(seed, p, boolean).statement()
    assert F.read_messages(F.write_messages(msgs)) == msgs
    tb, o, viol = oracle_run(msgs)
    z = zkb()
    b = z.GpuBackend(-1)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_buffers([F.write_messages(msgs)]))
    assert b.pending_error() is None       # programs are structurally valid
    compare_prefix(b, tb, complete=(viol == []))


@pytest.mark.gpu
@pytest.mark.parametrize("seed,p,boolean", CASES)
def test_gpu_matches_oracle_on_random_structured_programs(seed, p, boolean):
    msgs = Gen(seed, p, boolean).statement()
    tb, o, viol = oracle_run(msgs)
    z = zkb()
    b = z.GpuBackend(0)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_buffers([F.write_messages(msgs)]))
    if b.stats()["nlimb"]:
        b.finalize(True)
    assert e.get_violations() == viol
    if viol == [] and b.stats()["n_values"]:
        vals = [v for (k, _, v) in tb.trace if k != "copy"]
        assert b.read_values(0, list(range(len(vals))), 40) == vals
        for wid, w in o.values.items():
            assert e.get(wid) == w[1]
    elif b.stats()["n_values"]:
        # values recorded BEFORE the first failing assertion are defined in the reference: compare those
        vals = [v for (k, _, v) in tb.trace if k != "copy"]
        assert b.read_values(0, list(range(len(vals))), 40) == vals
