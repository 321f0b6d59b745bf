"""GPU parity of the Evaluator path (structured gates: For / Switch / Call / AnonCall) against the
oracle: violation strings, every recorded value, Evaluator::get."""
import glob
import os

import numpy as np
import pytest

from oracle import evaluator as ev
from oracle import fixtures as fx
from oracle import ir
from oracle import sieve_fbs as F
from tests.test_host_evaluator import GOLDEN, STATEMENTS
from tests.util import zkb

pytestmark = pytest.mark.gpu


def run_gpu(msgs, keep_all=True):
    z = zkb()
    b = z.GpuBackend(0)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_buffers([F.write_messages(msgs)]))
    if keep_all and b.stats()["nlimb"]:
        b.finalize(True)
    return z, b, e, e.get_violations()


def check_all_values(b, msgs):
    tb = ev.TracingBackend()
    o = ev.Evaluator.from_messages(msgs, tb)
    vals = [v for (k, _, v) in tb.trace if k != "copy"]
    n = b.stats()["n_values"]
    assert n == len(vals)
    got = b.read_values(0, list(range(n)), 40)
    assert got == vals
    return o


@pytest.mark.parametrize("name", list(STATEMENTS))
def test_statement_true_and_every_value(name):
    msgs = STATEMENTS[name]()
    z, b, e, viol = run_gpu(msgs)
    assert viol == ev.evaluate(msgs) == []
    o = check_all_values(b, msgs)
    for wid, val in o.values.items():            # Evaluator::get on the live top-scope wires
        assert e.get(wid) == val[1]


def test_example_wrong_witness_exact_text():
    # consumers/evaluator.rs:1082-1104
    msgs = [fx.example_instance(), fx.example_witness_incorrect(), fx.example_relation()]
    _, _, _, viol = run_gpu(msgs)
    assert viol == ["Wire_9 (may be weighted) should be 0, while it is not"]


def test_boolean_wrong_witness():
    msgs = [fx.boolean_example_instance(), fx.boolean_example_witness_incorrect(), fx.boolean_example_relation()]
    _, _, _, viol = run_gpu(msgs)
    assert viol == ev.evaluate(msgs) == ["Wire_22 (may be weighted) should be 0, while it is not"]


def test_r1cs_example_wire_values():
    # producers/from_r1cs.rs:209-217
    msgs, _ = fx.r1cs_to_gates(*fx.zkif_example())
    _, _, e, viol = run_gpu(msgs, keep_all=False)
    assert viol == []
    assert [e.get(i) for i in range(7)] == [1, 100, 3, 4, 25, 9, 16]


def test_reference_binary_fixtures_true():
    z = zkb()
    e = z.Evaluator.from_messages(z.Source.from_directory(GOLDEN), device=0)
    assert e.get_violations() == []


def test_no_gate_and_latched_errors():
    h = fx.example_header()
    _, _, _, viol = run_gpu([ir.Relation(h, ir.ARITH, ir.SIMPLE, [], [])])
    assert viol == ["Did not receive any gate to verify."]
    # a failing assertion earlier in program order wins over a later structural error
    rel = ir.Relation(h, ir.ARITH, ir.SIMPLE, [], [("Constant", 0, b"\x05"), ("AssertZero", 0), ("Add", 3, 1, 2)])
    _, _, _, viol = run_gpu([rel])
    assert viol == ev.evaluate([rel]) == ["Wire_0 (may be weighted) should be 0, while it is not"]
    rel = ir.Relation(h, ir.ARITH, ir.SIMPLE, [], [("Constant", 0, b"\x00"), ("AssertZero", 0), ("Add", 3, 1, 2)])
    _, _, _, viol = run_gpu([rel])
    assert viol == ev.evaluate([rel]) == ["No value given for wire_1"]


def test_is_boolean_needs_all_three_gates():
    # SURVEY.md 8a trap 3: "@xor,@and" over p = 2 runs Switch weights with ARITHMETIC ops
    h = fx.boolean_header()
    sw = ("Switch", 0, [ir.Wire(2)], [b"\x01", b"\x00"], [
        ("AbstractAnonCall", [ir.Wire(1)], 0, 0, [("Xor", 0, 1, 1)]),
        ("AbstractAnonCall", [ir.Wire(1)], 0, 0, [("And", 0, 1, 1)])])
    for mask in (ir.XOR | ir.AND, ir.BOOL):
        rel = ir.Relation(h, mask, ir.SWITCH, [], [("Witness", 0), ("Witness", 1), sw, ("AssertZero", 2)])
        for w in ([1, 1], [0, 1], [1, 0]):
            msgs = [ir.Witness(h, [bytes([x]) for x in w]), rel]
            z, b, e, viol = run_gpu(msgs)
            assert viol == ev.evaluate(msgs)
            if not viol:
                check_all_values(b, msgs)


def test_nested_structures_random(seed=5):
    """nested For / Switch / Call / AnonCall with Free and wire re-use, field 101 and Goldilocks"""
    rng = np.random.default_rng(seed)
    for p in (101, (1 << 64) - (1 << 32) + 1):
        h = ir.Header(ir.le_bytes(p))
        m1 = ir.le_bytes(p - 1)
        sq = ir.Function("sq_add", 1, 2, 0, 1, [("Mul", 3, 1, 1), ("Witness", 4), ("Add", 5, 3, 2), ("Add", 0, 5, 4)])
        inner = ir.Function("inner_sw", 2, 2, 1, 1, [
            ("Switch", 2, [ir.WireRange(0, 1)], [b"\x01", b"\x02", b"\x07"], [
                ("AbstractGateCall", "two", [ir.Wire(2), ir.Wire(3)]),
                ("AbstractAnonCall", [ir.Wire(3)], 1, 1, [("Instance", 5), ("Witness", 6), ("Mul", 0, 2, 5), ("Add", 1, 6, 2),
                                                          ("AssertZero", 6)]),
                ("AbstractGateCall", "two", [ir.Wire(3), ir.Wire(2)]),
            ])])
        two = ir.Function("two", 2, 2, 1, 0, [("Instance", 4), ("Add", 0, 2, 4), ("MulConstant", 1, 3, m1)])
        gates = [("Witness", 0), ("Witness", 1), ("Constant", 2, b"\x03")]
        I, C = (lambda n: ("Name", n)), (lambda v: ("Const", v))
        gates.append(("For", "i", 0, 5, [ir.WireRange(3, 8)], ("IterExprCall", "sq_add",
                      [("Single", ("Add", I("i"), C(3)))], [("Single", ("Add", I("i"), C(1))), ("Single", ("Add", I("i"), C(2)))])))
        gates.append(("For", "i", 0, 2, [ir.WireRange(9, 14)], ("IterExprAnonCall",
                      [("Range", ("Add", ("Mul", I("i"), C(2)), C(9)), ("Add", ("Mul", I("i"), C(2)), C(10)))],
                      [("Single", ("Add", I("i"), C(3))), ("Single", C(0))], 1, 1,
                      [("Call", "inner_sw", [ir.WireRange(0, 1)], [ir.Wire(3), ir.Wire(2)])])))
        gates += [("Free", 3, 8), ("Add", 3, 9, 14), ("AddConstant", 4, 3, b"\x09"), ("Free", 9, 13)]
        rel = ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [sq, two, inner], gates)
        for trial in range(4):
            wit = [ir.le_bytes(int(rng.integers(0, 3))) for _ in range(2)] + [ir.le_bytes(int.from_bytes(rng.bytes(16), "little") % p) for _ in range(12)]
            wit[1] = ir.le_bytes(int(rng.integers(1, 3)))
            inst = [ir.le_bytes(int.from_bytes(rng.bytes(16), "little") % p) for _ in range(6)]
            msgs = [ir.Instance(h, inst), ir.Witness(h, wit), rel]
            expected = ev.evaluate(msgs)
            z, b, e, viol = run_gpu(msgs)
            assert viol == expected, (p, trial)
            if not expected:
                o = check_all_values(b, msgs)
                for wid, val in o.values.items():
                    assert e.get(wid) == val[1]


def test_batch_through_evaluator_program():
    """record once through the Evaluator, then check a whole batch of witnesses with the backend"""
    z = zkb()
    msgs = [fx.example_instance(), fx.example_witness(), fx.example_relation()]
    b = z.GpuBackend(0)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_buffers([F.write_messages(msgs)]))
    assert e.get_violations() == []
    good = [3, 4, 0, 17711 % 101]
    bad = [3, 5, 1, 40]
    wit = np.zeros((6, 4, 4), dtype=np.uint8)
    for j in range(6):
        wit[j, :, 0] = bad if j in (1, 4) else good
    inst = np.zeros((3, 4), dtype=np.uint8)
    inst[:, 0] = [25, 0, 1]
    v = b.evaluate(inst, wit, 6)
    assert [int(x) for x in v["ok"]] == [1, 0, 1, 1, 0, 1]
    assert b.assert_wire(int(v[1]["first_fail_seq"])) == 9


def test_failing_assertion_before_a_missing_witness_value_is_the_violation():
    """The reference evaluates as it goes: an assertion that fails before the Witness gate that finds its queue empty
    latches its error (evaluator.rs:213-221, 357-362) and the panic of PlaintextBackend::witness(None) (:944-946) is never
    reached.  The deferred backend records first, so it resolves the recorded prefix when the panic condition is met."""
    z = zkb()
    h = fx.example_header()
    rel = ir.Relation(h, ir.ARITH, ir.SIMPLE, [], [
        ("Witness", 0), ("Instance", 1), ("Add", 2, 0, 1), ("AssertZero", 2),       # 3 + 25 != 0 (mod 101)
        ("Witness", 3), ("Witness", 4), ("Mul", 5, 3, 4), ("AssertZero", 5)])
    wit_short = ir.Witness(h, [ir.literal32(3)])
    msgs = [fx.example_instance(), wit_short, rel]
    assert ev.evaluate(msgs) == ["Wire_2 (may be weighted) should be 0, while it is not"]
    e = z.Evaluator(z.GpuBackend(0))
    e.ingest_source(z.Source.from_buffers([F.write_messages(msgs)]))
    assert e.get_violations() == ["Wire_2 (may be weighted) should be 0, while it is not"]
    # with a witness that satisfies the first assertion the missing value is reached: the reference panics
    wit_ok = ir.Witness(h, [ir.literal32(101 - 25)])
    e = z.Evaluator(z.GpuBackend(0))
    with pytest.raises(z.ZkbError) as err:
        e.ingest_source(z.Source.from_buffers([F.write_messages([fx.example_instance(), wit_ok, rel])]))
    assert err.value.code == z.ZKB_E_FATAL and "Missing witness value" in str(err.value)
