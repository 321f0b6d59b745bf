"""Pin the oracle (oracle/*.py) against every golden vector the reference's own
tests hold for the evaluation path.  CPU only."""
import glob
import json
import os
import re

import pytest

from oracle import evaluator as ev
from oracle import fixtures as fx
from oracle import ir
from oracle import sieve_fbs as F

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
GOLDILOCKS = (1 << 64) - (1 << 32) + 1
BLS12_381_FR = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def test_exponentiation_kats():
    # consumers/evaluator.rs:949-984
    b = ev.PlaintextBackend()
    for m, base, e, expected in [
        (16249742125730185677094195492597105093, 2, 2206000150907221872269901214599500635,
         5834907326474057072663503101785122138),
        (101, 42, 100, 1),
    ]:
        b.set_field(ir.le_bytes(m), 1, False)
        assert ev.exp(b, base, e, m, False) == expected


def test_example_true():
    # consumers/evaluator.rs:986-1004
    assert ev.evaluate([fx.example_instance(), fx.example_witness(), fx.example_relation()]) == []


def test_example_wrong_result():
    # consumers/evaluator.rs:1082-1104
    v = ev.evaluate([fx.example_instance(), fx.example_witness_incorrect(), fx.example_relation()])
    assert v == ["Wire_9 (may be weighted) should be 0, while it is not"]


def test_example_all_wires_freed_and_op_counts():
    # op counts derived in SURVEY.md §8c (283 callbacks incl. 6 assert_zero)
    tb = ev.TracingBackend()
    e = ev.Evaluator.from_messages([fx.example_instance(), fx.example_witness(), fx.example_relation()], tb)
    assert e.get_violations() == []
    assert e.values == {}
    c = tb.counts()
    assert c == {"copy": 147, "mul": 57, "add": 43, "constant": 11, "witness": 6, "instance": 6,
                 "assert_zero": 6, "mulc": 5, "addc": 2}
    # Stats cross-check (consumers/stats.rs:302-329): 21 mul gates in the unrolled IR =
    # 57 backend muls minus 2 cases * 18 exponentiation muls
    assert c["mul"] - 2 * 18 == 21


def test_example_incorrect_stops_after_31_callbacks():
    tb = ev.TracingBackend()
    e = ev.Evaluator.from_messages(
        [fx.example_instance(), fx.example_witness_incorrect(), fx.example_relation()], tb)
    assert len(tb.trace) + len(tb.asserts) == 31
    assert e.get_violations() == ["Wire_9 (may be weighted) should be 0, while it is not"]


@pytest.mark.parametrize("p,muls", [(GOLDILOCKS, 229), (BLS12_381_FR, 813)])
def test_example_other_fields(p, muls):
    h = fx.example_header(p)
    tb = ev.TracingBackend()
    e = ev.Evaluator.from_messages([fx.example_instance(h), fx.example_witness(h), fx.example_relation(h)], tb)
    assert e.get_violations() == []
    assert tb.counts()["mul"] == muls


def test_boolean_example():
    # cli.rs:602-624 (bool-example -> valid-eval-metrics), boolean_examples.rs
    msgs = [fx.boolean_example_instance(), fx.boolean_example_witness(), fx.boolean_example_relation()]
    tb = ev.TracingBackend()
    e = ev.Evaluator.from_messages(msgs, tb)
    assert e.get_violations() == []
    assert tb.counts() == {"copy": 49, "xor": 21, "and": 19, "not": 14, "instance": 10, "witness": 5,
                           "assert_zero": 4, "constant": 3}
    v = ev.evaluate([fx.boolean_example_instance(), fx.boolean_example_witness_incorrect(),
                     fx.boolean_example_relation()])
    assert v == ["Wire_22 (may be weighted) should be 0, while it is not"]


@pytest.mark.parametrize("build", [fx.builder_with_function, fx.builder_with_several_functions,
                                   fx.builder_switch, fx.builder_switch_nested_in_function])
def test_builder_circuits_true(build):
    # producers/builder.rs:726-1175
    assert ev.evaluate(build()) == []


def test_r1cs_example_wire_values_and_counts():
    # producers/from_r1cs.rs:178-286
    msgs, _ = fx.r1cs_to_gates(*fx.zkif_example())
    e = ev.Evaluator.from_messages(msgs, ev.PlaintextBackend())
    assert [e.get(i) for i in range(7)] == [1, 100, 3, 4, 25, 9, 16]
    assert e.get_violations() == []
    rel = msgs[-1]
    kinds = [g[0] for g in rel.gates]
    assert kinds.count("Constant") == 12 and kinds.count("Mul") == 15 and kinds.count("Add") == 4
    assert kinds.count("AssertZero") == 3 and kinds.count("Instance") == 3 and kinds.count("Witness") == 2
    assert rel.header.field_characteristic == bytes([101])


def _flatc_json(path):
    txt = open(path).read()
    txt = re.sub(r"(?m)^(\s*)([A-Za-z_]+):", r'\1"\2":', txt)
    return json.loads(txt)


def test_reference_binary_fixtures_parse_like_their_json_twins():
    # rust/examples/*.sieve with flatc-JSON twins
    inst = F.read_message(open(os.path.join(GOLDEN, "000_instance.sieve"), "rb").read())
    j = _flatc_json(os.path.join(GOLDEN, "000_instance.json"))
    assert [list(v) for v in inst.common_inputs] == [x["value"] for x in j["message"]["common_inputs"]]
    assert list(inst.header.field_characteristic) == j["message"]["header"]["field_characteristic"]["value"]
    wit = F.read_message(open(os.path.join(GOLDEN, "001_witness.sieve"), "rb").read())
    j = _flatc_json(os.path.join(GOLDEN, "001_witness.json"))
    assert [list(v) for v in wit.short_witness] == [x["value"] for x in j["message"]["short_witness"]]
    rel = F.read_message(open(os.path.join(GOLDEN, "002_relation.sieve"), "rb").read())
    j = _flatc_json(os.path.join(GOLDEN, "002_relation.json"))["message"]
    assert ir.create_gateset_string(rel.gate_mask) == j["gateset"]
    assert ir.create_feature_string(rel.feat_mask) == j["features"]
    assert [f.name for f in rel.functions] == [f["name"] for f in j["functions"]]
    assert ["Gate" + g[0] for g in rel.gates] == [d["directive_type"] for d in j["directives"]]
    sw = j["directives"][1]["directive"]
    assert rel.gates[1][3] == [bytes(c["value"]) for c in sw["cases"]]
    assert [len(b[4]) for b in rel.gates[1][4]] == [len(b["invocation"]["subcircuit"]) for b in sw["branches"]]


def test_reference_binary_fixtures_evaluate_true():
    bufs = [open(p, "rb").read() for p in sorted(glob.glob(os.path.join(GOLDEN, "*.sieve")))]
    msgs = [m for b in bufs for m in F.read_messages(b)]
    tb = ev.TracingBackend()
    e = ev.Evaluator.from_messages(msgs, tb)
    assert e.get_violations() == []
    assert len(tb.trace) + len(tb.asserts) == 210   # SURVEY.md §8c derived count


@pytest.mark.parametrize("make", [
    lambda: [fx.example_instance(), fx.example_witness(), fx.example_relation()],
    lambda: [fx.boolean_example_instance(), fx.boolean_example_witness(), fx.boolean_example_relation()],
    fx.builder_switch, fx.builder_switch_nested_in_function,
])
def test_writer_reader_round_trip(make):
    # mirrors producers/examples.rs:243-259
    msgs = make()
    buf = F.write_messages(msgs)
    assert F.read_messages(buf) == msgs


def test_wirelist_and_iterexpr_semantics():
    # structs/wire.rs:205-219
    assert ir.expand_wirelist([ir.WireRange(0, 2), ir.Wire(5)]) == [0, 1, 2, 5]
    with pytest.raises(ValueError):
        ir.expand_wirelist([ir.WireRange(0, 1), ir.WireRange(2, 2), ir.Wire(5)])
    with pytest.raises(ValueError):
        ir.expand_wirelist([ir.WireRange(4, 2)])
    with pytest.raises(ir.OraclePanic):
        ir.evaluate_iterexpr_list([("Single", ("Name", "j"))], {"i": 1})
    assert ir.evaluate_iterexpr_list([("Range", ("Const", 3), ("Const", 2))], {}) == []


def test_unreduced_inputs_are_kept_raw():
    # SURVEY.md §8a trap 1: a witness equal to p fails AssertZero although it is 0 mod p
    h = fx.example_header()
    rel = ir.Relation(h, ir.ARITH, ir.SIMPLE, [], [("Witness", 0), ("AssertZero", 0)])
    v = ev.evaluate([ir.Witness(h, [bytes([101])]), rel])
    assert v == ["Wire_0 (may be weighted) should be 0, while it is not"]
    rel2 = ir.Relation(h, ir.ARITH, ir.SIMPLE, [], [("Witness", 0), ("AddConstant", 1, 0, bytes([0])), ("AssertZero", 1)])
    assert ev.evaluate([ir.Witness(h, [bytes([101])]), rel2]) == []


def test_no_gate_and_missing_inputs():
    h = fx.example_header()
    assert ev.evaluate([ir.Relation(h, ir.ARITH, ir.SIMPLE, [], [])]) == ["Did not receive any gate to verify."]
    assert ev.evaluate([ir.Relation(h, ir.ARITH, ir.SIMPLE, [], [("Instance", 0)])]) == \
        ["Not enough instance to consume"]
    with pytest.raises(ir.OraclePanic):
        ev.evaluate([ir.Relation(h, ir.ARITH, ir.SIMPLE, [], [("Witness", 0)])])
