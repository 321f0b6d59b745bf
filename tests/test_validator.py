"""Validator (rust/src/consumers/validator.rs): the oracle restatement is pinned on the reference's own tests,
the C++ validator behind the C ABI must then produce the same violation list, text for text, on valid statements,
on the reference's violation cases and on randomly mutated structured programs."""
import copy

import numpy as np
import pytest

from oracle import evaluator as ev
from oracle import fixtures as fx
from oracle import flattening as fl
from oracle import ir
from oracle import sieve_fbs as F
from oracle import validator as ov
from tests.gen_programs import Gen
from tests.test_host_evaluator import STATEMENTS
from tests.util import zkb


def ours(msgs, as_prover=True, per_message=False):
    z = zkb()
    v = z.Validator(as_prover)
    if per_message:
        for m in msgs:
            v.ingest_message(F.write_message(m))
    else:
        v.ingest_source(z.Source.from_buffers([F.write_messages(msgs)]))
    return v.get_violations()


# ---- the oracle against the reference's golden tests ---------------------------------------------------
def test_oracle_validator_valid_example():                       # validator.rs:831-850
    msgs = [fx.example_instance(), fx.example_witness(), fx.example_relation()]
    assert ov.validate(msgs, as_prover=True) == []


def test_oracle_validator_as_verifier():                         # validator.rs:852-870
    assert ov.validate([fx.example_instance(), fx.example_relation()], as_prover=False) == []


def reference_violation_case():                                  # validator.rs:872-902
    instance, witness, relation = fx.example_instance(), fx.example_witness(), fx.example_relation()
    instance.common_inputs[0] = bytes(instance.header.field_characteristic)
    witness.short_witness.pop()
    relation.header = copy.deepcopy(relation.header)
    relation.header.field_characteristic = bytes([10])
    return [instance, witness, relation]


EXPECTED_VIOLATIONS = [
    "The instance value [101, 0, 0, 0] cannot be represented in the field specified in Header (101 >= 101).",
    "The field_characteristic field is not consistent across headers.",
    "Not enough Witness value to consume.",
]


def reference_free_case():                                       # validator.rs:904-930
    relation = fx.example_relation()
    relation.gates = list(relation.gates) + [("Free", 1, 2), ("Free", 4, None)]
    return [fx.example_instance(), fx.example_witness(), relation]


EXPECTED_FREE = [
    "The wire 1 is used but was not assigned a value, or has been freed already.",
    "The wire 2 is used but was not assigned a value, or has been freed already.",
    "The wire 4 is used but was not assigned a value, or has been freed already.",
]


def test_oracle_validator_violations():
    assert ov.validate(reference_violation_case()) == EXPECTED_VIOLATIONS


def test_oracle_validator_free_violations():
    assert ov.validate(reference_free_case()) == EXPECTED_FREE


def test_oracle_validates_flattening():                          # flattening.rs:200-225
    f = fl.IRFlattener()
    ev.Evaluator.from_messages([fx.example_instance(), fx.example_witness(), fx.example_relation()], f)
    assert ov.validate(f.finish()) == []


def test_oracle_is_probably_prime():                             # value.rs:58-65
    assert not ov.is_probably_prime(bytes([187]))
    assert ov.is_probably_prime(bytes([101]))


@pytest.mark.parametrize("name", list(STATEMENTS))
def test_oracle_validator_accepts_the_reference_statements(name):   # cli.rs:602-624 validates what it produces
    assert ov.validate(STATEMENTS[name]()) == []


# ---- the product against the reference's tests and against the oracle ----------------------------------
def test_reference_cases_through_the_c_abi():
    assert ours([fx.example_instance(), fx.example_witness(), fx.example_relation()]) == []
    assert ours([fx.example_instance(), fx.example_relation()], as_prover=False) == []
    assert ours(reference_violation_case()) == EXPECTED_VIOLATIONS
    assert ours(reference_free_case()) == EXPECTED_FREE
    assert ours(reference_free_case(), per_message=True) == EXPECTED_FREE


@pytest.mark.parametrize("name", list(STATEMENTS))
@pytest.mark.parametrize("as_prover", [True, False])
def test_validator_matches_oracle_on_statements(name, as_prover):
    msgs = STATEMENTS[name]()
    assert ours(msgs, as_prover) == ov.validate(msgs, as_prover)


def mutate(msgs, rng):
    """damage a statement in ways the Validator has a check for"""
    msgs = copy.deepcopy(msgs)
    rel = [m for m in msgs if isinstance(m, ir.Relation)][-1]
    kind = int(rng.integers(0, 12))

    def walk(gates, out):
        for i, g in enumerate(gates):
            out.append((gates, i))
            if g[0] == "AnonCall":
                walk(g[5], out)
            elif g[0] == "For" and g[5][0] == "IterExprAnonCall":
                walk(g[5][5], out)
            elif g[0] == "Switch":
                for br in g[4]:
                    if br[0] == "AbstractAnonCall":
                        walk(br[4], out)
    sites = []
    walk(rel.gates, sites)
    for f in rel.functions:
        walk(f.body, sites)
    gates, i = sites[int(rng.integers(0, len(sites)))]
    g = gates[i]
    if kind == 0:
        gates.pop(i)                                             # a definition disappears: use-before-set / missing output
    elif kind == 1:
        gates.insert(i, g)                                       # duplicate: SSA violation / double free
    elif kind == 2 and g[0] in ("Add", "Mul", "And", "Xor"):
        gates[i] = (g[0], g[1], g[2] + 1000, g[3])               # undefined operand
    elif kind == 3:
        rel.gate_mask = ir.BOOL if rel.gate_mask == ir.ARITH else ir.ARITH   # gates not allowed in this gateset
    elif kind == 4:
        rel.feat_mask = ir.SIMPLE                                # features not allowed
    elif kind == 5:
        for m in msgs:
            if isinstance(m, ir.Witness) and m.short_witness:
                m.short_witness.pop()                            # not enough witness values
    elif kind == 6:
        for m in msgs:
            if isinstance(m, ir.Instance):
                m.common_inputs.append(b"\x01")                  # too many instance values
    elif kind == 7:
        rel.header = copy.deepcopy(rel.header)
        rel.header.version = "1.0"                               # inconsistent + (if first) malformed version
    elif kind == 8:
        p = int.from_bytes(rel.header.field_characteristic, "little")
        gates.insert(i, ("Constant", 99999, ir.le_bytes(p + 3)))  # constant not in the field
    elif kind == 9:
        gates.insert(i, ("Free", 5, 3))                          # Free with last <= first
    elif kind == 10 and rel.functions:
        rel.functions.append(copy.deepcopy(rel.functions[0]))    # duplicate function name
    elif kind == 11:
        gates.insert(i, ("Call", "no.such::function", [ir.Wire(77777)], []))
    return msgs


@pytest.mark.parametrize("seed", range(40))
def test_validator_matches_oracle_on_mutated_structured_programs(seed):
    rng = np.random.default_rng(1000 + seed)
    boolean = seed % 4 == 3
    p = 2 if boolean else [101, (1 << 61) - 1, 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001][seed % 3]
    g = Gen(seed, p, boolean=boolean)
    msgs = g.statement() if hasattr(g, "statement") else None
    if msgs is None:
        pytest.skip("generator has no statement()")
    for as_prover in (True, False):
        assert ours(msgs, as_prover) == ov.validate(msgs, as_prover)
    for _ in range(4):
        bad = mutate(msgs, rng)
        want = ov.validate(bad, True)
        assert ours(bad, True) == want
        assert ours(bad, False) == ov.validate(bad, False)


def test_header_checks():
    h = ir.Header(bytes([187]), "1.0", 2)                         # composite, bad version, degree 2
    rel = ir.Relation(h, ir.ARITH | ir.BOOL, ir.SIMPLE, [], [("Constant", 0, b""), ("AssertZero", 0)])
    assert ov.validate([rel]) == ["The field_characteristic should be a prime.", "field_degree must be = 1",
                                  "The profile version should match the following format <major>.<minor>.<patch>.",
                                  "Cannot mix arithmetic and boolean gates",
                                  "With boolean profile the field characteristic can only be 2.",
                                  "The Gate::Constant constant is empty."]
    # on the wire the mixed mask cannot exist: create_gateset_string writes "arithmetic" (relation.rs:183)
    rel = ir.Relation(h, ir.BOOL, ir.SIMPLE, [], [("Constant", 0, b""), ("AssertZero", 0), ("Add", 1, 0, 0)])
    want = ov.validate([rel])
    assert "With boolean profile the field characteristic can only be 2." in want
    assert "The gate @add is not allowed in this circuit." in want
    assert ours([rel]) == want
    # `.` in the version pattern is any character, digits included
    for version, ok in [("1.0.0", True), ("1x2y3", True), ("12345", True), ("1234", False), ("1.0.", False), (" 1.0.0 ", True),
                        ("1.0.0\n", True), ("a.0.0", False)]:
        rel = ir.Relation(ir.Header(bytes([101]), version, 1), ir.ARITH, ir.SIMPLE, [], [])
        want = ov.validate([rel])
        assert (want == []) == ok, version
        assert ours([rel]) == want
    # names: function names are trimmed before matching, iterator names are not
    for name, ok in [("f", True), ("com.example::mul", True), ("a:b", False), ("9a", False), ("a..b", False), ("_x9.y_", True),
                     (" f ", True), ("a::", False)]:
        fn = ir.Function(name, 0, 0, 0, 0, [])
        rel = ir.Relation(fx.example_header(), ir.ARITH, ir.FOR_FUNCTION_SWITCH, [fn], [])
        want = ov.validate([rel])
        assert (want == []) == ok, name
        assert ours([rel]) == want
        loop = ("For", name, 0, 0, [], ("IterExprAnonCall", [], [], 0, 0, []))
        rel = ir.Relation(fx.example_header(), ir.ARITH, ir.FOR_FUNCTION_SWITCH, [], [loop])
        want = ov.validate([rel])
        assert (want == []) == (ok and name.strip() == name), name
        assert ours([rel]) == want


def test_switch_checks():
    h = fx.example_header()
    sw = ("Switch", 0, [ir.Wire(1)], [b"\x03", b"\x03", bytes([200])],
          [("AbstractAnonCall", [ir.Wire(0)], 0, 1, [("Witness", 0)]),
           ("AbstractGateCall", "nope", [ir.Wire(0)])])
    rel = ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [], [("Constant", 0, b"\x03"), sw, ("Switch", 0, [ir.Wire(5)], [], [])])
    msgs = [ir.Witness(h, [b"\x01"]), rel]
    want = ov.validate(msgs)
    assert "Gate::Switch: The number of cases value does not match the number of branches." in want
    assert "Gate::Switch: The cases values contain duplicates." in want
    assert "The Gate::Switch case value: 200 cannot be represented in the field specified in Header (200 >= 101)." in want
    assert "Switch: no case given while non-empty list of output wires." in want
    assert "Unknown Function gate nope" in want
    assert ours(msgs) == want
    assert ours(msgs, False) == ov.validate(msgs, False)


def test_validator_panics_and_limits():
    z = zkb()
    h = fx.example_header()
    loop = ("For", "i", 0, 1, [], ("IterExprAnonCall", [("Single", ("Name", "j"))], [], 0, 0, [("Constant", 0, b"\x01")]))
    rel = ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [], [loop])
    with pytest.raises(ir.OraclePanic):
        ov.validate([rel])
    v = z.Validator(True)
    with pytest.raises(z.ZkbError) as e:
        v.ingest_message(F.write_message(rel))
    assert e.value.code == z.ZKB_E_FATAL and str(e.value) == "Unknown iterator name j"
    hostile = ("For", "i", 0, 1 << 60, [], ("IterExprAnonCall", [], [], 0, 0, []))
    v = z.Validator(True)
    v.set_limits(1 << 16)
    with pytest.raises(z.ZkbError) as e:
        v.ingest_message(F.write_message(ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [], [hostile])))
    assert str(e.value) == "zkb: resource limit exceeded (max_steps)"
    v = z.Validator(True)
    with pytest.raises(z.ZkbError) as e:
        v.ingest_message(b"\x10\x00\x00\x00" + bytes(16))
    assert e.value.code == z.ZKB_E_FORMAT
