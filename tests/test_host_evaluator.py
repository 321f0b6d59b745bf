"""Host preparation parity (no GPU): the C++ Evaluator/reader behind the C ABI records, op for op,
the program the oracle's Evaluator asks of a ZKBackend (oracle/evaluator.py TracingBackend)."""
import glob
import os

import numpy as np
import pytest

from oracle import evaluator as ev
from oracle import fixtures as fx
from oracle import ir
from oracle import sieve_fbs as F
from tests.util import zkb

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
KIND = {0: "constant", 1: "instance", 2: "witness", 3: "add", 4: "mul", 5: "addc", 6: "mulc", 7: "and", 8: "xor", 9: "not"}

STATEMENTS = {
    "example": lambda: [fx.example_instance(), fx.example_witness(), fx.example_relation()],
    "example_goldilocks": lambda: (lambda h: [fx.example_instance(h), fx.example_witness(h), fx.example_relation(h)])(
        fx.example_header((1 << 64) - (1 << 32) + 1)),
    "example_bls": lambda: (lambda h: [fx.example_instance(h), fx.example_witness(h), fx.example_relation(h)])(
        fx.example_header(0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001)),
    "boolean": lambda: [fx.boolean_example_instance(), fx.boolean_example_witness(), fx.boolean_example_relation()],
    "builder_function": fx.builder_with_function,
    "builder_several_functions": fx.builder_with_several_functions,
    "builder_switch": fx.builder_switch,
    "builder_switch_nested": fx.builder_switch_nested_in_function,
    "r1cs_example": lambda: fx.r1cs_to_gates(*fx.zkif_example())[0],
}


def record(msgs_bytes):
    z = zkb()
    b = z.GpuBackend(-1)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_buffers([msgs_bytes]))
    return z, b, e


def compare_with_oracle_trace(b, msgs):
    tb = ev.TracingBackend()
    o = ev.Evaluator.from_messages(msgs, tb)
    st = b.stats()
    # same number of callbacks of every kind (copies are aliases on our side but still counted)
    oc = tb.counts()
    for k, v in st["callbacks"].items():
        assert v == oc.get(k, 0), (k, v, oc.get(k, 0))
    kinds, a, bb = b.program()
    # oracle ssa id -> our value handle (copy = alias of its operand)
    alias = {}
    nxt = 0
    inst_pos = {}
    p = tb.m
    for i, (k, ops, val) in enumerate(tb.trace):
        if k == "copy":
            alias[i] = alias[ops[0]]
            continue
        h = nxt
        nxt += 1
        alias[i] = h
        assert KIND[int(kinds[h])] == k, (i, k, KIND[int(kinds[h])])
        if k in ("add", "mul", "and", "xor"):
            assert (int(a[h]), int(bb[h])) == (alias[ops[0]], alias[ops[1]]), (i, k)
        elif k in ("addc", "mulc", "not"):
            assert int(a[h]) == alias[ops[0]], (i, k)
        if k == "constant":
            assert b.const_value(int(bb[h])) == val % p
    assert nxt == st["n_values"]
    # assertions: same count, same asserted values, in order
    assert st["n_asserts"] == len(tb.asserts)
    for s, (_, ssa, _) in enumerate(tb.asserts):
        assert b.assert_value(s) == alias[ssa]
    return o


@pytest.mark.parametrize("name", list(STATEMENTS))
def test_flattening_matches_oracle_trace(name):
    msgs = STATEMENTS[name]()
    z, b, e = record(F.write_messages(msgs))
    compare_with_oracle_trace(b, msgs)
    # no GPU in this process: evaluation must fail loudly, never fall back
    with pytest.raises(z.ZkbError) as err:
        e.get_violations()
    assert err.value.code == z.ZKB_E_CUDA


def test_reference_binary_fixtures_through_source_paths():
    z = zkb()
    b = z.GpuBackend(-1)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_directory(GOLDEN))   # *.sieve, ordered instance < witness < relation
    bufs = [open(p, "rb").read() for p in sorted(glob.glob(os.path.join(GOLDEN, "*.sieve")))]
    msgs = [m for bf in bufs for m in F.read_messages(bf)]
    compare_with_oracle_trace(b, msgs)
    st = b.stats()
    assert sum(st["callbacks"].values()) == 210


def test_source_orders_files_like_the_reference(tmp_path):
    # source.rs:69-89: lexical sort, then stable sort instance(0) < witness(1) < relation(3) < other(4)
    msgs = STATEMENTS["example"]()
    (tmp_path / "zz_relation.sieve").write_bytes(F.write_message(msgs[2]))
    (tmp_path / "b_witness.sieve").write_bytes(F.write_message(msgs[1]))
    (tmp_path / "c_instance.sieve").write_bytes(F.write_message(msgs[0]))
    (tmp_path / "notes.txt").write_bytes(b"not a sieve file")
    z = zkb()
    b = z.GpuBackend(-1)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_directory(tmp_path))
    assert b.stats()["callbacks"]["witness"] == 6 and b.pending_error() is None


def test_structural_errors_have_the_reference_text():
    h = fx.example_header()
    cases = [
        ([("Add", 2, 0, 1)], "No value given for wire_0"),
        ([("Constant", 0, b"\x01"), ("Constant", 0, b"\x02")], "Wire_0 already has a value in this scope."),
        ([("Instance", 0)], "Not enough instance to consume"),
        ([("Call", "nope", [], [])], "Unknown function"),
        ([("Constant", 0, b"\x01"), ("AnonCall", [ir.WireRange(3, 3)], [ir.Wire(0)], 0, 0, [("Copy", 0, 1)])],
         "In WireRange, last WireId (3) must be strictly greater than first WireId (3)."),
        ([("Free", 5, None)], "No value given for wire_5"),
    ]
    for gates, text in cases:
        rel = ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [], gates)
        assert ev.evaluate([rel]) == [text]                 # the oracle agrees with the reference text
        z, b, e = record(F.write_messages([rel]))
        assert b.pending_error() == text
    fn = ir.Function("f", 1, 2, 0, 0, [("Mul", 0, 1, 2)])
    rel = ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [fn],
                      [("Constant", 0, b"\x01"), ("Call", "f", [ir.Wire(1)], [ir.Wire(0)])])
    text = "Wrong number of input variables in call to function f (Expected 2 / Got 1)."
    assert ev.evaluate([rel]) == [text]
    z, b, e = record(F.write_messages([rel]))
    assert b.pending_error() == text


def test_set_field_errors():
    z = zkb()
    for mod, degree, text in [(b"\x00", 1, "Modulus cannot be zero."), (b"\x65", 2, "Field should be of degree 1")]:
        b = z.GpuBackend(-1)
        with pytest.raises(z.ZkbError) as e:
            b.set_field(mod, degree)
        assert str(e.value) == text and e.value.code == z.ZKB_E_SEMANTIC
    b = z.GpuBackend(-1)
    with pytest.raises(z.ZkbError) as e:
        b.set_field(100)
    assert e.value.code == z.ZKB_E_UNSUPPORTED
    b.set_field(101)
    assert b.minus_one() == bytes([100]) and b.one() == b"\x01" and b.zero() == b"\x00"


def test_panics_become_fatal_errors():
    z = zkb()
    h = fx.example_header()
    rel = ir.Relation(h, ir.ARITH, ir.SIMPLE, [], [("Witness", 0)])
    b = z.GpuBackend(-1)
    e = z.Evaluator(b)
    with pytest.raises(z.ZkbError) as err:
        e.ingest_source(z.Source.from_buffers([F.write_messages([rel])]))
    assert err.value.code == z.ZKB_E_FATAL and str(err.value) == "Missing witness value for PlaintextBackend"
    loop = ("For", "i", 0, 1, [ir.WireRange(0, 1)], ("IterExprAnonCall", [("Single", ("Name", "j"))], [], 0, 0,
                                                     [("Constant", 0, b"\x01")]))
    rel = ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [], [loop])
    e = z.Evaluator(z.GpuBackend(-1))
    with pytest.raises(z.ZkbError) as err:
        e.ingest_source(z.Source.from_buffers([F.write_messages([rel])]))
    assert err.value.code == z.ZKB_E_FATAL and str(err.value) == "Unknown iterator name j"


def test_malformed_messages_are_format_errors():
    z = zkb()
    good = F.write_message(fx.example_relation())
    for bad in (good[:40], good[:4] + b"\xff" * 60, b"\x10\x00\x00\x00" + b"\x00" * 16):
        e = z.Evaluator(z.GpuBackend(-1))
        with pytest.raises(z.ZkbError) as err:
            e.ingest_message(bad)
        assert err.value.code == z.ZKB_E_FORMAT


def test_library_exports_every_declared_symbol():
    import ctypes
    import re
    z = zkb()
    hdr = open(os.path.join(os.path.dirname(GOLDEN), "..", "include", "zkb.h")).read()
    declared = set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", hdr))
    lib = ctypes.CDLL(z.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(z.EXPORTED), declared ^ set(z.EXPORTED)


def test_levelizer_invariants():
    c = __import__("importlib").import_module("zkir_b200.circuits")
    z = zkb()
    circ = c.random_circuit(20000, 100, c.GOLDILOCKS, seed=3, n_tracked=10)
    b = z.GpuBackend(-1)
    b.set_field(c.GOLDILOCKS)
    b.push_gates(circ.gates, circ.const_pool)
    b.finalize(False)
    st = b.stats()
    assert st["ir_gates"] == 20000
    assert st["n_asserts"] == circ.n_asserts
    # every assertion here tests the result of an Add: all are fused into their producer
    assert st["n_device_ops"] == circ.hist["add"] + circ.hist["mul"]
    assert st["algo_bytes_per_witness"] == c.algorithmic_bytes_per_witness(circ)
    assert 1 <= st["n_levels"] < 200


def test_relation_split_over_several_messages():
    # builder.rs:72,90-95: a producer flushes a Relation message every 100 000 gates; functions defined in an
    # earlier message stay known, the wire namespace and the value queues continue (evaluator.rs:158-170, 273-300)
    msgs = STATEMENTS["example"]()
    inst, wit, rel = msgs
    parts = [ir.Relation(rel.header, rel.gate_mask, rel.feat_mask, rel.functions, rel.gates[:3]),
             ir.Relation(rel.header, rel.gate_mask, rel.feat_mask, [], rel.gates[3:8]),
             ir.Relation(rel.header, rel.gate_mask, rel.feat_mask, [], rel.gates[8:])]
    split = [inst, wit] + parts
    assert ev.evaluate(split) == []
    z, b, e = record(F.write_messages(split))
    compare_with_oracle_trace(b, split)
    z2, b2, e2 = record(F.write_messages(msgs))
    k1, k2 = b.program(), b2.program()
    assert all((x == y).all() for x, y in zip(k1, k2))
    # values arriving AFTER the first relation message are still consumed by later ones
    late = [inst, parts[0]]
    tb = ev.TracingBackend()
    with pytest.raises(ir.OraclePanic):          # ... but a witness that is not there yet is a panic in the reference
        ev.Evaluator.from_messages(late, tb)


def test_reader_survives_corrupted_buffers():
    """bounds-checked FlatBuffers walker: random corruption must end in ZKB_E_FORMAT / a latched error / success,
    never in a crash"""
    z = zkb()
    rng = np.random.default_rng(11)
    base = bytearray(F.write_message(fx.example_relation()))
    outcomes = {"ok": 0, "format": 0, "fatal": 0}
    for trial in range(400):
        buf = bytearray(base)
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(4, len(buf)))
            buf[pos] = int(rng.integers(0, 256))
        b = z.GpuBackend(-1)
        b.set_limits(max_values=1 << 20, max_steps=1 << 22)  # a corrupted For bound must not unroll for ever
        e = z.Evaluator(b)
        # the witness the (possibly still valid) relation needs
        e.ingest_message(F.write_message(fx.example_instance()))
        e.ingest_message(F.write_message(fx.example_witness()))
        try:
            e.ingest_message(bytes(buf))
            outcomes["ok"] += 1
        except z.ZkbError as err:
            assert err.code in (z.ZKB_E_FORMAT, z.ZKB_E_FATAL, z.ZKB_E_UNSUPPORTED, z.ZKB_E_SEMANTIC), err.code
            outcomes["format" if err.code == z.ZKB_E_FORMAT else "fatal"] += 1
    assert outcomes["format"] > 0 and outcomes["ok"] > 0, outcomes
    # truncations
    for cut in range(8, len(base), 97):
        e = z.Evaluator(z.GpuBackend(-1))
        with pytest.raises(z.ZkbError):
            e.ingest_message(bytes(base[:cut]))


def test_resource_limits_stop_hostile_loops():
    """zkb_set_limits: a For over 2^60 iterations / a 2^31-wire range ends in a latched error, not in an exhausted host"""
    z = zkb()
    h = fx.example_header()
    loop = ("For", "i", 0, (1 << 60), [], ("IterExprAnonCall", [], [], 0, 0, []))
    wide = ("AnonCall", [ir.WireRange(0, 1 << 31)], [], 0, 0, [])
    for gate, limits in [(loop, dict(max_steps=1 << 16)), (wide, dict(max_steps=1 << 16))]:
        rel = ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [], [gate])
        b = z.GpuBackend(-1)
        b.set_limits(**limits)
        e = z.Evaluator(b)
        e.ingest_source(z.Source.from_buffers([F.write_messages([rel])]))
        assert b.pending_error() == "zkb: resource limit exceeded (max_steps)"
    body = [("Constant", 0, b"\x01")]
    loop = ("For", "i", 0, 999, [ir.WireRange(0, 999)], ("IterExprAnonCall", [("Single", ("Name", "i"))], [], 0, 0, body))
    rel = ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [], [loop])
    b = z.GpuBackend(-1)
    b.set_limits(max_values=100)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_buffers([F.write_messages([rel])]))
    assert b.pending_error() == "zkb: resource limit exceeded (max_values)"
    with pytest.raises(z.ZkbError):
        b.set_limits(max_values=1 << 33)
