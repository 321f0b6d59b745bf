"""Every consumer of `.sieve` bytes (Evaluator in flatten mode, Validator, Stats, reader -> writer) must survive corrupted
and truncated messages: an error code or a result, never a crash or an exhausted host."""
import json
import subprocess
import sys
import os

import numpy as np
import pytest

from oracle import fixtures as fx
from oracle import sieve_fbs as F
from tests.util import ROOT, zkb

OK_CODES = None


def corrupted(base, rng):
    buf = bytearray(base)
    for _ in range(int(rng.integers(1, 6))):
        buf[int(rng.integers(4, len(buf)))] = int(rng.integers(0, 256))
    return bytes(buf)


@pytest.mark.parametrize("which", ["example", "boolean"])
def test_consumers_survive_corrupted_relations(which):
    z = zkb()
    rng = np.random.default_rng(23)
    if which == "example":
        pre, rel = [fx.example_instance(), fx.example_witness()], fx.example_relation()
    else:
        pre, rel = [fx.boolean_example_instance(), fx.boolean_example_witness()], fx.boolean_example_relation()
    pre_bytes = [F.write_message(m) for m in pre]
    base = F.write_message(rel)
    allowed = (z.ZKB_E_FORMAT, z.ZKB_E_FATAL, z.ZKB_E_UNSUPPORTED, z.ZKB_E_SEMANTIC, z.ZKB_E_ARG)
    outcomes = {"ok": 0, "err": 0}
    host = z.GpuBackend(-1)
    for trial in range(300):
        bad = corrupted(base, rng)
        # Validator
        v = z.Validator(True)
        v.set_limits(1 << 20)
        try:
            for m in pre_bytes:
                v.ingest_message(m)
            v.ingest_message(bad)
            assert isinstance(v.get_violations(), list)
            outcomes["ok"] += 1
        except z.ZkbError as e:
            assert e.code in allowed
            outcomes["err"] += 1
        # Stats
        st = z.Stats()
        try:
            st.ingest_message(bad)
            json.loads(st.to_json_pretty())
        except z.ZkbError as e:
            assert e.code in allowed
        # flatten
        b = z.GpuBackend(-1)
        b.set_limits(max_values=1 << 20, max_steps=1 << 22)
        ev = z.Evaluator(b, flatten=True)
        try:
            for m in pre_bytes:
                ev.ingest_message(m)
            ev.ingest_message(bad)
            bufs = ev.flatten()
            for part in bufs:
                for m in F.split_messages(part):
                    F.read_message(m)           # whatever was written is well-formed
        except z.ZkbError as e:
            assert e.code in allowed
        # reader -> writer
        try:
            again = host.rewrite_message(bad)
            assert host.rewrite_message(again) == again
        except z.ZkbError as e:
            assert e.code in allowed
    assert outcomes["ok"] > 0 and outcomes["err"] > 0, outcomes
    for cut in range(8, len(base), 131):
        for consume in (lambda b_: z.Validator(True).ingest_message(b_), lambda b_: z.Stats().ingest_message(b_),
                        lambda b_: host.rewrite_message(b_)):
            with pytest.raises(z.ZkbError):
                consume(base[:cut])


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--log2-gates", "16", "--cpu-sample-log2-gates", "14"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["config"]["workload"].startswith("C3")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_deep_nesting_is_a_format_error_not_a_stack_overflow():
    from oracle import ir
    z = zkb()
    h = fx.example_header()
    g = ("Constant", 0, b"\x01")
    for _ in range(600):
        g = ("AnonCall", [ir.Wire(0)], [], 0, 0, [g])
    sys.setrecursionlimit(20000)
    buf = F.write_message(ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [], [g]))
    for consume in (lambda: z.Validator(True).ingest_message(buf), lambda: z.Stats().ingest_message(buf),
                    lambda: z.Evaluator(z.GpuBackend(-1)).ingest_message(buf), lambda: z.GpuBackend(-1).rewrite_message(buf)):
        with pytest.raises(z.ZkbError) as e:
            consume()
        assert e.value.code == z.ZKB_E_FORMAT and "nested" in str(e.value)
    e = ("Const", 1)
    for _ in range(600):
        e = ("Add", e, ("Const", 1))
    loop = ("For", "i", 0, 0, [], ("IterExprAnonCall", [("Single", e)], [], 0, 0, []))
    buf = F.write_message(ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [], [loop]))
    with pytest.raises(z.ZkbError) as err:
        z.Validator(True).ingest_message(buf)
    assert err.value.code == z.ZKB_E_FORMAT


def test_recursive_functions_end_in_an_error_not_a_stack_overflow():
    """Functions are registered before the relation's gates run, so a body may call itself or a sibling; the reference
    recurses until its stack overflows, here the nesting bound of the reader (512) also bounds dynamic nesting."""
    from oracle import ir
    z = zkb()
    h = fx.example_header()
    inst, wit = F.write_message(fx.example_instance()), F.write_message(fx.example_witness())
    selfrec = ir.Function("f", 0, 0, 0, 0, [("Call", "f", [], [])])
    ping = ir.Function("ping", 0, 0, 0, 0, [("Call", "pong", [], [])])
    pong = ir.Function("pong", 0, 0, 0, 0, [("AnonCall", [], [], 0, 0, [("Call", "ping", [], [])])])
    for fns, entry in (([selfrec], "f"), ([ping, pong], "ping")):
        rel = F.write_message(ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, fns, [("Call", entry, [], [])]))
        ev = z.Evaluator(z.GpuBackend(-1), flatten=True)
        ev.ingest_message(inst)
        ev.ingest_message(wit)
        ev.ingest_message(rel)            # the error latches like any evaluation error (evaluator.rs:213-221)
        assert "nested too deep" in (ev.backend.pending_error() or "")


def test_huge_wire_ids_cost_memory_proportional_to_the_work_not_to_the_id():
    """a 1 KB relation naming wire 2^28-1 inside four nested AnonCalls took 4 GiB when scopes were dense tables"""
    import resource
    from oracle import ir
    z = zkb()
    h = fx.example_header()
    big = (1 << 28) - 1
    g = ("Constant", big, b"\x01")
    body = [g, ("Free", big, None)]
    for _ in range(4):
        body = [("Constant", big, b"\x02"), ("AnonCall", [], [], 0, 0, body), ("Free", big, None)]
    far = [("Constant", (1 << 40) + 5, b"\x03"), ("Copy", 1 << 33, (1 << 40) + 5), ("Free", 1 << 33, None)]
    rel = F.write_message(ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [], body + far))
    before = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    for mk in (lambda: z.Evaluator(z.GpuBackend(-1), flatten=True), lambda: z.Validator(True)):
        c = mk()
        c.ingest_message(rel)
        if isinstance(c, z.Validator):
            assert c.get_violations() == ["The variable 1099511627781 is still live"] or len(c.get_violations()) <= 2
    after = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    assert after - before < 200 * 1024, (before, after)      # KiB


def test_shared_subtables_cannot_expand_a_small_message_into_millions_of_gates():
    """FlatBuffers offsets may point every entry of a vector at one shared table: 8 entries per level, 9 levels deep is
    2^27 AnonCall gates out of ~2 KB; the reader charges every table view against the message length"""
    import struct
    from oracle import ir
    z = zkb()

    def shared_vec(k, child):
        def th(w):
            w.align(4)
            pos = len(w.b)
            w.b += struct.pack("<I", k) + bytes(4 * k)
            c = child(w)
            for i in range(k):
                struct.pack_into("<I", w.b, pos + 4 + 4 * i, c - (pos + 4 + 4 * i))
            return pos
        return th

    def anon(depth):
        body = shared_vec(8, anon(depth - 1)) if depth else F._w_gates([])
        sub = lambda w: w.table([(4, "ref", F._w_wirelist([])), (10, "ref", body)])
        return lambda w: w.table([(4, "u8", F.DS_ID["AnonCall"]),
                                  (6, "ref", lambda w2: w2.table([(4, "ref", F._w_wirelist([])), (6, "ref", sub)]))])

    h = fx.example_header()
    w = F._W()
    w.b += bytes(12)
    w.b[8:12] = b"siev"
    body = lambda w2: w2.table([(4, "ref", F._w_header(h)), (6, "ref", F._w_str("arithmetic")),
                                (8, "ref", F._w_str("@for,@function,@switch")),
                                (10, "ref", lambda w3: w3.table_vec([])), (12, "ref", shared_vec(8, anon(8)))])
    root = w.table([(4, "u8", F.MSG_RELATION), (6, "ref", body)])
    w.align(4)
    struct.pack_into("<I", w.b, 4, root - 4)
    struct.pack_into("<I", w.b, 0, len(w.b) - 4)
    buf = bytes(w.b)
    assert len(buf) < 4096
    for consume in (lambda: z.Validator(True).ingest_message(buf), lambda: z.Stats().ingest_message(buf),
                    lambda: z.Evaluator(z.GpuBackend(-1)).ingest_message(buf), lambda: z.GpuBackend(-1).rewrite_message(buf)):
        with pytest.raises(z.ZkbError) as e:
            consume()
        assert e.value.code == z.ZKB_E_FORMAT and "more tables" in str(e.value)


def test_input_arrays_smaller_than_the_batch_are_refused_before_the_c_side_reads_them():
    z = zkb()
    b = z.GpuBackend(-1)
    b.set_field(101)
    g = np.zeros(3, dtype=z.GATE_DTYPE)
    g["op"] = [z.G_INSTANCE, z.G_WITNESS, z.G_WITNESS]
    g["out"] = [0, 1, 2]
    b.push_gates(g)
    b.finalize()
    one = np.zeros((1, 4), np.uint8)
    two = np.zeros((2, 4), np.uint8)
    with pytest.raises(z.ZkbError) as e:
        b.evaluate(one, one, 1)                    # one witness value, the program consumes two: the reference panics
    assert e.value.code == z.ZKB_E_FATAL and "Missing witness value" in str(e.value)
    with pytest.raises(z.ZkbError) as e:
        b.evaluate(np.zeros((0, 4), np.uint8), two, 1)
    assert e.value.code == z.ZKB_E_SEMANTIC and str(e.value) == "Not enough instance to consume"
    with pytest.raises(z.ZkbError) as e:
        b.evaluate(one, np.zeros((3, 2, 4), np.uint8), 5)   # three value sets for a batch of five
    assert e.value.code == z.ZKB_E_ARG
    with pytest.raises(z.ZkbError) as e:
        b.evaluate(one, two, 1)                    # well-formed: only the missing device is left to complain about
    assert e.value.code == z.ZKB_E_CUDA
