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
