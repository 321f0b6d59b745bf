"""Flat relation messages recorded in bulk from the FlatBuffers tables (zkb_evaluator_ingest_buffer, SURVEY.md section 8f
row 1: no owned Gate structs) must leave exactly the state the serial gate loop leaves (evaluator.rs:288-301): the same
SSA program, assertions, callback counts, live wires and plan — and anything irregular (Copy, Free and re-use, a wire
bound twice, an operand bound later or never, too few witness values) must surface the reference's error text, i.e.
fall back to the serial path.  Host-only: no device needed."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import evaluator as ev, ir, sieve_fbs as F
from tests.util import FIELDS, ROOT, circuits, zkb

_CHILD = r"""
import os, sys, json, numpy as np
sys.path.insert(0, {root!r})
import zkb_loader
z = zkb_loader.load()
buf = open(sys.argv[1], "rb").read()
be = z.GpuBackend(-1)
e = z.Evaluator(be)
out = {{}}
try:
    e.ingest_source(z.Source.from_buffers([buf]))
    out["pending"] = be.pending_error()
    st = be.stats()
    k, a, b = be.program()
    out["stats"] = {{kk: st[kk] for kk in ("n_values", "n_asserts", "n_instance", "n_witness", "n_consts", "ir_gates", "callbacks")}}
    import hashlib
    out["program"] = hashlib.sha256(k.tobytes() + a.tobytes() + b.tobytes()).hexdigest()
    out["asserts"] = [(be.assert_wire(i), be.assert_value(i)) for i in range(min(st["n_asserts"], 2000))]
    if not out["pending"]:
        wires = [int(w) for w in sys.argv[2].split(",") if w]
        out["live"] = [e.value_handle(w) for w in wires]
        be.finalize()
        out["plan"] = be.plan_hash()
except z.ZkbError as err:
    out["error"] = [err.code, str(err)]
print(json.dumps(out))
"""


def _run(tmp_path, buf, wires, fast):
    script = tmp_path / "child.py"
    script.write_text(_CHILD.format(root=ROOT))
    f = tmp_path / "stmt.sieve"
    f.write_bytes(buf)
    env = dict(os.environ, ZKB_PARALLEL_INGEST_MIN_BYTES="0", ZKB_PARSE_THREADS="4", ZKB_TIMING="1")
    if not fast:
        env["ZKB_NO_FLAT_INGEST"] = "1"
    r = subprocess.run([sys.executable, str(script), str(f), ",".join(str(w) for w in wires)], capture_output=True, text=True, env=env,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    return json.loads(r.stdout.strip().splitlines()[-1]), r.stderr


def _messages(p, gates, chunk, n_wit_values, n_inst_values=0):
    h = ir.Header(ir.le_bytes(p))
    msgs = []
    if n_inst_values:
        msgs.append(ir.Instance(h, [ir.le_bytes(3 + i) for i in range(n_inst_values)]))
    msgs.append(ir.Witness(h, [ir.le_bytes(5 + i) for i in range(n_wit_values)]))
    for i in range(0, len(gates), chunk):
        msgs.append(ir.Relation(h, ir.ARITH | ir.BOOL, ir.SIMPLE, [], gates[i:i + chunk]))
    return msgs


def _regular_gates(rng, n, n_wit, n_inst):
    """every value-defining simple gate kind, operands anywhere earlier (also in earlier messages)"""
    gates = [("Witness", i) for i in range(n_wit)] + [("Instance", n_wit + i) for i in range(n_inst)]
    nxt = n_wit + n_inst
    gates.append(("Constant", nxt, ir.le_bytes(7)))
    nxt += 1
    kinds = ["Add", "Mul", "And", "Xor", "Not", "AddConstant", "MulConstant", "Constant", "AssertZero", "Witness"]
    wit_extra = 0
    for _ in range(n):
        k = kinds[int(rng.integers(0, len(kinds)))]
        a, b = int(rng.integers(0, nxt)), int(rng.integers(0, nxt))
        if k in ("Add", "Mul", "And", "Xor"):
            gates.append((k, nxt, a, b))
        elif k == "Not":
            gates.append((k, nxt, a))
        elif k in ("AddConstant", "MulConstant"):
            gates.append((k, nxt, a, ir.le_bytes(int(rng.integers(0, 90)))))
        elif k == "Constant":
            gates.append((k, nxt, ir.le_bytes(int(rng.integers(0, 90)))))
        elif k == "Witness":
            gates.append((k, nxt))
            wit_extra += 1
        else:
            gates.append((k, a))
            continue
        nxt += 1
    return gates, nxt, wit_extra


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_bulk_recording_equals_the_serial_gate_loop(seed, tmp_path):
    rng = np.random.default_rng(seed)
    p = FIELDS["p101"] if seed != 3 else FIELDS["bls381"]
    gates, n_wires, wit_extra = _regular_gates(rng, 3000, 6, 3 if seed == 2 else 0)
    msgs = _messages(p, gates, 137, 6 + wit_extra, 3 if seed == 2 else 0)
    buf = F.write_messages(msgs)
    wires = list(range(0, n_wires, 11))
    fast, err_fast = _run(tmp_path, buf, wires, True)
    slow, _ = _run(tmp_path, buf, wires, False)
    n_rel = sum(isinstance(m, ir.Relation) for m in msgs)
    assert f"({n_rel} of {len(msgs)} messages)" in err_fast      # the bulk path took every relation message
    assert fast == slow
    assert fast["stats"]["n_values"] == n_wires and "error" not in fast
    # and the oracle's Evaluator drives its backend through the same callbacks, the same number of times
    tb = ev.TracingBackend()
    ev.Evaluator.from_messages(msgs, tb)
    counts = {}
    for k, _, _ in tb.trace:
        counts[k] = counts.get(k, 0) + 1
    if not any(v for v in ev.Evaluator.from_messages(msgs, ev.PlaintextBackend()).get_violations()):
        names = {"constant": "constant", "instance": "instance", "witness": "witness", "add": "add", "mul": "mul", "addc": "addc",
                 "mulc": "mulc", "and": "and", "xor": "xor", "not": "not", "copy": "copy", "assert_zero": "assert_zero"}
        for k, v in fast["stats"]["callbacks"].items():
            assert counts.get(names[k], 0) == v, k


def _irregular_cases():
    W = [("Witness", 0), ("Witness", 1)]
    return {
        "copy": W + [("Add", 2, 0, 1), ("Copy", 3, 2), ("Mul", 4, 3, 0), ("AssertZero", 4)],
        "free_and_reuse": W + [("Add", 2, 0, 1), ("Free", 2, 2), ("Mul", 2, 0, 1), ("AssertZero", 2)],
        "bound_twice_same_message": W + [("Add", 2, 0, 1), ("Mul", 2, 0, 1)],
        "bound_twice_across_messages": W + [("Add", 2, 0, 1), ("Add", 3, 2, 1), ("Add", 4, 3, 1), ("Mul", 2, 0, 1)],
        "operand_bound_later": W + [("Add", 2, 0, 3), ("Mul", 3, 0, 1)],
        "operand_never_bound": W + [("Add", 2, 0, 1), ("Mul", 3, 2, 77)],
        "assert_on_unbound_wire": W + [("Add", 2, 0, 1), ("AssertZero", 9)],
        "too_few_witness_values": W + [("Add", 2, 0, 1), ("Witness", 3), ("Witness", 4), ("Add", 5, 3, 4)],
        "too_few_instance_values": W + [("Add", 2, 0, 1), ("Instance", 3)],
        "huge_wire_id": W + [("Add", (1 << 40) + 5, 0, 1), ("Mul", 3, (1 << 40) + 5, 1)],
    }


@pytest.mark.parametrize("name", list(_irregular_cases()))
def test_irregular_windows_take_the_serial_path_and_report_like_the_reference(name, tmp_path):
    gates = _irregular_cases()[name]
    p = 101
    msgs = _messages(p, gates, 3, 2)
    buf = F.write_messages(msgs)
    fast, _ = _run(tmp_path, buf, [0, 1], True)
    slow, _ = _run(tmp_path, buf, [0, 1], False)
    assert fast == slow, name
    want = None
    try:
        viol = ev.evaluate(msgs)
        want = viol[-1] if viol else None
    except ir.OraclePanic as e:        # the reference aborts (missing witness value)
        assert "error" in fast and fast["error"][0] == zkb().ZKB_E_FATAL and str(e) in fast["error"][1]
        return
    if name in ("copy", "free_and_reuse"):
        assert fast.get("pending") is None and "error" not in fast
    elif name == "huge_wire_id":
        assert fast.get("pending") is None
    else:
        assert fast.get("pending") == want, (fast, want)
