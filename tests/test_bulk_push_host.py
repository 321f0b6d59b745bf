"""zkb_push_gates on long regular runs takes a threaded bulk path (csrc/backend.cu: push_gates_bulk — count / define / resolve
over chunks of the gate array).  It must leave exactly the state the gate loop leaves (ZKB_NO_BULK_PUSH=1), and anything
irregular must fall back to the gate loop and report what the reference reports (evaluator.rs:775-797)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.util import FIELDS, ROOT, circuits, zkb

SNIPPET = r'''
import sys, importlib, hashlib
sys.path.insert(0, %r)
import numpy as np
import zkb_loader
z = zkb_loader.load()
c = importlib.import_module("zkir_b200.circuits")
out = []
for p, n, window in ((c.BLS12_381_FR, 1 << 17, 0), (c.GOLDILOCKS, 1 << 17, 512), ((1 << 31) - 1, 1 << 18, 64)):
    circ = c.random_circuit(n, 256, p, 77, window=window)
    gates, pool = circ.gates, circ.const_pool
    b = z.GpuBackend(-1)
    b.set_field(p)
    half = len(gates) // 2
    b.push_gates(gates[:half], pool)          # two calls: the second one resolves wires the first one bound
    b.push_gates(gates[half:], pool)
    k, a, bb = b.program()
    st = b.stats()
    h = hashlib.sha256(k.tobytes() + a.tobytes() + bb.tobytes()).hexdigest()
    asserts = [(b.assert_value(s), b.assert_wire(s)) for s in range(0, st["n_asserts"], 97)]
    b.finalize(False)
    out.append((h, st["n_values"], st["n_asserts"], st["ir_gates"], sorted(st["callbacks"].items()), st["n_instance"], st["n_witness"],
                asserts, b.plan_hash(), [b.scope_lookup(int(w)) for w in gates["out"][gates["op"] != c.G_ASSERT_ZERO][::1013]]))
print(out)
'''


def run(env_extra):
    env = dict(os.environ, ZKB_PLAN_THREADS="6", **env_extra)
    r = subprocess.run([sys.executable, "-c", SNIPPET % ROOT], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()[-1]


def test_bulk_push_leaves_the_state_of_the_gate_loop():
    assert run({}) == run({"ZKB_NO_BULK_PUSH": "1"})


def _long_circuit():
    c = circuits()
    p = FIELDS["goldilocks"]
    circ = c.random_circuit(1 << 17, 64, p, 9)
    return c, p, circ


@pytest.mark.parametrize("damage", ["redefine", "undefined", "later", "copy", "free", "const_index"])
def test_irregular_runs_fall_back_and_report_like_the_reference(damage):
    z = zkb()
    c, p, circ = _long_circuit()
    g = circ.gates.copy()
    value_gates = np.flatnonzero((g["op"] == c.G_ADD) | (g["op"] == c.G_MUL))
    i = int(value_gates[len(value_gates) // 2])
    expect_err = None
    if damage == "redefine":
        g["out"][i] = g["out"][i - 1] if g["op"][i - 1] != c.G_ASSERT_ZERO else g["out"][int(value_gates[0])]
        expect_err = f"Wire_{int(g['out'][i])} already has a value in this scope."
    elif damage == "undefined":
        g["a"][i] = int(g["out"].max()) + 1000
        expect_err = f"No value given for wire_{int(g['a'][i])}"
    elif damage == "later":
        j = int(value_gates[-1])
        g["b"][i] = g["out"][j]             # bound later in program order
        expect_err = f"No value given for wire_{int(g['b'][i])}"
    elif damage == "copy":
        extra = np.zeros(1, dtype=c.GATE_DTYPE)
        extra["op"], extra["out"], extra["a"] = c.G_COPY, int(g["out"].max()) + 1, g["out"][i]
        g = np.concatenate([g, extra])
    elif damage == "free":
        extra = np.zeros(1, dtype=c.GATE_DTYPE)
        extra["op"], extra["a"], extra["b"] = c.G_FREE, g["out"][i], g["out"][i]
        g = np.concatenate([g, extra])
    elif damage == "const_index":
        extra = np.zeros(1, dtype=c.GATE_DTYPE)
        extra["op"], extra["out"], extra["b"] = c.G_CONSTANT, int(g["out"].max()) + 1, 10 ** 6
        g = np.concatenate([g, extra])
        expect_err = "constant index out of range"
    results = []
    for env in ({}, {"ZKB_NO_BULK_PUSH": "1"}):
        for k in ("ZKB_NO_BULK_PUSH",):
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            b = z.GpuBackend(-1)
            b.set_field(p)
            try:
                b.push_gates(g, circ.const_pool)
                err = None
            except z.ZkbError as e:
                err = str(e)
            st = b.stats()
            results.append((err, st["n_values"], st["n_asserts"], sorted(st["callbacks"].items()), b.pending_error()))
        finally:
            os.environ.pop("ZKB_NO_BULK_PUSH", None)
    assert results[0] == results[1]
    if expect_err:
        assert results[0][0] == expect_err
    elif damage in ("copy", "free"):
        assert results[0][0] is None
