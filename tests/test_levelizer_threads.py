"""The levelizer's threaded passes (ZKB_PLAN_THREADS) must build exactly the plan the sequential form builds: flat random
circuits with slot re-use (verdicts only), live wires, keep-all; a wide shallow Boolean loop nest; a deep narrow chain."""
import os
import subprocess
import sys

import pytest

from tests.util import ROOT

SNIPPET = r'''
import sys, importlib
sys.path.insert(0, %r)
import numpy as np
import zkb_loader
z = zkb_loader.load()
c = importlib.import_module("zkir_b200.circuits")
from oracle import ir, sieve_fbs as F, workloads as wl
out = []
for p, lg, window in ((c.BLS12_381_FR, 19, 0), (c.GOLDILOCKS, 19, 4096)):
    circ = c.random_circuit(1 << lg, 256, p, 31 + lg, window=window)
    for mode in ("live", "all", "verdicts"):
        b = z.GpuBackend(-1)
        b.set_field(p)
        b.push_gates(circ.gates, circ.const_pool)
        b.finalize(keep_all_values=(mode == "all"), verdicts_only=(mode == "verdicts"))
        st = b.stats()
        out.append((mode, st["n_slots"], st["n_levels"], st["n_device_ops"], b.plan_hash()))
rel, n_leaf = wl.boolean_for_relation(9, 8, 2048)
buf = F.write_messages([ir.Witness(rel.header, [b"\0"] * 2048), rel])
for keep in (0, 2):
    b = z.GpuBackend(-1)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_buffers([buf]))
    b.finalize(verdicts_only=(keep == 2))
    out.append(("c5", keep, b.stats()["n_slots"], b.plan_hash()))
print(out)
'''


def run(threads):
    env = dict(os.environ, ZKB_PLAN_THREADS=str(threads))
    r = subprocess.run([sys.executable, "-c", SNIPPET % ROOT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    return r.stdout.strip().splitlines()[-1]


def test_threaded_levelizer_builds_the_sequential_plan():
    seq = run(1)
    assert seq == run(4)
    assert seq == run(7)


SNIPPET_CALLS = r'''
import sys, hashlib
sys.path.insert(0, %r)
import zkb_loader
z = zkb_loader.load()
from oracle import ir, sieve_fbs as F, fixtures as fx
from tests.gen_programs import Gen
out = []
for seed in range(24):
    boolean = seed %% 3 == 2
    msgs = Gen(seed, 2 if boolean else 101, boolean=boolean).statement()
    b = z.GpuBackend(-1)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_buffers([F.write_messages(msgs)]))
    k, a, bb = b.program()
    st = b.stats()
    out.append((hashlib.sha1(k.tobytes() + a.tobytes() + bb.tobytes()).hexdigest(), st["n_asserts"], st["ir_gates"],
                tuple(sorted(st["callbacks"].items())), b.pending_error()))
# error cases must surface identically: undefined wire inside a body, output written twice, output never written
h = fx.example_header()
bad_bodies = [[("Add", 0, 1, 5)], [("Add", 0, 1, 2), ("Mul", 0, 1, 2)], [("Add", 3, 1, 2)]]
for body in bad_bodies:
    fn = ir.Function("f", 1, 2, 0, 0, body)
    rel = ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [fn], [("Constant", 7, b"\x02"), ("Constant", 8, b"\x03"),
                      ("Call", "f", [ir.Wire(9)], [ir.Wire(7), ir.Wire(8)]), ("Call", "f", [ir.Wire(9)], [ir.Wire(7), ir.Wire(8)])])
    b = z.GpuBackend(-1)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_buffers([F.write_messages([rel])]))
    out.append((b.pending_error(), b.stats()["n_values"]))
good = ir.Function("g", 1, 2, 0, 0, [("Mul", 0, 1, 2)])
rel = ir.Relation(h, ir.ARITH, ir.FOR_FUNCTION_SWITCH, [good], [("Constant", 7, b"\x02"), ("Call", "g", [ir.Wire(9)], [ir.Wire(7), ir.Wire(7)]),
                  ("Call", "g", [ir.Wire(9)], [ir.Wire(7), ir.Wire(9)]), ("Call", "g", [ir.Wire(10)], [ir.Wire(7), ir.Wire(11)])])
b = z.GpuBackend(-1)
e = z.Evaluator(b)
e.ingest_source(z.Source.from_buffers([F.write_messages([rel])]))
out.append((b.pending_error(), b.stats()["n_values"], tuple(sorted(b.stats()["callbacks"].items()))))
print(out)
'''


def run_calls(no_simple):
    env = dict(os.environ)
    if no_simple:
        env["ZKB_NO_SIMPLE_CALLS"] = "1"
    r = subprocess.run([sys.executable, "-c", SNIPPET_CALLS % ROOT], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    return r.stdout.strip().splitlines()[-1]


def test_simple_function_fast_path_records_what_the_generic_path_records():
    fast = run_calls(False)
    assert fast == run_calls(True)
    assert "already has a value" in fast and "No value given for wire_" in fast
