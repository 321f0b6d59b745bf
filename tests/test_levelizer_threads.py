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
