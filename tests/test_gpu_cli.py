"""The `evaluate` verb (zkinterface-ir_b200/zkb) on committed `.sieve` workspaces: same lines on stderr and the
same exit behaviour as `zki_sieve evaluate` (rust/src/cli.rs:315-320, 557-571)."""
import os
import subprocess

import pytest

from oracle import evaluator as ev
from oracle import sieve_fbs as F
from tests.util import ROOT, zkb

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "zkinterface-ir_b200", "zkb")

EXPECT = {
    "": [],                                                              # the reference's own binary fixtures
    "example": [],
    "example_incorrect": ["Wire_9 (may be weighted) should be 0, while it is not"],
    "boolean_example": [],
    "boolean_example_incorrect": ["Wire_22 (may be weighted) should be 0, while it is not"],
    "builder_switch": [],
    "builder_switch_nested": [],
    "r1cs_example": [],
}


@pytest.mark.parametrize("ws", list(EXPECT))
def test_cli_evaluate(ws):
    path = os.path.join(GOLDEN, ws) if ws else GOLDEN
    # the committed bytes mean what the oracle says they mean
    msgs = []
    for fn in ("000_instance.sieve", "001_witness.sieve", "002_relation.sieve"):
        p = os.path.join(path, fn)
        if os.path.exists(p):
            msgs += F.read_messages(open(p, "rb").read())
    assert ev.evaluate(msgs) == EXPECT[ws]
    r = subprocess.run([CLI, "evaluate", path], capture_output=True, text=True, timeout=120)
    if EXPECT[ws]:
        assert r.returncode != 0
        assert "The statement is NOT TRUE!" in r.stderr
        assert f"- {EXPECT[ws][0]}" in r.stderr
        assert "Found 1 violations." in r.stderr
    else:
        assert r.returncode == 0, r.stderr
        assert "The statement is TRUE!" in r.stderr


def test_python_source_on_the_same_workspaces():
    z = zkb()
    for ws, want in EXPECT.items():
        path = os.path.join(GOLDEN, ws) if ws else GOLDEN
        files = [os.path.join(path, f) for f in sorted(os.listdir(path)) if f.endswith(".sieve")]
        e = z.Evaluator.from_messages(z.Source.from_dirs_and_files(files), device=0)
        assert e.get_violations() == want


def test_cli_valid_eval_metrics_and_stdin():
    """cli.rs:333-363: validator + evaluator + stats over the same messages; `-` reads the statement from stdin"""
    import json
    from oracle import stats as os_
    ws = os.path.join(GOLDEN, "example_incorrect")
    r = subprocess.run([CLI, "valid-eval-metrics", ws], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "The statement is COMPLIANT with the specification!" in r.stderr
    assert "The statement is NOT TRUE!" in r.stderr and "- Wire_9 (may be weighted) should be 0, while it is not" in r.stderr
    msgs = [m for n in sorted(os.listdir(ws)) for m in F.read_messages(open(os.path.join(ws, n), "rb").read())]
    assert json.loads(r.stdout) == os_.stats(msgs).as_dict()
    stream = b"".join(open(os.path.join(GOLDEN, "example", n), "rb").read() for n in sorted(os.listdir(os.path.join(GOLDEN, "example"))))
    r = subprocess.run([CLI, "valid-eval-metrics", "-"], input=stream, capture_output=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert b"The statement is TRUE!" in r.stderr and b"COMPLIANT" in r.stderr
    r = subprocess.run([CLI, "evaluate", "-"], input=stream, capture_output=True, timeout=120)
    assert r.returncode == 0 and b"The statement is TRUE!" in r.stderr
