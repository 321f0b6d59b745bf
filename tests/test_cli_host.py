"""Host-only verbs of the `zkb` CLI: `validate` (cli.rs:302-313) and `flatten` (cli.rs:442-472), output strings and
exit codes as the reference's print_violations (cli.rs:557-571)."""
import os
import subprocess

from oracle import evaluator as ev
from oracle import fixtures as fx
from oracle import sieve_fbs as F
from oracle import validator as ov

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "zkinterface-ir_b200", "zkb")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def run(*args, stdin=None):
    return subprocess.run([CLI, *args], input=stdin, capture_output=True, timeout=60)


def test_validate_verb(tmp_path):
    r = run("validate", os.path.join(GOLDEN, "example"))
    assert r.returncode == 0 and r.stderr.decode() == "\nThe statement is COMPLIANT with the specification!\n"
    bad = fx.example_relation()
    bad.gates = list(bad.gates) + [("Free", 1, 2), ("Free", 4, None)]
    (tmp_path / "000_instance.sieve").write_bytes(F.write_message(fx.example_instance()))
    (tmp_path / "001_witness.sieve").write_bytes(F.write_message(fx.example_witness()))
    (tmp_path / "002_relation.sieve").write_bytes(F.write_message(bad))
    r = run("validate", str(tmp_path))
    want = ov.validate([fx.example_instance(), fx.example_witness(), bad])
    assert r.returncode == 1
    assert r.stderr.decode() == ("\nThe statement is NOT COMPLIANT with the specification!\nViolations:\n- " + "\n- ".join(want) +
                                 f"\n\nError: Found {len(want)} violations.\n")


def test_flatten_verb_to_dir_and_stdout(tmp_path):
    src = os.path.join(GOLDEN, "builder_switch")
    out = tmp_path / "flat"
    r = run("flatten", "--out", str(out), src)
    assert r.returncode == 0, r.stderr
    assert sorted(os.listdir(out)) == ["000_instance.sieve", "001_witness.sieve", "002_relation.sieve"]
    msgs = [m for name in sorted(os.listdir(out)) for m in F.read_messages((out / name).read_bytes())]
    assert ev.evaluate(msgs) == [] and ov.validate(msgs) == []
    # --out - : the three buffers on stdout; "-" as input: the statement from stdin
    stream = b"".join(open(os.path.join(src, n), "rb").read() for n in sorted(os.listdir(src)))
    r2 = run("flatten", "--out", "-", "-", stdin=stream)
    assert r2.returncode == 0
    assert r2.stdout == b"".join((out / n).read_bytes() for n in sorted(os.listdir(out)))
    # a .sieve path is not a directory (cli.rs:459-460); files and stdin cannot be combined (source.rs:170-173)
    r3 = run("flatten", "--out", str(tmp_path / "x.sieve"), src)
    assert r3.returncode == 1 and b"IR flattening requires a directory as output value" in r3.stderr
    r4 = run("validate", src, "-", stdin=b"")
    assert r4.returncode == 1 and b"Cannot combine files and stdin" in r4.stderr
