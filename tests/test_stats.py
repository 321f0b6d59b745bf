"""Stats (rust/src/consumers/stats.rs): oracle pinned on the reference's `test_stats` expected counts, the C++ mirror
behind the C ABI must print the same JSON object for every statement."""
import json

import pytest

from oracle import fixtures as fx
from oracle import sieve_fbs as F
from oracle import stats as os_
from tests.gen_programs import Gen
from tests.test_host_evaluator import STATEMENTS
from tests.util import zkb

EXPECTED_EXAMPLE = {            # stats.rs:302-341
    "instance_variables": 3, "witness_variables": 4, "constants_gates": 1, "assert_zero_gates": 6, "copy_gates": 0, "add_gates": 25,
    "mul_gates": 21, "add_constant_gates": 0, "mul_constant_gates": 1, "and_gates": 0, "xor_gates": 0, "not_gates": 0,
    "variables_freed": 51, "functions_defined": 1, "functions_called": 20, "switches": 1, "branches": 2, "for_loops": 2,
    "instance_messages": 1, "witness_messages": 1, "relation_messages": 1,
}


def test_oracle_stats_pinned_on_test_stats():
    s = os_.stats([fx.example_instance(), fx.example_witness(), fx.example_relation()])
    assert s.gate_stats == EXPECTED_EXAMPLE
    assert s.field_characteristic == bytes([101, 0, 0, 0]) and s.field_degree == 1
    mul = os_.new_gate_stats()
    mul["mul_gates"] = 1
    assert s.functions == {"com.example::mul": (mul, 0, 0)}


def ours(msgs):
    z = zkb()
    st = z.Stats()
    st.ingest_source(z.Source.from_buffers([F.write_messages(msgs)]))
    return st.to_json_pretty()


@pytest.mark.parametrize("name", list(STATEMENTS))
def test_stats_match_oracle(name):
    msgs = STATEMENTS[name]()
    want = os_.stats(msgs)
    got = ours(msgs)
    assert json.loads(got) == want.as_dict()
    # same pretty layout as serde_json (two-space indent, declaration order); functions sorted by name here
    d = want.as_dict()
    d["functions"] = dict(sorted(d["functions"].items()))
    assert got == json.dumps(d, indent=2)
    assert list(json.loads(got)["gate_stats"]) == os_.GATE_FIELDS


@pytest.mark.parametrize("seed", range(12))
def test_stats_match_oracle_on_random_structured_programs(seed):
    boolean = seed % 3 == 2
    msgs = Gen(seed, 2 if boolean else 101, boolean=boolean).statement()
    assert json.loads(ours(msgs)) == os_.stats(msgs).as_dict()


def test_metrics_and_valid_eval_metrics_verbs_host_side(tmp_path):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "zkinterface-ir_b200", "zkb")
    ws = os.path.join(root, "tests", "golden", "example")
    r = subprocess.run([cli, "metrics", ws], capture_output=True, timeout=60)
    assert r.returncode == 0
    msgs = [m for n in sorted(os.listdir(ws)) for m in F.read_messages(open(os.path.join(ws, n), "rb").read())]
    assert json.loads(r.stdout) == os_.stats(msgs).as_dict()
    assert r.stdout.endswith(b"}\n")
