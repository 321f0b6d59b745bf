"""CPU restatement of `rust/src/consumers/stats.rs` (oracle, test infrastructure): the gate / message counts the
`metrics` and `valid-eval-metrics` verbs print (cli.rs:322-363).  Pinned on the reference's `test_stats`
(stats.rs:288-345) in tests/test_stats.py."""
from __future__ import annotations

import json
from typing import Dict

from . import ir

GATE_FIELDS = ["instance_variables", "witness_variables", "constants_gates", "assert_zero_gates", "copy_gates", "add_gates",
               "mul_gates", "add_constant_gates", "mul_constant_gates", "and_gates", "xor_gates", "not_gates", "variables_freed",
               "functions_defined", "functions_called", "switches", "branches", "for_loops", "instance_messages",
               "witness_messages", "relation_messages"]                 # stats.rs:11-41, declaration order
_SIMPLE = {"Constant": "constants_gates", "AssertZero": "assert_zero_gates", "Copy": "copy_gates", "Add": "add_gates",
           "Mul": "mul_gates", "AddConstant": "add_constant_gates", "MulConstant": "mul_constant_gates", "And": "and_gates",
           "Xor": "xor_gates", "Not": "not_gates", "Instance": "instance_variables", "Witness": "witness_variables"}
_CALL_FIELDS = ["constants_gates", "assert_zero_gates", "copy_gates", "add_gates", "mul_gates", "add_constant_gates",
                "mul_constant_gates", "and_gates", "xor_gates", "not_gates", "variables_freed", "switches", "branches", "for_loops",
                "functions_called"]                                     # ingest_call_stats, stats.rs:268-286


def new_gate_stats() -> Dict[str, int]:
    return {k: 0 for k in GATE_FIELDS}


def ingest_subcircuit(sub, known) -> Dict[str, int]:                    # stats.rs:114-123
    local = new_gate_stats()
    for g in sub:
        ingest_gate(local, g, known)
    return local


def _call_stats(st, other):
    for k in _CALL_FIELDS:
        st[k] += other[k]


def _named_call(st, name, known):
    st["functions_called"] += 1
    if name in known:
        fs, ic, wc = known[name]
        _call_stats(st, fs)
        return ic, wc
    return None                                                           # "WARNING Stats: function not defined"


def ingest_gate(st, gate, known):                                       # stats.rs:126-266
    k = gate[0]
    if k in _SIMPLE:
        st[_SIMPLE[k]] += 1
    elif k == "Free":
        first, last = gate[1], gate[2]
        st["variables_freed"] += ((first if last is None else last) - first + 1)
    elif k == "Call":
        r = _named_call(st, gate[1], known)
        if r:
            st["instance_variables"] += r[0]
            st["witness_variables"] += r[1]
    elif k == "AnonCall":
        _call_stats(st, ingest_subcircuit(gate[5], known))
        st["instance_variables"] += gate[3]
        st["witness_variables"] += gate[4]
    elif k == "Switch":
        st["switches"] += 1
        st["branches"] += len(gate[4])
        mi = mw = 0
        for br in gate[4]:
            if br[0] == "AbstractGateCall":
                ic, wc = _named_call(st, br[1], known) or (0, 0)
            else:
                _call_stats(st, ingest_subcircuit(br[4], known))
                ic, wc = br[2], br[3]
            mi, mw = max(mi, ic), max(mw, wc)
        st["instance_variables"] += mi
        st["witness_variables"] += mw
    elif k == "For":
        st["for_loops"] += 1
        body = gate[5]
        for _ in range(gate[2], gate[3] + 1):
            if body[0] == "IterExprCall":
                r = _named_call(st, body[1], known)
                if r:
                    st["instance_variables"] += r[0]
                    st["witness_variables"] += r[1]
            else:
                _call_stats(st, ingest_subcircuit(body[5], known))
                st["instance_variables"] += body[3]
                st["witness_variables"] += body[4]
    else:
        raise ValueError(k)


class Stats:
    def __init__(self):
        self.field_characteristic = b""
        self.field_degree = 0
        self.gate_stats = new_gate_stats()
        self.functions = {}

    def ingest_message(self, m):                                         # stats.rs:61-112
        self.field_characteristic = bytes(m.header.field_characteristic)
        self.field_degree = m.header.field_degree
        if isinstance(m, ir.Instance):
            self.gate_stats["instance_messages"] += 1
        elif isinstance(m, ir.Witness):
            self.gate_stats["witness_messages"] += 1
        else:
            self.gate_stats["relation_messages"] += 1
            for f in m.functions:
                self.gate_stats["functions_defined"] += 1
                self.functions[f.name] = (ingest_subcircuit(f.body, self.functions), f.instance_count, f.witness_count)
            for g in m.gates:
                ingest_gate(self.gate_stats, g, self.functions)

    def as_dict(self):
        """the shape serde_json gives `Stats` (tuples become arrays, Vec<u8> an array of numbers)"""
        return {"field_characteristic": list(self.field_characteristic), "field_degree": self.field_degree,
                "gate_stats": dict(self.gate_stats),
                "functions": {k: [dict(v[0]), v[1], v[2]] for k, v in self.functions.items()}}

    def to_json_pretty(self) -> str:                                     # serde_json::to_writer_pretty: two-space indent
        return json.dumps(self.as_dict(), indent=2)


def stats(messages) -> Stats:
    s = Stats()
    for m in messages:
        s.ingest_message(m)
    return s
