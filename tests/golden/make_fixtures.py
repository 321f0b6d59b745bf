#!/usr/bin/env python
"""Regenerates the `.sieve` workspaces under tests/golden/ from the oracle's restatement of the reference's
example statements (oracle/fixtures.py <- rust/src/producers/{examples,boolean_examples,builder,from_r1cs}.rs),
serialised with oracle/sieve_fbs.py using the reference's file naming (producers/sink.rs:84-100).
The three files directly in tests/golden/ are the reference's own binary fixtures (rust/examples/*.sieve)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import fixtures as fx  # noqa: E402
from oracle import ir  # noqa: E402
from oracle import sieve_fbs as F  # noqa: E402

WORKSPACES = {
    "example": [fx.example_instance(), fx.example_witness(), fx.example_relation()],
    "example_incorrect": [fx.example_instance(), fx.example_witness_incorrect(), fx.example_relation()],
    "boolean_example": [fx.boolean_example_instance(), fx.boolean_example_witness(), fx.boolean_example_relation()],
    "boolean_example_incorrect": [fx.boolean_example_instance(), fx.boolean_example_witness_incorrect(),
                                  fx.boolean_example_relation()],
    "builder_switch": fx.builder_switch(),
    "builder_switch_nested": fx.builder_switch_nested_in_function(),
    "r1cs_example": fx.r1cs_to_gates(*fx.zkif_example())[0],
}
NAMES = {ir.Instance: "000_instance.sieve", ir.Witness: "001_witness.sieve", ir.Relation: "002_relation.sieve"}

if __name__ == "__main__":
    for name, msgs in WORKSPACES.items():
        d = os.path.join(HERE, name)
        os.makedirs(d, exist_ok=True)
        for m in msgs:
            with open(os.path.join(d, NAMES[type(m)]), "ab" if False else "wb") as f:
                f.write(F.write_message(m))
        print(name, [NAMES[type(m)] for m in msgs])
