"""CPU oracle for the zkb hot path — TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference's satisfiability-checking
path (`zki_sieve evaluate` = `Evaluator<PlaintextBackend>`,
reference `rust/src/consumers/evaluator.rs`).  It exists to CHECK the CUDA
product path, never to serve it:

  * only `tests/` (incl. the measurement harness `tests/bench_configs.py`, which uses it as workload generator
    and checker), `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
    `--impl reference` legs may import or execute anything under `oracle/`;
  * nothing under `zkinterface-ir_b200/` imports it, and the product fails
    loudly when its CUDA library or a GPU is missing.

Parity pinning: the Rust reference cannot be built in this
image (no cargo/rustc), so the oracle is pinned against every golden vector
the reference's own tests hold for this path — see `tests/test_oracle_golden.py`:
`test_exponentiation` KATs, the example statement (TRUE / "Wire_9 ..."),
the boolean example, the four `GateBuilder` circuits, the R1CS example wire
values, the `Stats` gate counts and the binary `.sieve` fixtures shipped in
`rust/examples/`; the widened consumers are pinned on their own reference tests:
`flattening.py` on `test_validate_flattening` / `test_evaluate_flattening`
(tests/test_flatten.py), `validator.py` on the four `test_validator*` cases and
`test_is_probably_prime` (tests/test_validator.py), `stats.py` on `test_stats`
(tests/test_stats.py).

Modules
  ir.py            owned data model (mirror of rust/src/structs/*.rs)
  sieve_fbs.py     FlatBuffers reader + writer for sieve_ir.fbs (no flatc here)
  evaluator.py     Evaluator / ZKBackend / PlaintextBackend restatement
  fixtures.py      the reference's example statements and builder circuits
  plaintext_flat.c C restatement of the PlaintextBackend gate loop for flat
                   circuits (big sizes, CPU baseline timing); flat.py binds it
  flattening.py    IRFlattener + the GateBuilder / MessageBuilder slice it drives
  validator.py     Validator restatement (violation texts character for character)
  stats.py         Stats restatement (the `metrics` JSON)
  workloads.py     the C5 nested-For boolean relation (generator)
"""
