"""Pins the C oracle (oracle/plaintext_flat.c, the checker behind smoke(), tests/test_gpu_flat.py,
tests/test_gpu_raw_inputs.py, tests/test_gpu_full_size.py and bench.py's cpu_baseline) before anything trusts it:

  1. its big-integer kernel (`bn_mul` + Knuth-D `bn_mod`, exported as `flat_bn_mulmod`) against the reference's own
     known answers — both `test_exponentiation` KATs (rust/src/consumers/evaluator.rs:955-970), driven through the same
     left-to-right square-and-multiply as `exp` (:801-820) — and against Python integers on random operands for ten
     moduli (1 to 8 digits, top bit set or not, operands reduced or not);
  2. its gate loop (`flat_eval_one`) against oracle/evaluator.py — the line-for-line Python restatement of
     `Evaluator<PlaintextBackend>` that tests/test_oracle_golden.py pins on the reference's golden vectors — on random
     flat programs: every gate kind, Free + wire-id re-use, instance / witness values >= p, constants >= p, and every
     status the gate loop can end with (assertion failed, no value, already set, not enough instance, missing witness).
"""
import numpy as np
import pytest

from oracle import evaluator as oev
from oracle import flat, ir
from tests.util import FIELDS, circuits, random_flat_program

MODULI = [
    101, (1 << 31) - 1, (1 << 32) - 5, (1 << 61) - 1, (1 << 64) - (1 << 32) + 1,
    16249742125730185677094195492597105093,            # evaluator.rs:956 (124 bits)
    FIELDS["bn254"], FIELDS["bls381"], (1 << 256) - 189,  # 8 digits, top bit set
    (1 << 255) + 95,                                   # 8 digits, top digit 0x80000000 (normalisation shift 0)
]


def _exp_with(mulmod, base, exponent, modulus):
    """`exp` (evaluator.rs:801-820): left-to-right square-and-multiply, recursion on exponent >> 1"""
    if exponent == 1:
        return base % modulus if base >= modulus else base
    sq = _exp_with(mulmod, base, exponent >> 1, modulus)
    sq = mulmod(sq, sq, modulus)
    if exponent & 1:
        sq = mulmod(sq, base, modulus)
    return sq


def test_reference_exponentiation_kats_through_the_c_bignum():
    # rust/src/consumers/evaluator.rs:955-970
    assert _exp_with(flat.mulmod, 2, 2206000150907221872269901214599500635,
                     16249742125730185677094195492597105093) == 5834907326474057072663503101785122138
    assert _exp_with(flat.mulmod, 42, 100, 101) == 1


@pytest.mark.parametrize("m", MODULI)
def test_mulmod_against_python_integers(m):
    rng = np.random.default_rng(m % (1 << 32))
    nb = (m.bit_length() + 7) // 8
    edge = [0, 1, 2, m - 1, m - 2, m, m + 1, (1 << (8 * nb)) - 1, 1 << (m.bit_length() - 1)]
    for a in edge:
        for b in edge:
            assert flat.mulmod(a, b, m) == a * b % m, (a, b, m)
    for it in range(2000):
        # operands below p, and raw operands wider than p (the reference keeps inputs unreduced, :862-864)
        wide = it % 4 == 3
        a = int.from_bytes(rng.bytes(nb + (8 if wide else 0)), "little")
        b = int.from_bytes(rng.bytes(nb), "little")
        if not wide:
            a %= m
            b %= m
        assert flat.mulmod(a, b, m) == a * b % m, (a, b, m)


# ---------------------------------------------------------------------------------------------
# gate loop vs the Python restatement of Evaluator<PlaintextBackend>
# ---------------------------------------------------------------------------------------------
def _to_ir_gates(c, gates, pool):
    out = []
    for g in gates:
        op, o, a, b = int(g["op"]), int(g["out"]), int(g["a"]), int(g["b"])
        if op == c.G_CONSTANT:
            out.append(("Constant", o, pool[b].tobytes()))
        elif op == c.G_ASSERT_ZERO:
            out.append(("AssertZero", a))
        elif op == c.G_COPY:
            out.append(("Copy", o, a))
        elif op in (c.G_ADD, c.G_MUL, c.G_AND, c.G_XOR):
            out.append(({c.G_ADD: "Add", c.G_MUL: "Mul", c.G_AND: "And", c.G_XOR: "Xor"}[op], o, a, b))
        elif op in (c.G_ADD_CONSTANT, c.G_MUL_CONSTANT):
            out.append(("AddConstant" if op == c.G_ADD_CONSTANT else "MulConstant", o, a, pool[b].tobytes()))
        elif op == c.G_NOT:
            out.append(("Not", o, a))
        elif op == c.G_INSTANCE:
            out.append(("Instance", o))
        elif op == c.G_WITNESS:
            out.append(("Witness", o))
        elif op == c.G_FREE:
            out.append(("Free", a, b))
        else:
            raise AssertionError(op)
    return out


def _python_oracle(c, p, gates, pool, inst, wit):
    """(violations, live top-scope values, panicked) from oracle/evaluator.py"""
    hdr = ir.Header(field_characteristic=ir.le_bytes(p))
    msgs = [ir.Instance(hdr, [v.tobytes() for v in inst] if inst is not None else []),
            ir.Witness(hdr, [v.tobytes() for v in wit] if wit is not None else []),
            # gate_mask 0x000F = arithmetic: with @and,@xor,@not all set is_boolean would flip (evaluator.rs:262); the
            # flat loop has no Switch, so is_boolean never matters for it
            ir.Relation(hdr, 0x000F, 0x1000, [], _to_ir_gates(c, gates, pool))]
    be = oev.PlaintextBackend()
    try:
        ev = oev.Evaluator.from_messages(msgs, be)
    except ir.OraclePanic:
        return None, None, True
    return ev.get_violations(), dict(ev.values), False


def _compare(c, p, gates, pool, inst, wit, n_wires):
    eb = pool.shape[1]
    res, dump = flat.eval_dump(gates, pool, ir.le_bytes(p, eb), inst, wit, n_wires, stride=eb + 8)
    viol, values, panicked = _python_oracle(c, p, gates, pool, inst, wit)
    if panicked:
        assert int(res["status"]) == flat.EV_MISSING_WITNESS_PANIC
        return int(res["status"])
    assert flat.violation_text(res) == viol, (flat.violation_text(res), viol)
    if int(res["status"]) == flat.EV_ASSERT_FAILED:
        # which assertion (program order) — the index the GPU path reports — and how far the loop got
        k = [i for i, g in enumerate(gates) if g["op"] == c.G_ASSERT_ZERO][int(res["fail_assert_seq"])]
        assert int(res["gates_done"]) == k and int(gates[k]["a"]) == int(res["fail_wire"])
    # the wire store at the end (or at the point of the error): same live ids, same RAW integers
    for w in range(n_wires):
        got = dump[w]
        if w in values:
            assert int.from_bytes(got.tobytes(), "little") == values[w], (w, values[w])
        else:
            assert (got == 0xFF).all(), w
    return int(res["status"])


@pytest.mark.parametrize("name", ["p101", "goldilocks", "kat124", "bls381", "p256full"])
def test_gate_loop_against_python_evaluator_random_programs(name):
    c = circuits()
    p = FIELDS[name]
    eb = c.elem_bytes(p)
    seen = set()
    for seed in range(12):
        n_inst, n_wit = 3, 6
        gates, pool, n_wires = random_flat_program(p, 220, n_inst, n_wit, seed=100 * seed + 7, bool_ops=True)
        rng = np.random.default_rng(seed)
        # values: canonical, plus RAW inputs >= p that the reference keeps unreduced (evaluator.rs:862-864, 896-898)
        def val(i):
            v = int.from_bytes(rng.bytes(eb), "little")
            if (seed + i) % 3:
                v %= p
            return np.frombuffer(v.to_bytes(eb, "little"), dtype=np.uint8)
        inst = np.stack([val(i) for i in range(n_inst)])
        wit = np.stack([val(10 + i) for i in range(n_wit)])
        if seed % 4 == 1:   # a constant >= p too
            pool = pool.copy()
            pool[3] = np.frombuffer(((1 << (8 * eb)) - 1).to_bytes(eb, "little"), dtype=np.uint8)
        seen.add(_compare(c, p, gates, pool, inst, wit, n_wires))
    assert flat.EV_TRUE in seen or flat.EV_ASSERT_FAILED in seen


def _g(c, rows):
    g = np.zeros(len(rows), dtype=c.GATE_DTYPE)
    for i, (op, out, a, b) in enumerate(rows):
        g[i]["op"], g[i]["out"], g[i]["a"], g[i]["b"] = op, out, a, b
    return g


def test_every_status_code():
    c = circuits()
    p = FIELDS["bn254"]
    eb = 32
    pool = np.stack([c.le_bytes(v, eb) for v in (0, 1, p - 1, p, p + 5)])
    one = np.stack([c.le_bytes(7, eb)])
    cases = {
        flat.EV_TRUE: [(c.G_WITNESS, 0, 0, 0), (c.G_MUL_CONSTANT, 1, 0, 2), (c.G_ADD, 2, 0, 1), (c.G_ASSERT_ZERO, 0, 2, 0)],
        flat.EV_ASSERT_FAILED: [(c.G_WITNESS, 0, 0, 0), (c.G_ASSERT_ZERO, 0, 0, 0)],
        flat.EV_NO_VALUE: [(c.G_WITNESS, 0, 0, 0), (c.G_ADD, 1, 0, 5)],
        flat.EV_ALREADY_SET: [(c.G_WITNESS, 0, 0, 0), (c.G_COPY, 0, 0, 0)],
        flat.EV_NOT_ENOUGH_INSTANCE: [(c.G_INSTANCE, 0, 0, 0), (c.G_INSTANCE, 1, 0, 0)],
        flat.EV_MISSING_WITNESS_PANIC: [(c.G_WITNESS, 0, 0, 0), (c.G_WITNESS, 1, 0, 0)],
    }
    for want, rows in cases.items():
        got = _compare(c, p, _g(c, rows), pool, one, one, 8)
        assert got == want, (want, got)
    # Free of a dead wire / double Free / re-use after Free
    assert _compare(c, p, _g(c, [(c.G_WITNESS, 0, 0, 0), (c.G_FREE, 0, 0, 0), (c.G_FREE, 0, 0, 0)]), pool, one, one, 4) == flat.EV_NO_VALUE
    assert _compare(c, p, _g(c, [(c.G_WITNESS, 3, 0, 0), (c.G_FREE, 0, 3, 3), (c.G_CONSTANT, 3, 0, 1), (c.G_NOT, 0, 3, 0),
                                 (c.G_ASSERT_ZERO, 0, 0, 0)]), pool, one, one, 4) == flat.EV_TRUE
    # Free(first, last) with last < first frees nothing (evaluator.rs:434-438)
    assert _compare(c, p, _g(c, [(c.G_WITNESS, 2, 0, 0), (c.G_FREE, 0, 2, 1)]), pool, one, one, 4) == flat.EV_TRUE


def test_unreduced_values_keep_their_raw_meaning():
    """SURVEY.md section 8a trap 1: constant / instance / witness / copy keep the RAW integer; assert_zero and not test
    it; add / mul / and / xor reduce (evaluator.rs:862-864, 896-906, 924-938)"""
    c = circuits()
    p = FIELDS["goldilocks"]
    eb = 16
    pool = np.stack([c.le_bytes(v, eb) for v in (p, 2 * p, 0, 1)])
    wit = np.stack([c.le_bytes(p, eb), c.le_bytes(p + 6, eb)])
    # a witness equal to p is 0 mod p but FAILS AssertZero
    assert _compare(c, p, _g(c, [(c.G_WITNESS, 0, 0, 0), (c.G_ASSERT_ZERO, 0, 0, 0)]), pool, None, wit, 4) == flat.EV_ASSERT_FAILED
    # ... its copy too, while p + 0 passes (the addition reduces)
    assert _compare(c, p, _g(c, [(c.G_WITNESS, 0, 0, 0), (c.G_COPY, 1, 0, 0), (c.G_ADD_CONSTANT, 2, 1, 2), (c.G_ASSERT_ZERO, 0, 2, 0),
                                 (c.G_ASSERT_ZERO, 0, 1, 0)]), pool, None, wit, 4) == flat.EV_ASSERT_FAILED
    # not(p) = 0, not(constant 2p) = 0, and/xor act on the raw integers and reduce afterwards
    assert _compare(c, p, _g(c, [(c.G_WITNESS, 0, 0, 0), (c.G_WITNESS, 1, 0, 0), (c.G_NOT, 2, 0, 0), (c.G_ASSERT_ZERO, 0, 2, 0),
                                 (c.G_CONSTANT, 3, 0, 1), (c.G_NOT, 4, 3, 0), (c.G_ASSERT_ZERO, 0, 4, 0), (c.G_AND, 5, 0, 1),
                                 (c.G_XOR, 6, 0, 1), (c.G_XOR, 7, 3, 1)]), pool, None, wit, 8) == flat.EV_TRUE
