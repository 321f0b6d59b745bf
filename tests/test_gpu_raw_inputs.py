"""SURVEY.md section 8a trap 1: instance / witness / constant values >= p stay RAW in the reference
(evaluator.rs:862-864, 896-898): AssertZero and Not test the raw integer, Add/Mul reduce.  The device resolves
this itself (raw flags written by the input kernel; and / xor re-read the raw bytes); every case is compared with the oracle."""
import numpy as np
import pytest

from oracle import evaluator as ev
from oracle import ir
from oracle import sieve_fbs as F
from tests.util import zkb

pytestmark = pytest.mark.gpu


def both(msgs):
    z = zkb()
    e = z.Evaluator.from_messages(z.Source.from_buffers([F.write_messages(msgs)]), device=0)
    return e.get_violations(), ev.evaluate(msgs), e


@pytest.mark.parametrize("p", [101, (1 << 64) - (1 << 32) + 1, 2])
def test_assert_and_not_on_unreduced_inputs(p):
    h = ir.Header(ir.le_bytes(p))
    boolean = p == 2
    mask = ir.BOOL if boolean else ir.ARITH
    addz = (lambda o, a: ("Xor", o, a, a)) if boolean else (lambda o, a: ("AddConstant", o, a, b"\x00"))
    cases = []
    for raw in (0, 1, p, 2 * p, p + 1, 3 * p):
        w = ir.le_bytes(raw)
        # assert directly on the witness: fails iff the RAW integer is non-zero
        cases.append(([("Witness", 0), ("AssertZero", 0)], [w], []))
        # assert after an arithmetic gate: reduced first
        cases.append(([("Witness", 0), addz(1, 0), ("AssertZero", 1)] if not boolean else
                      [("Witness", 0), ("Witness", 1), ("Xor", 2, 0, 1), ("AssertZero", 2)],
                      [w] if not boolean else [w, b"\x00"], []))
        # not(witness): 1 iff the RAW integer is zero; assert(not) fails iff raw == 0
        cases.append(([("Witness", 0), ("Not", 1, 0), ("AssertZero", 1)], [w], []))
        # through copies and as an instance
        cases.append(([("Instance", 0), ("Copy", 1, 0), ("Copy", 2, 1), ("Not", 3, 2), ("AssertZero", 3), ("AssertZero", 2)], [], [w]))
        # a constant >= p
        cases.append(([("Constant", 0, w), ("Not", 1, 0), ("AssertZero", 1)], [], []))
        cases.append(([("Constant", 0, w), ("AssertZero", 0)], [], []))
    for gates, wit, inst in cases:
        rel = ir.Relation(h, mask if not boolean else ir.BOOL, ir.SIMPLE, [], gates)
        msgs = [ir.Instance(h, inst), ir.Witness(h, wit), rel]
        got, want, e = both(msgs)
        assert got == want, (p, gates, wit, inst)


def test_batch_with_some_unreduced_witnesses():
    z = zkb()
    p = 101
    h = ir.Header(ir.le_bytes(p))
    gates = [("Witness", 0), ("Witness", 1), ("Add", 2, 0, 1), ("AssertZero", 2), ("Not", 3, 1), ("Mul", 4, 3, 0), ("AssertZero", 4),
             ("AssertZero", 1)]
    rel = ir.Relation(h, ir.ARITH | ir.NOT, ir.SIMPLE, [], gates)
    rows = [(0, 0), (101, 0), (0, 101), (5, 96), (202, 0), (0, 202), (3, 98), (101, 101)]
    b = z.GpuBackend(0)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_buffers([F.write_messages([ir.Witness(h, [ir.le_bytes(rows[0][0]), ir.le_bytes(rows[0][1])]), rel])]))
    assert e.get_violations() == []
    W = np.zeros((len(rows), 2, 4), dtype=np.uint8)
    for j, (x, y) in enumerate(rows):
        W[j, 0, 0], W[j, 1, 0] = x, y
    v = b.evaluate(None, W, len(rows))
    for j, (x, y) in enumerate(rows):
        want = ev.evaluate([ir.Witness(h, [ir.le_bytes(x), ir.le_bytes(y)]), rel])
        got = [] if v[j]["ok"] else [f"Wire_{b.assert_wire(int(v[j]['first_fail_seq']))} (may be weighted) should be 0, while it is not"]
        assert got == want, (j, x, y)


@pytest.mark.parametrize("p,stride", [(101, 4), (101, 12), ((1 << 64) - (1 << 32) + 1, 8), ((1 << 64) - (1 << 32) + 1, 24),
                                      (0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001, 32),
                                      (0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001, 40)])
def test_bitwise_gates_on_unreduced_inputs_use_the_raw_integers(p, stride):
    """`and` / `xor` over an odd field are (a & b) % m and (a ^ b) % m on the integers the reference holds
    (evaluator.rs:924-930): an instance / witness / constant >= p takes part UNREDUCED, whatever its width.  A batch mixes
    canonical and raw values per witness; every recorded value is compared with the oracle."""
    z = zkb()
    h = ir.Header(ir.le_bytes(p))
    rng = np.random.default_rng(stride)
    big = ir.le_bytes(p + 12345 if p > 101 else 3 * p + 2)          # a constant >= p
    gates = [("Witness", 0), ("Witness", 1), ("Instance", 2), ("Constant", 3, big), ("Constant", 4, b"\x07"),
             ("And", 5, 0, 1), ("Xor", 6, 0, 1), ("And", 7, 0, 3), ("Xor", 8, 3, 1), ("And", 9, 2, 4), ("Xor", 10, 2, 2),
             ("Mul", 11, 0, 1), ("And", 12, 11, 0), ("Xor", 13, 1, 11), ("Copy", 14, 0), ("Xor", 15, 14, 3), ("And", 16, 3, 3),
             ("Add", 17, 5, 6), ("AssertZero", 10), ("Xor", 18, 4, 3)]
    rel = ir.Relation(h, ir.ARITH | ir.BOOL, ir.SIMPLE, [], gates)
    n = 40
    W = rng.integers(0, 256, size=(n, 2, stride), dtype=np.uint8)
    I = rng.integers(0, 256, size=(n, 1, stride), dtype=np.uint8)
    eb = (p.bit_length() + 7) // 8
    for j in range(0, n, 3):        # canonical values in a third of the witnesses
        for arr, k in ((W, 0), (W, 1), (I, 0)):
            v = int.from_bytes(arr[j, k].tobytes(), "little") % p
            arr[j, k] = np.frombuffer(v.to_bytes(stride, "little"), dtype=np.uint8)
    W[1, 0] = 0
    W[2, 1, :] = np.frombuffer(p.to_bytes(stride, "little"), dtype=np.uint8) if stride >= eb else W[2, 1]
    b = z.GpuBackend(0)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_buffers([F.write_messages([ir.Instance(h, [I[0, 0].tobytes()]),
                                                             ir.Witness(h, [W[0, 0].tobytes(), W[0, 1].tobytes()]), rel])]))
    b.finalize(True)
    v = b.evaluate(I, W, n)
    n_vals = b.stats()["n_values"]
    for j in range(n):
        msgs = [ir.Instance(h, [I[j, 0].tobytes()]), ir.Witness(h, [W[j, 0].tobytes(), W[j, 1].tobytes()]), rel]
        tb = ev.TracingBackend()
        o = ev.Evaluator.from_messages(msgs, tb)
        assert o.get_violations() == []          # x ^ x == 0 for the raw integer too
        assert v[j]["ok"] == 1
        vals = [val for (k, _, val) in tb.trace if k != "copy"]
        assert len(vals) == n_vals
        assert b.read_values(j, list(range(n_vals)), 64) == vals, j


def test_bitwise_gate_on_unreduced_input_through_the_evaluator():
    z = zkb()
    p = 101
    h = ir.Header(ir.le_bytes(p))
    rel = ir.Relation(h, ir.ARITH | ir.BOOL, ir.SIMPLE, [], [("Witness", 0), ("Witness", 1), ("And", 2, 0, 1), ("AssertZero", 2)])
    for w0 in (3, 101 + 3, 101 + 4, 4, 2 * 101 + 4):     # (w0 & 4) % 101
        msgs = [ir.Witness(h, [ir.le_bytes(w0), b"\x04"]), rel]
        e = z.Evaluator.from_messages(z.Source.from_buffers([F.write_messages(msgs)]), device=0)
        assert e.get_violations() == ev.evaluate(msgs), w0


@pytest.mark.parametrize("p,stride", [(101, 8), ((1 << 64) - (1 << 32) + 1, 20), (0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, 48)])
def test_witness_values_wider_than_the_field_element(p, stride):
    """`Value`s are byte strings of any length; the reference reduces them at the first Add/Mul.  The input
    kernel reduces arbitrarily wide raw values (Horner over element-sized chunks)."""
    z = zkb()
    h = ir.Header(ir.le_bytes(p))
    rng = np.random.default_rng(9)
    gates = [("Witness", 0), ("Witness", 1), ("Mul", 2, 0, 1), ("Add", 3, 2, 0), ("MulConstant", 4, 3, ir.le_bytes(p - 1)),
             ("Add", 5, 4, 3), ("AssertZero", 5), ("AssertZero", 1)]
    rel = ir.Relation(h, ir.ARITH, ir.SIMPLE, [], gates)
    n = 12
    W = rng.integers(0, 256, size=(n, 2, stride), dtype=np.uint8)
    W[3, 1] = 0                      # raw zero: the direct assertion holds
    W[5, 1, :] = np.frombuffer((p * 7).to_bytes(stride, "little"), dtype=np.uint8)   # 0 mod p but non-zero: fails
    b = z.GpuBackend(0)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_buffers([F.write_messages([ir.Witness(h, [W[0, 0].tobytes(), W[0, 1].tobytes()]), rel])]))
    b.finalize(True)
    e.get_violations()
    v = b.evaluate(None, W, n)
    for j in range(n):
        msgs = [ir.Witness(h, [W[j, 0].tobytes(), W[j, 1].tobytes()]), rel]
        tb = ev.TracingBackend()
        o = ev.Evaluator.from_messages(msgs, tb)
        want = o.get_violations()
        got = [] if v[j]["ok"] else [f"Wire_{b.assert_wire(int(v[j]['first_fail_seq']))} (may be weighted) should be 0, while it is not"]
        assert got == want, j
        # computed values (reduced by the first gate) and the raw inputs themselves
        vals = b.read_values(j, [2, 3, 4, 5], 64)
        x, y = int.from_bytes(W[j, 0].tobytes(), "little"), int.from_bytes(W[j, 1].tobytes(), "little")
        m = x * y % p
        assert vals == [m, (m + x) % p, (m + x) * (p - 1) % p, 0]
        assert b.read_values(j, [0, 1], 64) == [x, y]
