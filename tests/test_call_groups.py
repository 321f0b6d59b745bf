"""Loop-structured recording (csrc/program.h CallGroup): a For loop over a
plain named function is kept as ONE descriptor and expanded on the device.  Everything the reference's gate-by-gate
unrolling (rust/src/consumers/evaluator.rs:495-559 + :441-471 + :698-746) makes observable must stay the same:
callback counts, verdicts, violation texts, every wire value of the enclosing scope.

Host part (no GPU): state equality of the grouped recording against the serial one and against the oracle's trace.
GPU part: values and verdicts against oracle/evaluator.py on random relations built to hit every operand form
(affine slots, slot tables, stride 0, groups reading groups), the refusals (Not on an input, AssertZero / Copy in the
body, unreduced constants, loop-carried inputs) and raw witness values >= 2."""
import os

import numpy as np
import pytest

from oracle import evaluator as ev
from oracle import ir
from oracle import sieve_fbs as F
from oracle import workloads as wl
from tests.util import zkb

I = lambda n: ("Name", n)
C = lambda v: ("Const", v)
add = lambda l, r: ("Add", l, r)
mul = lambda l, r: ("Mul", l, r)


class Rel:
    """random Boolean relation: n_wit witness bits, a few plain functions, For loops over them, a tail of checks"""

    def __init__(self, seed, n_wit=96, raw_witness=False, allow_refused=True):
        self.rng = np.random.default_rng(seed)
        self.h = ir.Header(bytes([2]))
        self.n_wit = n_wit
        self.raw = raw_witness
        self.funcs = []
        self.gates = []
        self.next = 0
        self.blocks = []   # (first wire, count) of defined, never freed wires
        self.allow_refused = allow_refused
        self.expect_groups = 0

    def ri(self, lo, hi):
        return int(self.rng.integers(lo, hi + 1))

    def function(self, name, n_out, n_in, kind):
        """kind: 'plain' qualifies; the others must be refused by the recorder and run gate by gate"""
        g, avail = [], list(range(n_out, n_out + n_in))
        nxt = n_out + n_in
        computed = []
        n_body = self.ri(max(n_out, 2), 14)
        targets = [None] * (n_body - n_out) + list(range(n_out))
        for k, t in enumerate(targets):
            w = t if t is not None else nxt
            if t is None:
                nxt += 1
            op = ["Xor", "And", "Not", "Add", "Mul", "AddConstant", "MulConstant", "Constant"][self.ri(0, 7)]
            a, b = avail[self.ri(0, len(avail) - 1)], avail[self.ri(0, len(avail) - 1)]
            if op == "Not":
                if not computed:
                    op = "Xor"
                else:
                    a = computed[self.ri(0, len(computed) - 1)]
            if op in ("Xor", "And", "Add", "Mul"):
                g.append((op, w, a, b))
            elif op == "Not":
                g.append((op, w, a))
            elif op in ("AddConstant", "MulConstant"):
                g.append((op, w, a, bytes([self.ri(0, 3)])))     # constants of an AddConstant / MulConstant may be >= 2
            else:
                g.append((op, w, bytes([self.ri(0, 1)])))
            avail.append(w)
            computed.append(w)
        if kind == "not_on_input":
            g.insert(0, ("Not", nxt, n_out))
            nxt += 1
        elif kind == "assert":
            g.append(("Xor", nxt, 0, 0))
            g.append(("AssertZero", nxt))
            nxt += 1
        elif kind == "copy":
            g.insert(0, ("Copy", nxt, n_out))
            nxt += 1
        elif kind == "raw_const":
            g.insert(0, ("Constant", nxt, bytes([2])))
            g.insert(1, ("Not", nxt + 1, nxt))
            nxt += 2
        self.funcs.append(ir.Function(name, n_out, n_in, 0, 0, g))
        return (name, n_out, n_in, kind)

    def witnesses(self):
        n = self.n_wit
        self.gates.append(("For", "k", 0, n - 1, [ir.WireRange(0, n - 1)],
                           ("IterExprAnonCall", [("Single", I("k"))], [], 0, 1, [("Witness", 0)])))
        self.next = n
        self.blocks.append((0, n))

    def loop(self, f, n_calls=None, carried=False, allow_div=True):
        name, n_out, n_in, kind = f
        n = n_calls if n_calls is not None else self.ri(4, 40)
        base = self.next
        S = n_out + self.ri(0, 2)
        if carried:     # call 0 reads a seed wire placed where "call -1" would have written its output 0
            self.gates.append(("Xor", base, 0, 1))
            base += S
        # outputs: call c writes base + S*c + (0 .. n_out-1); S > n_out leaves holes
        if n_out > 1 and self.ri(0, 1):
            outs = [("Range", add(C(base), mul(I("i"), C(S))), add(C(base + n_out - 1), mul(C(S), I("i"))))]
        else:
            outs = [("Single", add(C(base + k), mul(I("i"), C(S)))) for k in range(n_out)]
        ins = []
        for k in range(n_in):
            first, count = self.blocks[self.ri(0, len(self.blocks) - 1)]
            st = self.ri(0, 3)
            while st * (n - 1) >= count:
                st -= 1
            off = self.ri(0, count - 1 - st * (n - 1))
            if carried and k == 0:      # reads the previous call's output 0: cannot be a group
                ins.append(("Single", add(C(base - S), mul(I("i"), C(S)))))
            elif st == 0:
                ins.append(("Single", C(first + off)))
            elif allow_div and self.ri(0, 3) == 0 and st == 2:
                ins.append(("Single", add(C(first + off), ("DivConst", mul(I("i"), C(4)), 2))))   # not affine: per-gate path
                kind = "divconst"
            else:
                ins.append(("Single", add(C(first + off), mul(I("i"), C(st)))))
        self.gates.append(("For", "i", 0, n - 1, [ir.WireRange(base, base + S * (n - 1) + n_out - 1)],
                           ("IterExprCall", name, outs, ins)))
        self.next = base + S * n
        if S == n_out:
            self.blocks.append((base, S * n))
        else:
            self.blocks.append((base, n_out))
        if kind == "plain" and not carried and n >= 4:
            self.expect_groups += 1

    def tail(self, n_checks=6, failing=False):
        for _ in range(n_checks):
            first, count = self.blocks[self.ri(0, len(self.blocks) - 1)]
            t = first + self.ri(0, count - 1)
            s = self.next
            self.next += 1
            self.gates.append(("Xor", s, t, t))
            self.gates.append(("AssertZero", s))
        first, count = self.blocks[-1]
        u = self.next
        self.gates.append(("And", u, first, first + count - 1))
        self.gates.append(("Not", u + 1, first))           # Not directly on a group output
        self.gates.append(("Xor", u + 2, u + 1, u + 1))
        self.gates.append(("AssertZero", u + 2))
        self.next += 3
        if failing:
            self.gates.append(("AssertZero", first + count // 2))   # data dependent, directly on a loop output

    def messages(self, w):
        rel = ir.Relation(self.h, ir.BOOL | ir.ARITH, ir.FOR | ir.FUNCTION, self.funcs, self.gates)
        return [ir.Witness(self.h, [bytes([int(x)]) for x in w]), rel]

    def witness(self):
        if self.raw:
            return self.rng.integers(0, 5, self.n_wit).astype(np.uint8)   # raw integers up to 4 (trap 1)
        return self.rng.integers(0, 2, self.n_wit).astype(np.uint8)


def build(seed, raw=False, failing=False):
    r = Rel(seed, raw_witness=raw)
    kinds = ["plain", "plain", "plain", "not_on_input", "assert", "copy", "raw_const"]
    fs = [r.function(f"f{k}", r.ri(1, 3), r.ri(1, 8) if k % 2 else r.ri(1, 4), kinds[k % len(kinds)] if k else "plain") for k in range(r.ri(2, 5))]
    r.witnesses()
    r.loop(fs[0], allow_div=False)
    for k in range(r.ri(2, 6)):
        f = fs[r.ri(0, len(fs) - 1)]
        r.loop(f, n_calls=3 if r.ri(0, 9) == 0 else None, carried=(r.ri(0, 7) == 0 and f[2] > 0))
    r.tail(failing=failing)
    return r


def record(device, msgs, no_groups=False):
    z = zkb()
    if no_groups:
        os.environ["ZKB_NO_CALL_GROUPS"] = "1"
    try:
        b = z.GpuBackend(device)
        e = z.Evaluator(b)
        e.ingest_source(z.Source.from_buffers([F.write_messages(msgs)]))
    finally:
        os.environ.pop("ZKB_NO_CALL_GROUPS", None)
    return b, e


SEEDS = list(range(24))


@pytest.mark.parametrize("seed", SEEDS)
def test_grouped_recording_keeps_the_observable_state(seed):
    r = build(seed)
    msgs = r.messages(r.witness())
    bg, eg = record(-1, msgs)
    bs, es = record(-1, msgs, no_groups=True)
    sg, ss = bg.stats(), bs.stats()
    assert ss["n_call_groups"] == 0
    assert 1 <= sg["n_call_groups"] <= r.expect_groups      # a loop that reads the outputs of a refused loop is refused too
    for k in ("n_asserts", "n_instance", "n_witness", "ir_gates", "callbacks"):
        assert sg[k] == ss[k], k
    assert sg["n_values"] < ss["n_values"]
    assert bg.pending_error() is None and bs.pending_error() is None
    # the oracle's trace asks the backend for the same callbacks
    tb = ev.TracingBackend()
    ev.Evaluator.from_messages(msgs, tb)
    oc = tb.counts()
    for k, v in sg["callbacks"].items():
        assert v == oc.get(k, 0), (k, v, oc.get(k, 0))
    # the same wires are bound in the top-level scope
    for w in range(r.next):
        hg = hs = None
        try:
            hg = eg.value_handle(w)
        except Exception:
            pass
        try:
            hs = es.value_handle(w)
        except Exception:
            pass
        assert (hg is None) == (hs is None), w
    # asserted values: same assertion wires, in order
    for s in range(sg["n_asserts"]):
        assert bg.assert_wire(s) == bs.assert_wire(s)


def test_the_seeds_cover_every_operand_form():
    depths, tables, groups = 0, 0, 0
    for seed in SEEDS:
        r = build(seed)
        b, e = record(-1, r.messages(r.witness()))
        b.finalize(True)
        st = b.stats()
        groups += st["n_call_groups"]
        depths = max(depths, st["n_group_launches"])
        tables += st["n_group_table_slots"]
    assert groups >= 2 * len(SEEDS) and depths >= 3, (groups, depths)
    r, _ = table_relation()
    b, e = record(-1, r.messages(r.witness()))
    b.finalize(True)
    assert b.stats()["n_call_groups"] == 2 and b.stats()["n_group_table_slots"] > 0


def table_relation(seed=5, n=40):
    """even wires are Witness gates, odd wires irregularly a Witness or an Xor: the handles of the even wires are an
    arithmetic progression, their slots are not (input slots are numbered over the inputs only) -> explicit slot tables"""
    r = Rel(seed)
    f = r.function("f0", 2, 3, "plain")
    n_w = 0
    for k in range(2 * n):
        if k % 2 == 0 or k < 3 or r.ri(0, 1):
            r.gates.append(("Witness", k))
            n_w += 1
        else:
            r.gates.append(("Xor", k, k - 2, k - 1))
    r.n_wit = n_w
    r.next = 2 * n
    for base_in in (0, 2):
        base = r.next
        r.gates.append(("For", "i", 0, n - 4, [ir.WireRange(base, base + 2 * (n - 3) - 1)],
                        ("IterExprCall", "f0", [("Range", add(C(base), mul(I("i"), C(2))), add(C(base + 1), mul(I("i"), C(2))))],
                         [("Single", add(C(base_in), mul(I("i"), C(2)))), ("Single", add(C(base_in + 4), mul(I("i"), C(2)))), ("Single", C(6))])))
        r.next = base + 2 * (n - 3)
        r.blocks.append((base, 2 * (n - 3)))
    r.tail(failing=True)
    return r, n_w


def test_specialised_group_kernel_is_generated_and_compiles(monkeypatch):
    """group_jit.cpp without a device: the templates of random relations as straight-line CUDA, compiled by NVRTC for
    sm_100a in the calling thread (ZKB_GROUP_JIT=2); every template of the plan is a case of the generated switch"""
    monkeypatch.setenv("ZKB_GROUP_JIT", "2")
    for seed in (0, 3, 8):
        r = build(seed)
        b, e = record(-1, r.messages(r.witness()))
        b.finalize(True)
        state, seconds = b.wait_group_jit()
        if state == -1:
            pytest.skip("NVRTC is not available here")
        assert state == 2, state
        src = b.group_jit_source()
        assert src.count("case ") >= 1 and "zkb_groups_w4" in src and "__shared__" not in src


def test_c5_shape_is_one_group_per_inner_loop():
    rel, n_leaf = wl.boolean_for_relation(4, 5, 256)
    msgs = [ir.Witness(rel.header, [b"\0"] * 256), rel]
    b, e = record(-1, msgs)
    st = b.stats()
    assert st["n_call_groups"] == 16 and st["n_group_calls"] == 16 * 32
    assert st["ir_gates"] == n_leaf + 8
    assert st["n_values"] == 256 + 16 * 32 * 2 + 7
    b.finalize(False)
    assert b.stats()["n_device_ops"] == 7      # the Xor chain (its AssertZero is fused); the loops are descriptors


# ----------------------------------------------------------------------------------------------------------------- GPU
def use_jit(b, jit):
    """jit: wait for the groups' specialised kernel (group_jit.cpp) so that the evaluation that follows runs it;
    returns False when this box cannot compile it (no NVRTC: the interpreter kernel stays in charge)"""
    if not jit:
        return False
    state, _ = b.wait_group_jit()
    assert state in (2, 3, -1), state
    return state != -1


def check_against_oracle(r, w, exact_groups=None, jit=False):
    msgs = r.messages(w)
    expected = ev.evaluate(msgs)
    b, e = record(0, msgs)
    if exact_groups is None:
        assert 1 <= b.stats()["n_call_groups"] <= r.expect_groups
    else:
        assert b.stats()["n_call_groups"] == exact_groups
    b.finalize(True)
    jitted = use_jit(b, jit)
    assert e.get_violations() == expected
    assert b.stats()["group_jit_state"] == (3 if jitted else b.stats()["group_jit_state"])
    o = ev.Evaluator.from_messages(msgs, ev.PlaintextBackend())
    if expected == []:
        for wid, val in o.values.items():
            assert e.get(wid) == val, wid
    return b, e, o


@pytest.fixture(params=["interpreter", "specialised"])
def jit(request, monkeypatch):
    """both expansions of a call group: the interpreter kernel (k_bool_groups; ZKB_GROUP_JIT=0 keeps the background
    compilation from taking over mid-test) and the run-time specialised kernel (waited for before the evaluation)"""
    if request.param == "interpreter":
        monkeypatch.setenv("ZKB_GROUP_JIT", "0")
        return False
    return True


@pytest.mark.gpu
@pytest.mark.parametrize("seed", SEEDS)
def test_gpu_values_and_verdicts(seed, jit):
    r = build(seed)
    check_against_oracle(r, r.witness(), jit=jit)


@pytest.mark.gpu
def test_gpu_slot_tables(jit):
    r, _ = table_relation()
    for trial in range(4):
        check_against_oracle(r, r.witness(), exact_groups=2, jit=jit)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [100, 101, 102, 103, 104, 105])
def test_gpu_raw_witness_values(seed, jit):
    """witness bytes 0..4: And / Xor / Add / Mul inside a call see the raw integers' low bits (evaluator.rs:908-930), the
    Not / AssertZero outside see what the calls computed"""
    r = build(seed, raw=True)
    check_against_oracle(r, r.witness(), jit=jit)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [200, 201, 202, 203])
def test_gpu_failing_assertion_on_a_group_output(seed):
    r = build(seed, failing=True)
    for trial in range(4):
        w = r.witness()
        msgs = r.messages(w)
        b, e = record(0, msgs)
        assert e.get_violations() == ev.evaluate(msgs)


@pytest.mark.gpu
@pytest.mark.parametrize("n_batch", [1, 33, 64, 100, 128, 300, 1000])
def test_gpu_batches_bit_sliced(n_batch, jit):
    """one recording, n_batch witnesses (ragged last word / tile): verdict per witness and probed values vs the oracle"""
    r = build(7, failing=True)
    w0 = r.witness()
    b, e = record(0, r.messages(w0))
    b.finalize(True)
    use_jit(b, jit)
    rng = np.random.default_rng(n_batch)
    W = rng.integers(0, 2, size=(n_batch, r.n_wit, 1)).astype(np.uint8)
    v = b.evaluate(None, W, n_batch)
    for j in sorted(set([0, n_batch - 1, n_batch // 2, min(31, n_batch - 1), min(32, n_batch - 1)])):
        msgs = r.messages(W[j, :, 0])
        viol = ev.evaluate(msgs)
        assert bool(v[j]["ok"]) == (viol == []), j
        o = ev.Evaluator.from_messages(msgs, ev.PlaintextBackend())
        wires = sorted(o.values)
        got = b.read_values(j, [e.value_handle(x) for x in wires], 4)
        if viol == []:
            assert got == [o.values[x] for x in wires], j


@pytest.mark.gpu
def test_gpu_c5_grouped_equals_serial(jit):
    """the C5 relation both ways on the device: same verdicts and outputs for 64 witnesses"""
    lo, li, n_wit = 5, 6, 512
    rel, _ = wl.boolean_for_relation(lo, li, n_wit)
    msgs = [ir.Witness(rel.header, [b"\0"] * n_wit), rel]
    rng = np.random.default_rng(11)
    W = rng.integers(0, 2, size=(64, n_wit, 1)).astype(np.uint8)
    res = []
    for no_groups in (False, True):
        b, e = record(0, msgs, no_groups=no_groups)
        assert (b.stats()["n_call_groups"] == 0) == no_groups
        b.finalize(True)
        jitted = use_jit(b, jit and not no_groups)
        v = b.evaluate(None, W, 64)
        if jitted:
            assert b.stats()["group_jit_state"] == 3
        n_out = (1 << lo) * (2 << li)
        vals = [b.read_values(j, [e.value_handle(n_wit + k) for k in range(n_out)], 4) for j in (0, 31, 32, 63)]
        res.append(([bool(x["ok"]) for x in v], vals))
    assert res[0] == res[1]
    for idx, j in enumerate((0, 31, 32, 63)):
        outs = wl.boolean_for_expected_outputs(W[j, :, 0], lo, li).reshape(-1)
        assert res[0][1][idx] == [int(x) for x in outs]
