"""Synthetic workloads of BASELINE.json's configs (numpy, deterministic).

These are the INPUTS of the benchmark and of the parity tests — flat relations in the
`zkb_gate` array form (the simple arms of the reference's `Gate`,
rust/src/structs/gates.rs:18-45) plus witness vectors.  Nothing here evaluates a gate.

random_circuit  "random Add/Mul/AssertZero circuit" (configs C2, C3):
    wire 0            Constant(p-1)                                   (not a counted gate)
    wires 1..n_in     Witness                                         (not counted)
    then slots drawn 45 % Add, 45 % Mul, 10 % assertion, operands uniform over all earlier
    wires (or over the last `window` wires).  An assertion slot is either
      identity:  n = Mul(t, w0); s = Add(t, n); AssertZero(s)     -- (p-1)*t + t == 0 for any t:
                 exercises modular mul/add, holds for every witness            (3 counted gates)
      tie:       s = Add(x_2k, x_2k+1); AssertZero(s)             -- holds iff the witness was
                 built with x_2k+1 = p - x_2k; corrupting x_2k+1 makes exactly this assertion
                 (and possibly later ones) fail                                (2 counted gates)
    so every witness vector is satisfiable WITHOUT evaluating the circuit, and a corrupted one
    has a first failing assertion known by construction.  Counted gates (Add + Mul +
    AssertZero) are exactly `n_gates`.
"""
from __future__ import annotations

import numpy as np

GATE_DTYPE = np.dtype([("op", "u1"), ("pad", "u1", (3,)), ("out", "<u4"), ("a", "<u4"), ("b", "<u4")])
(G_CONSTANT, G_ASSERT_ZERO, G_COPY, G_ADD, G_MUL, G_ADD_CONSTANT, G_MUL_CONSTANT, G_AND, G_XOR, G_NOT, G_INSTANCE,
 G_WITNESS, G_FREE) = range(1, 14)

GOLDILOCKS = (1 << 64) - (1 << 32) + 1
BLS12_381_FR = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
BN254_FR = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


def elem_bytes(p: int) -> int:
    bits = p.bit_length()
    return 4 if bits <= 32 else 8 if bits <= 64 else 16 if bits <= 128 else 32


def le_bytes(v: int, n: int) -> np.ndarray:
    return np.frombuffer(int(v).to_bytes(n, "little"), dtype=np.uint8)


class FlatCircuit:
    def __init__(self):
        self.p = 0
        self.gates = None          # GATE_DTYPE
        self.const_pool = None     # uint8 [n_consts, stride]
        self.n_inputs = 0          # Witness gates
        self.n_gates = 0           # counted gates: Add + Mul + AssertZero
        self.n_wires = 0
        self.n_ties = 0
        self.tie_assert_seq = None  # assert seq of tie k
        self.n_asserts = 0
        self.hist = {}


def random_circuit(n_gates: int, n_inputs: int, p: int, seed: int, n_ties: int = 32, window: int = 0,
                   assert_frac: float = 0.10) -> FlatCircuit:
    rng = np.random.default_rng(seed)
    n_ties = min(n_ties, n_inputs // 2)
    # --- draw slot kinds until the counted gates reach n_gates -----------------------------
    est = int(n_gates / (1.0 + 2.0 * assert_frac)) + 16
    kinds = rng.choice(3, size=est + 64, p=[(1 - assert_frac) / 2, (1 - assert_frac) / 2, assert_frac]).astype(np.int8)
    # kinds: 0 add, 1 mul, 2 identity-assert ; a few assertion slots become ties (kind 3)
    a_slots = np.flatnonzero(kinds == 2)
    if len(a_slots) < n_ties:
        n_ties = len(a_slots)
    tie_slots = np.sort(rng.choice(a_slots, size=n_ties, replace=False)) if n_ties else np.zeros(0, np.int64)
    kinds[tie_slots] = 3
    cost = np.array([1, 1, 3, 2], dtype=np.int64)[kinds]
    cum = np.cumsum(cost)
    n_slots = int(np.searchsorted(cum, n_gates, side="right"))
    kinds = kinds[:n_slots]
    cost = cost[:n_slots]
    missing = n_gates - (int(cum[n_slots - 1]) if n_slots else 0)
    if missing:  # pad with Add slots to hit n_gates exactly
        kinds = np.concatenate([kinds, np.zeros(missing, np.int8)])
        cost = np.concatenate([cost, np.ones(missing, np.int64)])
        n_slots += missing
    tie_slots = tie_slots[tie_slots < n_slots]
    n_ties = len(tie_slots)
    # --- wire ids ------------------------------------------------------------------------------
    wires_per = np.array([1, 1, 2, 1], dtype=np.int64)[kinds]
    first_wire = 1 + n_inputs + np.concatenate([[0], np.cumsum(wires_per)[:-1]])
    n_wires = int(1 + n_inputs + wires_per.sum())
    if n_wires >= (1 << 32) - 16:
        raise ValueError("circuit too large for 32-bit wire ids")
    emit_per = np.array([1, 1, 3, 2], dtype=np.int64)[kinds]  # gates emitted (all counted)
    first_gate = 1 + n_inputs + np.concatenate([[0], np.cumsum(emit_per)[:-1]])
    n_emit = int(1 + n_inputs + emit_per.sum())

    def pick(limit):  # operand uniform over wires [lo, limit)
        u = rng.random(len(limit))
        if window:
            lo = np.maximum(limit - window, 0)
            return (lo + (u * (limit - lo)).astype(np.int64)).astype(np.uint32)
        return (u * limit).astype(np.int64).astype(np.uint32)

    opa = pick(first_wire)
    opb = pick(first_wire)
    g = np.zeros(n_emit, dtype=GATE_DTYPE)
    g["op"][0] = G_CONSTANT
    g["out"][0] = 0
    g["b"][0] = 0
    g["op"][1:1 + n_inputs] = G_WITNESS
    g["out"][1:1 + n_inputs] = np.arange(1, 1 + n_inputs, dtype=np.uint32)
    for k, opc in ((0, G_ADD), (1, G_MUL)):
        s = np.flatnonzero(kinds == k)
        gi = first_gate[s]
        g["op"][gi] = opc
        g["out"][gi] = first_wire[s]
        g["a"][gi] = opa[s]
        g["b"][gi] = opb[s]
    s = np.flatnonzero(kinds == 2)  # identity: n = Mul(t, w0); s = Add(t, n); AssertZero(s)
    gi = first_gate[s]
    g["op"][gi] = G_MUL
    g["out"][gi] = first_wire[s]
    g["a"][gi] = opa[s]
    g["b"][gi] = 0
    g["op"][gi + 1] = G_ADD
    g["out"][gi + 1] = first_wire[s] + 1
    g["a"][gi + 1] = opa[s]
    g["b"][gi + 1] = first_wire[s]
    g["op"][gi + 2] = G_ASSERT_ZERO
    g["a"][gi + 2] = first_wire[s] + 1
    s = tie_slots  # tie k: s = Add(x_2k, x_2k+1); AssertZero(s)   (inputs are wires 1..n_in)
    gi = first_gate[s]
    kk = np.arange(n_ties, dtype=np.uint32)
    g["op"][gi] = G_ADD
    g["out"][gi] = first_wire[s]
    g["a"][gi] = 1 + 2 * kk
    g["b"][gi] = 2 + 2 * kk
    g["op"][gi + 1] = G_ASSERT_ZERO
    g["a"][gi + 1] = first_wire[s]

    c = FlatCircuit()
    c.p = p
    c.gates = g
    eb = elem_bytes(p)
    c.const_pool = le_bytes(p - 1, eb).reshape(1, eb).copy()
    c.n_inputs = n_inputs
    c.n_gates = int(emit_per.sum())
    assert c.n_gates == n_gates, (c.n_gates, n_gates)
    c.n_wires = n_wires
    c.n_ties = n_ties
    # assert sequence numbers (program order)
    is_assert = g["op"] == G_ASSERT_ZERO
    seq_of_gate = np.cumsum(is_assert) - 1
    c.tie_assert_seq = seq_of_gate[first_gate[tie_slots] + 1].astype(np.int64) if n_ties else np.zeros(0, np.int64)
    c.n_asserts = int(is_assert.sum())
    c.hist = {"add": int((g["op"] == G_ADD).sum()), "mul": int((g["op"] == G_MUL).sum()), "assert_zero": c.n_asserts}
    return c


def random_field_elements(rng, shape, p: int) -> np.ndarray:
    """uniform in [0, p) by rejection; returns uint8 [*shape, elem_bytes(p)] little-endian"""
    eb = elem_bytes(p)
    n = int(np.prod(shape))
    bits = p.bit_length()
    nl = eb // 4
    plimbs = np.array([(p >> (32 * i)) & 0xFFFFFFFF for i in range(nl)], dtype=np.uint32)
    out = np.zeros((n, nl), dtype=np.uint32)
    todo = np.arange(n)
    top_bits = bits - 32 * (nl - 1)
    top_mask = np.uint32((1 << top_bits) - 1) if top_bits < 32 else np.uint32(0xFFFFFFFF)
    while len(todo):
        cand = rng.integers(0, 1 << 32, size=(len(todo), nl), dtype=np.uint64).astype(np.uint32)
        cand[:, nl - 1] &= top_mask
        # cand < p, lexicographic from the top limb
        lt = np.zeros(len(todo), dtype=bool)
        eq = np.ones(len(todo), dtype=bool)
        for i in range(nl - 1, -1, -1):
            lt |= eq & (cand[:, i] < plimbs[i])
            eq &= cand[:, i] == plimbs[i]
        out[todo[lt]] = cand[lt]
        todo = todo[~lt]
    return out.view(np.uint8).reshape(*shape, eb)


def make_witnesses(c: FlatCircuit, n_batch: int, seed: int, corrupt=None) -> np.ndarray:
    """uint8 [n_batch, n_inputs, elem_bytes]; tie inputs satisfy x_2k+1 = p - x_2k.
    corrupt: {witness index: tie index} -> that tie is broken (x_2k+1 += 1 mod p)."""
    rng = np.random.default_rng(seed ^ 0x5EED)
    eb = elem_bytes(c.p)
    w = random_field_elements(rng, (n_batch, c.n_inputs), c.p)
    for k in range(c.n_ties):
        xs = w[:, 2 * k, :]
        for j in range(n_batch):
            x = int.from_bytes(xs[j].tobytes(), "little")
            y = (c.p - x) % c.p
            if corrupt and corrupt.get(j) == k:
                y = (y + 1) % c.p
            w[j, 2 * k + 1, :] = le_bytes(y, eb)
    return w


def expected_first_fail(c: FlatCircuit, n_batch: int, corrupt=None) -> np.ndarray:
    """first failing assert seq per witness (-1: TRUE), by construction"""
    out = np.full(n_batch, -1, dtype=np.int64)
    for j, k in (corrupt or {}).items():
        out[j] = int(c.tie_assert_seq[k])
    return out


def algorithmic_bytes_per_witness(c: FlatCircuit) -> int:
    """Add/Mul = 3E, AssertZero = E"""
    E = elem_bytes(c.p)
    return 3 * E * (c.hist["add"] + c.hist["mul"]) + E * c.hist["assert_zero"]


# ---------------------------------------------------------------------------------------------
# C4: "multiplication-gate" shaped R1CS
# ---------------------------------------------------------------------------------------------
class R1cs:
    def __init__(self):
        self.p = 0
        self.n_rows = 0
        self.n_vars = 0
        self.n_free = 0
        self.A = self.B = self.C = None   # (row_ptr uint64, col uint32, coef_idx uint32)
        self.coef_table = None            # uint8 [n_coefs, elem_bytes]
        self.coefs = None                 # python ints
        self.nnz = 0


def random_r1cs(n_rows: int, n_free: int, p: int, seed: int) -> R1cs:
    """Row r: (A_r . z)(B_r . z) = z[s_r] with s_r = n_free + 1 + r a fresh slack variable.
    A_r, B_r: 1 + Poisson(2) terms (capped at 8) over uniformly chosen ids < s_r; coefficients from a
    seeded 256-entry table plus {1, p-1}.  Variable 0 is the constant one."""
    rng = np.random.default_rng(seed)
    eb = elem_bytes(p)
    coefs = [1, p - 1] + [int.from_bytes(rng.bytes(eb), "little") % p for _ in range(256)]
    r = R1cs()
    r.p = p
    r.n_rows = n_rows
    r.n_free = n_free
    r.n_vars = 1 + n_free + n_rows
    r.coefs = coefs
    r.coef_table = np.stack([le_bytes(c, eb) for c in coefs])
    mats = []
    for _ in range(2):
        cnt = np.minimum(1 + rng.poisson(2.0, size=n_rows), 8).astype(np.int64)
        row_ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
        nnz = int(row_ptr[-1])
        rows = np.repeat(np.arange(n_rows, dtype=np.int64), cnt)
        limit = n_free + 1 + rows                       # ids < s_r
        col = (rng.random(nnz) * limit).astype(np.int64).astype(np.uint32)
        ci = rng.integers(0, len(coefs), size=nnz).astype(np.uint32)
        # half of the coefficients are the literal 1, as in real R1CS
        ci[rng.random(nnz) < 0.5] = 0
        mats.append((row_ptr, col, ci))
    r.A, r.B = mats
    r.C = (np.arange(n_rows + 1, dtype=np.uint64), (n_free + 1 + np.arange(n_rows)).astype(np.uint32),
           np.zeros(n_rows, dtype=np.uint32))
    r.nnz = int(r.A[0][-1] + r.B[0][-1] + n_rows)
    return r


def r1cs_assignment(r: R1cs, seed: int):
    """A satisfying assignment z (python ints): free variables random, slack z[s_r] := (A_r.z)(B_r.z) mod p."""
    rng = np.random.default_rng(seed ^ 0xA551)
    p = r.p
    free = random_field_elements(rng, (r.n_free,), p)
    z = [1] + [int.from_bytes(free[i].tobytes(), "little") for i in range(r.n_free)] + [0] * r.n_rows
    coefs = r.coefs
    (rpa, ca, ia), (rpb, cb, ib) = r.A, r.B
    rpa, ca, ia, rpb, cb, ib = (x.tolist() for x in (rpa, ca, ia, rpb, cb, ib))
    base = r.n_free + 1
    for row in range(r.n_rows):
        a = 0
        for e in range(rpa[row], rpa[row + 1]):
            a += coefs[ia[e]] * z[ca[e]]
        b = 0
        for e in range(rpb[row], rpb[row + 1]):
            b += coefs[ib[e]] * z[cb[e]]
        z[base + row] = (a % p) * (b % p) % p
    return z


def assignment_bytes(z, p: int) -> np.ndarray:
    eb = elem_bytes(p)
    buf = b"".join(int(v).to_bytes(eb, "little") for v in z)
    return np.frombuffer(buf, dtype=np.uint8).reshape(len(z), eb).copy()


def r1cs_first_row_reading(r: R1cs, var: int) -> int:
    """first row whose A, B or C mentions variable `var`"""
    best = r.n_rows
    for rp, col, _ in (r.A, r.B, r.C):
        hits = np.flatnonzero(col == var)
        if len(hits):
            row = int(np.searchsorted(rp, hits[0], side="right") - 1)
            best = min(best, row)
    return best


def r1cs_algorithmic_bytes(r: R1cs) -> int:
    """per nnz 4 B col + 4 B coef index + E gathered z; per row 3 x 4 B row_ptr"""
    return (8 + elem_bytes(r.p)) * r.nnz + 12 * r.n_rows


def r1cs_gate_equivalent(r: R1cs) -> int:
    """gates `zkif-to-ir` emits for this system: 2 per term with id != 0 (Constant + Mul), 1 per id-0 term,
    (terms - 1) Adds per LC, 4 per row (from_r1cs.rs:71-125)"""
    total = 0
    for rp, col, _ in (r.A, r.B, r.C):
        total += int(2 * (col != 0).sum() + (col == 0).sum())
        cnt = np.diff(rp.astype(np.int64))
        total += int(np.maximum(cnt - 1, 0).sum() + (cnt == 0).sum())
    return total + 4 * r.n_rows
