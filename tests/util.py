"""Shared helpers of the test-suite."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import zkb_loader  # noqa: E402


def zkb():
    return zkb_loader.load()


def circuits():
    zkb_loader.load()
    return importlib.import_module("zkir_b200.circuits")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


FIELDS = {
    "p101": 101,
    "m31": (1 << 31) - 1,
    "goldilocks": (1 << 64) - (1 << 32) + 1,
    "p61": (1 << 61) - 1,
    "p64full": (1 << 64) - 59,                           # largest 64-bit prime: one-limb (u64) Montgomery path, top bit set
    "kat124": 16249742125730185677094195492597105093,   # modulus of evaluator.rs:956
    "bn254": 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001,
    "bls381": 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    "p256full": (1 << 256) - 189,                        # top bit set: exercises the carry paths
}


def random_flat_program(p, n_ops, n_inst, n_wit, seed, bool_ops=False):
    """Random flat relation using EVERY simple gate kind, wire ids re-used after Free.
    Returns (gates GATE_DTYPE, const_pool uint8[n,stride], n_wires_upper_bound)."""
    c = circuits()
    rng = np.random.default_rng(seed)
    eb = c.elem_bytes(p)
    consts = [0, 1, p - 1] + [int(rng.integers(0, 1 << 62)) % p for _ in range(5)]
    if p > (1 << 64):
        consts += [int.from_bytes(rng.bytes(32), "little") % p for _ in range(4)]
    pool = np.stack([c.le_bytes(v, eb) for v in consts])
    gates = []
    live = []
    free_ids = []
    next_id = 0

    def new_id():
        nonlocal next_id
        if free_ids and rng.random() < 0.7:
            return free_ids.pop(int(rng.integers(0, len(free_ids))))
        next_id += 1
        return next_id - 1

    def emit(op, out=0, a=0, b=0):
        gates.append((op, out, a, b))

    for _ in range(n_inst):
        w = new_id(); emit(c.G_INSTANCE, w); live.append(w)
    for _ in range(n_wit):
        w = new_id(); emit(c.G_WITNESS, w); live.append(w)
    w = new_id(); emit(c.G_CONSTANT, w, 0, 2); live.append(w)
    arith = [c.G_ADD, c.G_MUL, c.G_ADD_CONSTANT, c.G_MUL_CONSTANT, c.G_COPY, c.G_CONSTANT]
    if bool_ops:
        arith += [c.G_AND, c.G_XOR, c.G_NOT]
    for _ in range(n_ops):
        r = rng.random()
        if r < 0.06 and len(live) > 8:
            # free a short run of consecutive live ids
            s = sorted(live)
            i = int(rng.integers(0, len(s)))
            j = i
            while j + 1 < len(s) and s[j + 1] == s[j] + 1 and j - i < 3:
                j += 1
            emit(c.G_FREE, 0, s[i], s[j])
            for k in s[i:j + 1]:
                live.remove(k); free_ids.append(k)
            continue
        op = arith[int(rng.integers(0, len(arith)))]
        a = live[int(rng.integers(0, len(live)))]
        b = live[int(rng.integers(0, len(live)))]
        if op == c.G_CONSTANT:
            w = new_id(); emit(op, w, 0, int(rng.integers(0, len(consts))))
        elif op in (c.G_ADD_CONSTANT, c.G_MUL_CONSTANT):
            w = new_id(); emit(op, w, a, int(rng.integers(0, len(consts))))
        elif op in (c.G_COPY, c.G_NOT):
            w = new_id(); emit(op, w, a)
        else:
            w = new_id(); emit(op, w, a, b)
        live.append(w)
        if rng.random() < 0.08:
            # an assertion that holds: t + (p-1)*t
            n = new_id(); emit(c.G_MUL_CONSTANT, n, w, 2); live.append(n)
            s_ = new_id(); emit(c.G_ADD, s_, w, n); live.append(s_)
            emit(c.G_ASSERT_ZERO, 0, s_)
        elif rng.random() < 0.01:
            emit(c.G_ASSERT_ZERO, 0, w)   # almost surely fails: data-dependent first failure
    g = np.zeros(len(gates), dtype=c.GATE_DTYPE)
    for i, (op, out, a, b) in enumerate(gates):
        g[i]["op"], g[i]["out"], g[i]["a"], g[i]["b"] = op, out, a, b
    return g, pool, next_id
