"""Structured synthetic relations (test / bench infrastructure; needs the FlatBuffers writer).

boolean_for_relation — config C5 of BASELINE.json: a Boolean-profile (field 2) relation whose
2^(a+b+3) leaf gates exist only after the nested For loops are unrolled on the host:

    Function "mix" (3 inputs -> 2 outputs, 8 gates: 4 Xor, 3 And, 1 Not)
    Witness bits  w[0 .. n_wit)                 (a For loop of Witness gates)
    For i in 0 ..= 2^a - 1   (anonymous body, iterators forwarded)
        For j in 0 ..= 2^b - 1   (IterExprCall "mix")
            outputs  (2j, 2j+1)            of the outer body's output block
            inputs   body wires  base + 2j,  base + 2j + 1,  base + n_wit/2 + i/4
    final chain: Xor-fold of eight outputs of the last block, AssertZero.

The inner iterations only read witness bits, so all 2^(a+b) calls are independent (depth = depth of
"mix"); with an all-zero witness every wire is 0 and the statement is TRUE.
"""
from __future__ import annotations

import numpy as np

from . import ir


def mix_function() -> ir.Function:
    # outputs 0,1 ; inputs 2,3,4 ; locals 5..
    return ir.Function("mix", 2, 3, 0, 0, [
        ("Xor", 5, 2, 3),      # g1 = a ^ b
        ("Xor", 0, 5, 4),      # out0 = g1 ^ c
        ("And", 6, 2, 3),      # g3 = a & b
        ("And", 7, 5, 4),      # g4 = g1 & c
        ("Xor", 8, 6, 7),      # g5 = carry
        ("Not", 9, 8),         # g6 = !g5
        ("And", 10, 9, 0),     # g7 = g6 & out0
        ("Xor", 1, 10, 8),     # out1 = g7 ^ g5
    ])


def mix_numpy(a, b, c):
    """the same function on numpy bit arrays (independent check for the full-size run)"""
    g1 = a ^ b
    out0 = g1 ^ c
    g5 = (a & b) ^ (g1 & c)
    out1 = ((1 - g5) & out0) ^ g5
    return out0, out1


def boolean_for_relation(log2_outer: int, log2_inner: int, n_wit: int = 4096):
    n_outer, n_inner = 1 << log2_outer, 1 << log2_inner
    assert 2 * n_inner <= n_wit // 2 and n_outer // 4 <= n_wit // 2
    block = 2 * n_inner
    I, C = (lambda n: ("Name", n)), (lambda v: ("Const", v))
    add = lambda l, r: ("Add", l, r)
    mul = lambda l, r: ("Mul", l, r)
    h = ir.Header(bytes([2]))
    gates = []
    gates.append(("For", "k", 0, n_wit - 1, [ir.WireRange(0, n_wit - 1)],
                  ("IterExprAnonCall", [("Single", I("k"))], [], 0, 1, [("Witness", 0)])))
    # outer body: outputs = its block (local 0..block-1), inputs = ALL witness wires (local block..block+n_wit-1)
    inner = ("For", "j", 0, n_inner - 1, [ir.WireRange(0, block - 1)],
             ("IterExprCall", "mix",
              [("Range", mul(I("j"), C(2)), add(mul(I("j"), C(2)), C(1)))],
              [("Single", add(C(block), mul(I("j"), C(2)))),
               ("Single", add(C(block + 1), mul(I("j"), C(2)))),
               ("Single", add(C(block + n_wit // 2), ("DivConst", I("i"), 4)))]))
    base = n_wit
    gates.append(("For", "i", 0, n_outer - 1, [ir.WireRange(base, base + n_outer * block - 1)],
                  ("IterExprAnonCall",
                   [("Range", add(C(base), mul(I("i"), C(block))), add(C(base + block - 1), mul(I("i"), C(block))))],
                   [("Range", C(0), C(n_wit - 1))], 0, 0, [inner])))
    last = base + (n_outer - 1) * block
    t = base + n_outer * block
    gates.append(("Xor", t, last, last + 1))
    for k in range(2, 8):
        gates.append(("Xor", t + k - 1, t + k - 2, last + k))
    gates.append(("AssertZero", t + 6))
    rel = ir.Relation(h, ir.BOOL, ir.FOR | ir.FUNCTION, [mix_function()], gates)
    n_leaf = n_outer * n_inner * 8
    return rel, n_leaf


def boolean_for_expected_outputs(w: np.ndarray, log2_outer: int, log2_inner: int):
    """all outputs of the unrolled loops for witness bits w (uint8[n_wit]) -> uint8 [n_outer, 2*n_inner]"""
    n_outer, n_inner = 1 << log2_outer, 1 << log2_inner
    n_wit = len(w)
    j = np.arange(n_inner)
    a = w[2 * j][None, :].repeat(n_outer, 0)
    b = w[2 * j + 1][None, :].repeat(n_outer, 0)
    c = w[n_wit // 2 + np.arange(n_outer) // 4][:, None].repeat(n_inner, 1)
    o0, o1 = mix_numpy(a, b, c)
    out = np.empty((n_outer, 2 * n_inner), dtype=np.uint8)
    out[:, 0::2] = o0
    out[:, 1::2] = o1
    return out
