"""The reference's own example statements, restated as data (oracle side).

  example_*            rust/src/producers/examples.rs:11-212
  boolean_example_*    rust/src/producers/boolean_examples.rs:5-239
  GateBuilder & co     rust/src/producers/builder.rs:136-724 (+ build_gates.rs,
                       structs/gates.rs:742-854 replace_output_wires) — only as far
                       as the four builder tests need it
  builder_*            the circuits of rust/src/producers/builder.rs:726-1175
  r1cs_to_gates        rust/src/producers/from_r1cs.rs:27-141
  zkif_example_*       the zkinterface 1.3.2 example used by from_r1cs.rs:178-221
                       (crate source absent; pinned by the gate counts and wire
                       values asserted in from_r1cs.rs:209-286)
"""
from __future__ import annotations

from typing import List, Optional

from . import ir
from .ir import Wire, WireRange, wirelist, wirelist_rep, literal32

EXAMPLE_MODULUS = 101


# ------------------------------------------------------------------ examples.rs
def example_header(modulus: int = EXAMPLE_MODULUS) -> ir.Header:
    # examples.rs:11-13, 31-36: literal32(101) for the default, to_bytes_le otherwise
    fc = literal32(modulus) if modulus < (1 << 32) else ir.le_bytes(modulus)
    return ir.Header(field_characteristic=fc)


def example_instance(header=None) -> ir.Instance:
    header = header or example_header()
    return ir.Instance(header, [literal32(25), literal32(0), literal32(1)])


def example_witness(header=None) -> ir.Witness:
    header = header or example_header()
    m = int.from_bytes(header.field_characteristic, "little")
    fib22 = 17711 % m
    # BigUint::to_bytes_le gives [0] for zero
    return ir.Witness(header, [literal32(3), literal32(4), literal32(0), ir.le_bytes(fib22)])


def example_witness_incorrect(header=None) -> ir.Witness:
    header = header or example_header()
    return ir.Witness(header, [literal32(3), literal32(5), literal32(1), literal32(40)])


def encode_negative_one(header: ir.Header) -> bytes:
    b = bytearray(header.field_characteristic)
    assert len(b) > 0 and b[0] > 0
    b[0] -= 1
    return bytes(b)


def example_relation(header=None) -> ir.Relation:
    header = header or example_header()
    mul = "com.example::mul"
    I = lambda n: ("Name", n)
    C = lambda v: ("Const", v)
    return ir.Relation(
        header=header,
        gate_mask=ir.ADD | ir.MUL | ir.MULC,
        feat_mask=ir.FUNCTION | ir.SWITCH | ir.FOR,
        functions=[ir.Function(mul, 1, 2, 0, 0, [("Mul", 0, 1, 2)])],
        gates=[
            ("Witness", 1),
            ("Switch", 1, wirelist(0, 2, 4, 5, 6, 9, 10, 11), [bytes([3]), bytes([5])], [
                ("AbstractAnonCall", wirelist(1), 3, 3, [
                    ("Instance", 0),
                    ("Witness", 1),
                    ("Call", mul, wirelist(2), wirelist_rep(8, 2)),
                    ("Call", mul, wirelist(3), wirelist_rep(1, 2)),
                    ("Add", 4, 2, 3),
                    ("Witness", 9),
                    ("AssertZero", 9),
                    ("Instance", 6),
                    ("AssertZero", 6),
                    ("Instance", 7),
                    ("Witness", 5),
                ]),
                ("AbstractAnonCall", wirelist(1), 3, 2, [
                    ("Instance", 0),
                    ("Call", mul, wirelist(1), wirelist(8, 0)),
                    ("Witness", 2),
                    ("Mul", 3, 1, 2),
                    ("Add", 4, 2, 3),
                    ("Instance", 5),
                    ("Instance", 6),
                    ("Witness", 7),
                    ("AssertZero", 5),
                    ("AssertZero", 0),
                ]),
            ]),
            ("Constant", 3, encode_negative_one(header)),
            ("Call", mul, wirelist(7), wirelist(3, 0)),
            ("Add", 8, 6, 7),
            ("Free", 0, 7),
            ("AssertZero", 8),
            ("For", "i", 0, 20, [WireRange(12, 32)],
             ("IterExprAnonCall",
              [("Single", ("Add", I("i"), C(12)))],
              [("Single", ("Add", I("i"), C(10))), ("Single", ("Add", I("i"), C(11)))],
              0, 0, [("Add", 0, 1, 2)])),
            ("MulConstant", 33, 32, encode_negative_one(header)),
            ("Add", 34, 9, 33),
            ("AssertZero", 34),
            ("For", "i", 35, 50, [WireRange(35, 50)],
             ("IterExprCall", mul,
              [("Single", I("i"))],
              [("Single", ("Sub", I("i"), C(1))), ("Single", ("Sub", I("i"), C(2)))])),
            ("Free", 8, 50),
        ])


# ---------------------------------------------------------- boolean_examples.rs
def boolean_header() -> ir.Header:
    return ir.Header(field_characteristic=bytes([2]))


def boolean_example_instance() -> ir.Instance:
    return ir.Instance(boolean_header(), [bytes([v]) for v in (0, 0, 0, 0, 0, 1, 0, 1)])


def boolean_example_witness() -> ir.Witness:
    return ir.Witness(boolean_header(), [bytes([v]) for v in (1, 0, 1, 0, 0)])


def boolean_example_witness_incorrect() -> ir.Witness:
    return ir.Witness(boolean_header(), [bytes([v]) for v in (1, 1, 1, 0, 0)])


def boolean_example_relation() -> ir.Relation:
    I = lambda n: ("Name", n)
    C = lambda v: ("Const", v)
    aff = lambda k: ("Add", ("Mul", I("i"), C(3)), C(k))   # 3i + k
    return ir.Relation(
        header=boolean_header(),
        gate_mask=ir.AND | ir.XOR | ir.NOT,
        feat_mask=ir.FUNCTION | ir.SWITCH | ir.FOR,
        functions=[ir.Function("two_bit_adder", 3, 4, 0, 0, [
            ("Xor", 2, 4, 6), ("And", 7, 4, 6), ("Xor", 8, 3, 5), ("Xor", 1, 7, 8), ("And", 9, 3, 5),
            ("Not", 10, 9), ("And", 11, 8, 7), ("Not", 12, 11), ("And", 13, 10, 12), ("Not", 0, 13),
            ("Free", 7, 13)])],
        gates=[
            ("For", "i", 0, 2, [WireRange(0, 2)],
             ("IterExprAnonCall", [("Single", I("i"))], [], 0, 1, [("Witness", 0)])),
            ("For", "i", 3, 8, [WireRange(3, 8)],
             ("IterExprAnonCall", [("Single", I("i"))], [], 1, 0, [("Instance", 0)])),
            ("For", "i", 0, 3, [WireRange(9, 20)],
             ("IterExprCall", "two_bit_adder",
              [("Range", aff(9), aff(11))],
              [("Single", aff(4)), ("Single", aff(5)), ("Single", aff(7)), ("Single", aff(8))])),
            ("Free", 3, 17),
            ("Xor", 21, 18, 0), ("Xor", 22, 19, 1), ("Xor", 23, 20, 2),
            ("AssertZero", 21), ("AssertZero", 22), ("AssertZero", 23),
            ("Free", 0, 2), ("Free", 18, 23),
            ("Witness", 24), ("Witness", 25),
            ("Switch", 24, wirelist(26), [bytes([1]), bytes([0])], [
                ("AbstractAnonCall", [], 2, 0, [("Instance", 1), ("Instance", 2), ("Xor", 0, 1, 2)]),
                ("AbstractAnonCall", [], 2, 0, [("Instance", 1), ("Instance", 2), ("And", 0, 1, 2)]),
            ]),
            ("Xor", 27, 26, 25),
            ("AssertZero", 27),
            ("Free", 24, 27),
        ])


# ------------------------------------------------------------------ builder.rs
NO_OUTPUT = (1 << 64) - 1
_HAS_OUTPUT = {"Constant", "Copy", "Add", "Mul", "AddConstant", "MulConstant", "And", "Xor", "Not",
               "Instance", "Witness"}


def _with_output(bg, out):
    """build_gates.rs:32-54.  bg = (kind, *args-without-output)."""
    k = bg[0]
    if k in ("AssertZero", "Free"):
        return bg
    if k in ("Instance", "Witness"):
        return (k, out)
    return (k, out) + tuple(bg[1:])


def _multiple_alloc(free_id, n):
    """builder.rs:213-224 -> (wirelist, new free_id)"""
    if n == 0:
        return [], free_id
    if n == 1:
        return [Wire(free_id)], free_id + 1
    return [WireRange(free_id, free_id + n - 1)], free_id + n


def _replace_in_wirelist(wl, old, new):
    ids = ir.expand_wirelist(wl)
    if old in ids:
        return [Wire(new if i == old else i) for i in ids]
    return wl


def replace_output_wires(gates: list, output_wires: List[int]):
    """structs/gates.rs:742-854"""
    if any(g[0] == "For" for g in gates):
        for i, w in enumerate(output_wires):
            gates.append(("Copy", i, w))
        return
    for i, old in enumerate(output_wires):
        new = i
        r = lambda w: new if w == old else w
        for n, g in enumerate(gates):
            k = g[0]
            if k == "Constant":
                gates[n] = (k, r(g[1]), g[2])
            elif k in ("Copy", "Not"):
                gates[n] = (k, r(g[1]), r(g[2]))
            elif k in ("Add", "Mul", "And", "Xor"):
                gates[n] = (k, r(g[1]), r(g[2]), r(g[3]))
            elif k in ("AddConstant", "MulConstant"):
                gates[n] = (k, r(g[1]), r(g[2]), g[3])
            elif k in ("Instance", "Witness", "AssertZero"):
                gates[n] = (k, r(g[1]))
            elif k == "Free":
                last = g[1] if g[2] is None else g[2]
                if g[1] <= old <= last:
                    raise ValueError("It is forbidden to free an output wire !")
            elif k == "AnonCall":
                gates[n] = (k, _replace_in_wirelist(g[1], old, new), _replace_in_wirelist(g[2], old, new)) + g[3:]
            elif k == "Call":
                gates[n] = (k, g[1], _replace_in_wirelist(g[2], old, new), _replace_in_wirelist(g[3], old, new))
            elif k == "Switch":
                brs = []
                for b in g[4]:
                    if b[0] == "AbstractAnonCall":
                        brs.append((b[0], _replace_in_wirelist(b[1], old, new)) + b[2:])
                    else:
                        brs.append((b[0], b[1], _replace_in_wirelist(b[2], old, new)))
                gates[n] = (k, r(g[1]), _replace_in_wirelist(g[2], old, new), g[3], brs)


class FunctionBuilder:
    """builder.rs:410-518"""

    def __init__(self, known, name, output_count, input_count):
        self.known = known
        self.name = name
        self.output_count = output_count
        self.input_count = input_count
        self.gates: list = []
        self.instance_count = 0
        self.witness_count = 0
        self.free_id = output_count + input_count

    def input_wire_ids(self):
        return list(range(self.output_count, self.output_count + self.input_count))

    def create_gate(self, *bg):
        k = bg[0]
        out = NO_OUTPUT
        if k in _HAS_OUTPUT:
            out = self.free_id
            self.free_id += 1
        if k == "Instance":
            self.instance_count += 1
        if k == "Witness":
            self.witness_count += 1
        self.gates.append(_with_output(bg, out))
        return out

    def create_complex_gate(self, cg):
        if cg[0] == "Call":
            p = self.known[cg[1]]
            assert ir.wirelist_len(cg[2]) == p["input_count"]
            oc, ic, wc = p["output_count"], p["instance_count"], p["witness_count"]
        else:
            oc, ic, wc = cg[4]["output_count"], cg[4]["instance_count"], cg[4]["witness_count"]
        outs, self.free_id = _multiple_alloc(self.free_id, oc)
        self.witness_count += wc
        self.instance_count += ic
        self.gates.append(_complex_with_output(cg, outs))
        return outs

    def finish(self, output_wires):
        assert len(output_wires) == self.output_count
        replace_output_wires(self.gates, output_wires)
        return ir.Function(self.name, self.output_count, self.input_count, self.instance_count,
                           self.witness_count, list(self.gates))


def _complex_with_output(cg, outs):
    """build_gates.rs:77-86"""
    if cg[0] == "Call":
        return ("Call", cg[1], outs, cg[2])
    return ("Switch", cg[1], outs, cg[2], cg[3])


class SwitchBuilder:
    """builder.rs:589-673"""

    def __init__(self, known, output_count):
        self.known = known
        self.output_count = output_count
        self.cases = []
        self.branches = []
        self.instance_count = 0
        self.witness_count = 0

    def push_branch_from(self, name, inputs, case: bytes):
        p = self.known[name]
        assert ir.wirelist_len(inputs) == p["input_count"]
        assert self.output_count == p["output_count"]
        assert case not in self.cases
        self.instance_count = max(self.instance_count, p["instance_count"])
        self.witness_count = max(self.witness_count, p["witness_count"])
        self.cases.append(case)
        self.branches.append(("AbstractGateCall", name, inputs))

    def finish(self, condition):
        assert self.branches
        return ("Switch", condition, self.cases, self.branches,
                dict(output_count=self.output_count, instance_count=self.instance_count,
                     witness_count=self.witness_count))


class GateBuilder:
    """builder.rs:136-408 with a MemorySink: produces [Instance?, Witness?, Relation]."""

    def __init__(self, header: ir.Header, gateset: int, features: int):
        self.header = header
        self.gateset = gateset
        self.features = features
        self.known = {}
        self.free_id = 0
        self.instance_vals: List[bytes] = []
        self.witness_vals: List[bytes] = []
        self.functions: List[ir.Function] = []
        self.gates: list = []

    def create_gate(self, *bg):
        k = bg[0]
        out = NO_OUTPUT
        if k in _HAS_OUTPUT:
            out = self.free_id
            self.free_id += 1
        if k == "Instance" and len(bg) > 1 and bg[1] is not None:
            self.instance_vals.append(bg[1])
        if k == "Witness" and len(bg) > 1 and bg[1] is not None:
            self.witness_vals.append(bg[1])
        self.gates.append(_with_output(bg, out))
        return out

    def create_complex_gate(self, cg, instances, witnesses):
        if cg[0] == "Call":
            p = self.known[cg[1]]
            assert ir.wirelist_len(cg[2]) == p["input_count"]
            assert len(instances) == p["instance_count"] and len(witnesses) == p["witness_count"]
            oc = p["output_count"]
        else:
            assert len(instances) == cg[4]["instance_count"] and len(witnesses) == cg[4]["witness_count"]
            oc = cg[4]["output_count"]
        self.instance_vals.extend(instances)
        self.witness_vals.extend(witnesses)
        outs, self.free_id = _multiple_alloc(self.free_id, oc)
        self.gates.append(_complex_with_output(cg, outs))
        return outs

    def new_function_builder(self, name, output_count, input_count):
        return FunctionBuilder(self.known, name, output_count, input_count)

    def new_switch_builder(self, output_count):
        return SwitchBuilder(self.known, output_count)

    def push_function(self, f: ir.Function):
        assert f.name not in self.known
        self.known[f.name] = dict(input_count=f.input_count, output_count=f.output_count,
                                  instance_count=f.instance_count, witness_count=f.witness_count)
        self.functions.append(f)

    def finish(self):
        msgs = []
        if self.instance_vals:
            msgs.append(ir.Instance(self.header, list(self.instance_vals)))
        if self.witness_vals:
            msgs.append(ir.Witness(self.header, list(self.witness_vals)))
        if self.gates or self.functions:
            msgs.append(ir.Relation(self.header, self.gateset, self.features, list(self.functions),
                                    list(self.gates)))
        return msgs


def _assert_equal_witness(b):
    fb = b.new_function_builder("assert_equal_witness", 0, 1)
    iw = fb.input_wire_ids()
    w = fb.create_gate("Witness", None)
    nw = fb.create_gate("MulConstant", w, bytes([100]))
    ar = fb.create_gate("Add", iw[0], nw)
    fb.create_gate("AssertZero", ar)
    return fb.finish([])


def _custom_sub_iw(b):
    fb = b.new_function_builder("custom_sub", 2, 2)
    iw = fb.input_wire_ids()
    inst = fb.create_gate("Instance", None)
    wit = fb.create_gate("Witness", None)
    ni = fb.create_gate("MulConstant", inst, bytes([100]))
    nw = fb.create_gate("MulConstant", wit, bytes([100]))
    o0 = fb.create_gate("Add", iw[0], ni)
    o1 = fb.create_gate("Add", iw[1], nw)
    return fb.finish([o0, o1])


def builder_with_function():
    """builder.rs:726-805"""
    b = GateBuilder(example_header(), ir.ARITH, ir.FOR_FUNCTION_SWITCH)
    fb = b.new_function_builder("custom_sub", 2, 4)
    iw = fb.input_wire_ids()
    n2 = fb.create_gate("MulConstant", iw[2], bytes([100]))
    n3 = fb.create_gate("MulConstant", iw[3], bytes([100]))
    o0 = fb.create_gate("Add", iw[0], n2)
    o1 = fb.create_gate("Add", iw[1], n3)
    b.push_function(fb.finish([o0, o1]))
    ids = [b.create_gate("Constant", bytes([v])) for v in (40, 30, 10, 5)]
    out = ir.expand_wirelist(b.create_complex_gate(("Call", "custom_sub", wirelist(*ids)), [], []))
    w0 = b.create_gate("Witness", bytes([30]))
    w1 = b.create_gate("Witness", bytes([25]))
    nw0 = b.create_gate("MulConstant", w0, bytes([100]))
    nw1 = b.create_gate("MulConstant", w1, bytes([100]))
    r0 = b.create_gate("Add", out[0], nw0)
    r1 = b.create_gate("Add", out[1], nw1)
    b.create_gate("AssertZero", r0)
    b.create_gate("AssertZero", r1)
    return b.finish()


def builder_with_several_functions():
    """builder.rs:807-898"""
    b = GateBuilder(example_header(), ir.ARITH, ir.FOR_FUNCTION_SWITCH)
    fb = b.new_function_builder("witness_square", 1, 0)
    ww = fb.create_gate("Witness", None)
    ow = fb.create_gate("Mul", ww, ww)
    b.push_function(fb.finish([ow]))
    fb = b.new_function_builder("sub_instance_witness_square", 1, 0)
    iw = fb.create_gate("Instance", None)
    wsw = ir.expand_wirelist(fb.create_complex_gate(("Call", "witness_square", [])))
    nws = fb.create_gate("MulConstant", wsw[0], bytes([100]))
    ow = fb.create_gate("Add", iw, nws)
    b.push_function(fb.finish([ow]))
    out = ir.expand_wirelist(b.create_complex_gate(
        ("Call", "sub_instance_witness_square", []), [bytes([25])], [bytes([5])]))
    b.create_gate("AssertZero", out[0])
    return b.finish()


def builder_switch():
    """builder.rs:900-1054"""
    b = GateBuilder(example_header(), ir.ARITH, ir.FOR_FUNCTION_SWITCH)
    b.push_function(_custom_sub_iw(b))
    fb = b.new_function_builder("custom_add", 2, 2)
    iw = fb.input_wire_ids()
    inst = fb.create_gate("Instance", None)
    wit = fb.create_gate("Witness", None)
    o0 = fb.create_gate("Add", iw[0], inst)
    o1 = fb.create_gate("Add", iw[1], wit)
    w2 = fb.create_gate("Witness", None)
    fb.create_gate("AssertZero", w2)
    b.push_function(fb.finish([o0, o1]))
    b.push_function(_assert_equal_witness(b))
    bi0 = b.create_gate("Constant", bytes([10]))
    bi1 = b.create_gate("Constant", bytes([15]))
    cond = b.create_gate("Constant", bytes([1]))
    sb = b.new_switch_builder(2)
    sb.push_branch_from("custom_sub", wirelist(bi0, bi1), bytes([0]))
    sb.push_branch_from("custom_add", wirelist(bi0, bi1), bytes([1]))
    sw = sb.finish(cond)
    bout = ir.expand_wirelist(b.create_complex_gate(sw, [bytes([5])], [bytes([15]), bytes([0])]))
    b.create_complex_gate(("Call", "assert_equal_witness", wirelist(bout[0])), [], [bytes([15])])
    b.create_complex_gate(("Call", "assert_equal_witness", wirelist(bout[1])), [], [bytes([30])])
    v55 = b.create_gate("Constant", bytes([55]))
    c60 = b.create_gate("Constant", bytes([60]))
    sb = b.new_switch_builder(0)
    sb.push_branch_from("assert_equal_witness", wirelist(v55), bytes([60]))
    b.create_complex_gate(sb.finish(c60), [], [bytes([55])])
    return b.finish()


def builder_switch_nested_in_function():
    """builder.rs:1056-1175"""
    b = GateBuilder(example_header(), ir.ARITH, ir.FOR_FUNCTION_SWITCH)
    b.push_function(_custom_sub_iw(b))
    fb = b.new_function_builder("custom_add", 2, 2)
    iw = fb.input_wire_ids()
    inst = fb.create_gate("Instance", None)
    wit = fb.create_gate("Witness", None)
    o0 = fb.create_gate("Add", iw[0], inst)
    o1 = fb.create_gate("Add", iw[1], wit)
    b.push_function(fb.finish([o0, o1]))
    id0 = b.create_gate("Constant", bytes([40]))
    id1 = b.create_gate("Constant", bytes([30]))
    cond = b.create_gate("Constant", bytes([1]))
    fb = b.new_function_builder("function_with_switch", 2, 3)
    iw = fb.input_wire_ids()
    sb = b.new_switch_builder(2)
    sb.push_branch_from("custom_sub", wirelist(iw[0], iw[1]), bytes([0]))
    sb.push_branch_from("custom_add", wirelist(iw[0], iw[1]), bytes([1]))
    out = ir.expand_wirelist(fb.create_complex_gate(sb.finish(iw[2])))
    b.push_function(fb.finish(out))
    out = ir.expand_wirelist(b.create_complex_gate(
        ("Call", "function_with_switch", wirelist(id0, id1, cond)), [bytes([10])], [bytes([5])]))
    b.push_function(_assert_equal_witness(b))
    b.create_complex_gate(("Call", "assert_equal_witness", wirelist(out[0])), [], [bytes([50])])
    b.create_complex_gate(("Call", "assert_equal_witness", wirelist(out[1])), [], [bytes([35])])
    return b.finish()


# --------------------------------------------------------------- from_r1cs.rs
def r1cs_to_gates(field_maximum: bytes, instance_vars, witness_ids, constraints, witness_values):
    """`FromR1CSConverter` (from_r1cs.rs:27-141) on in-memory R1CS.

    instance_vars : [(id, value_bytes)]      (zkif header.instance_variables)
    witness_ids   : [id] ascending            (zkif header.list_witness_ids())
    constraints   : [(A, B, C)], each LC = [(id, coeff_bytes)]
    witness_values: [value_bytes] in zkif-message order
    Returns [Instance, Witness, Relation] messages (flat, gateset ARITH, SIMPLE).
    """
    p = int.from_bytes(field_maximum, "little") + 1
    header = ir.Header(field_characteristic=ir.le_bytes(p))      # :143-158
    b = GateBuilder(header, ir.ARITH, ir.SIMPLE)
    r1cs_to_ir = {}
    one = b.create_gate("Constant", bytes([1]))                  # :40-42
    assert one == 0
    r1cs_to_ir[0] = one
    minus_one = b.create_gate("Constant", bytes(field_maximum))  # :45-47
    for vid, val in instance_vars:                               # :50-60
        if vid == 0:
            assert int.from_bytes(val, "little") == 1
        else:
            r1cs_to_ir[vid] = b.create_gate("Instance", bytes(val))
    for vid in witness_ids:                                      # :63-66
        r1cs_to_ir[vid] = b.create_gate("Witness", None)

    def build_term(vid, coeff):                                  # :71-92
        val = bytes(coeff) if len(coeff) else bytes([0])
        if vid == 0:
            return b.create_gate("Constant", val)
        c = b.create_gate("Constant", val)
        if vid not in r1cs_to_ir:
            raise ValueError(f"The WireId {vid} has not been defined yet.")
        return b.create_gate("Mul", r1cs_to_ir[vid], c)

    def add_lc(lc):                                              # :94-108
        if len(lc) == 0:
            return b.create_gate("Constant", bytes([0]))
        s = build_term(*lc[0])
        for t in lc[1:]:
            tid = build_term(*t)
            s = b.create_gate("Add", s, tid)
        return s

    for v in witness_values:                                     # :127-136 (test order: witness first)
        b.witness_vals.append(bytes(v))
    for A, B, C in constraints:                                  # :110-125
        sa, sb_, sc = add_lc(A), add_lc(B), add_lc(C)
        prod = b.create_gate("Mul", sa, sb_)
        negc = b.create_gate("Mul", minus_one, sc)
        claim = b.create_gate("Add", prod, negc)
        b.create_gate("AssertZero", claim)
    return b.finish(), r1cs_to_ir


def zkif_example(x=3, y=4, zz=25):
    """zkinterface 1.3.2 `producers::examples` (external crate, restated from its
    published example: x*x = xx ; y*y = yy ; 1*(xx+yy) = zz over p = 101)."""
    field_maximum = bytes([100])
    le4 = lambda v: int(v).to_bytes(4, "little")
    instance_vars = [(1, le4(x)), (2, le4(y)), (3, le4(zz))]
    witness_ids = [4, 5]
    one = bytes([1])
    constraints = [
        ([(1, one)], [(1, one)], [(4, one)]),
        ([(2, one)], [(2, one)], [(5, one)]),
        ([(0, one)], [(4, one), (5, one)], [(3, one)]),
    ]
    witness_values = [le4(x * x), le4(y * y)]
    return field_maximum, instance_vars, witness_ids, constraints, witness_values
