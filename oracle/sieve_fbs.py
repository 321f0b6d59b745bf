"""FlatBuffers reader + writer for `sieve_ir.fbs` (oracle side, test infrastructure).

There is no flatc and no flatbuffers runtime in this image, so both directions
are hand-written against the schema (`sieve_ir.fbs:1-333`) and the vtable slot
constants of the generated Rust (`rust/src/sieve_ir_generated.rs`, `VT_*`).
Reader behaviour (which sub-table is required, the error strings) follows
`rust/src/structs/{gates.rs:60-259, function.rs:29-46,132-172, header.rs:37-56,
value.rs:14-29, wire.rs:70-157, iterators.rs:36-326, relation.rs:43-85,
message.rs:15-36}`.  Framing follows `rust/src/consumers/utils.rs:6-41`:
each message is `[u32 LE size][size bytes]`, identifier "siev" at bytes 4..8 of
the body.

The writer exists so tests can produce `.sieve` inputs for the product's C++
reader; it is pinned by (a) the reference's own binary fixtures re-serialised
and re-read and (b) write->read round trips of every fixture relation.
"""
from __future__ import annotations

import struct
from typing import List, Tuple

from . import ir

# DirectiveSet discriminants, sieve_ir_generated.rs:422-442
DS = ["NONE", "Constant", "AssertZero", "Copy", "Add", "Mul", "AddConstant", "MulConstant",
      "And", "Xor", "Not", "Instance", "Witness", "Free", "Call", "AnonCall", "Switch", "For"]
DS_ID = {n: i for i, n in enumerate(DS)}
MSG_RELATION, MSG_INSTANCE, MSG_WITNESS = 1, 2, 3  # sieve_ir_generated.rs:23-29
ITER = ["NONE", "Const", "Name", "Add", "Sub", "Mul", "DivConst"]  # :218-227
ITER_ID = {n: i for i, n in enumerate(ITER)}


class FbsError(Exception):
    pass


# =====================================================================
#                               READER
# =====================================================================
class _T:
    """A table view: buffer + absolute position."""
    __slots__ = ("b", "pos", "vt", "vtlen")

    def __init__(self, b, pos):
        self.b = b
        self.pos = pos
        self.vt = pos - struct.unpack_from("<i", b, pos)[0]
        self.vtlen = struct.unpack_from("<H", b, self.vt)[0]

    def _off(self, slot):
        if slot >= self.vtlen:
            return 0
        return struct.unpack_from("<H", self.b, self.vt + slot)[0]

    def u8(self, slot, default=0):
        o = self._off(slot)
        return self.b[self.pos + o] if o else default

    def u32(self, slot, default=0):
        o = self._off(slot)
        return struct.unpack_from("<I", self.b, self.pos + o)[0] if o else default

    def u64(self, slot, default=0):
        o = self._off(slot)
        return struct.unpack_from("<Q", self.b, self.pos + o)[0] if o else default

    def _indirect(self, slot):
        o = self._off(slot)
        if not o:
            return None
        p = self.pos + o
        return p + struct.unpack_from("<I", self.b, p)[0]

    def table(self, slot):
        p = self._indirect(slot)
        return None if p is None else _T(self.b, p)

    def string(self, slot):
        p = self._indirect(slot)
        if p is None:
            return None
        n = struct.unpack_from("<I", self.b, p)[0]
        return bytes(self.b[p + 4:p + 4 + n]).decode("utf-8")

    def bytes_vec(self, slot):
        p = self._indirect(slot)
        if p is None:
            return None
        n = struct.unpack_from("<I", self.b, p)[0]
        return bytes(self.b[p + 4:p + 4 + n])

    def table_vec(self, slot):
        p = self._indirect(slot)
        if p is None:
            return None
        n = struct.unpack_from("<I", self.b, p)[0]
        out = []
        for i in range(n):
            q = p + 4 + 4 * i
            out.append(_T(self.b, q + struct.unpack_from("<I", self.b, q)[0]))
        return out


def _req(x, msg):
    if x is None:
        raise FbsError(msg)
    return x


def _wire_id(t, msg):
    return _req(t, msg).u64(4)


def _value(t):
    return _req(_req(t, "Missing value").bytes_vec(4), "Missing value")


def _wirelist(t):
    out = []
    for el in _req(t.table_vec(4), "Missing wire list elements"):
        ty = el.u8(4)
        if ty == 1:
            out.append(ir.Wire(_req(el.table(6), "Missing wire").u64(4)))
        elif ty == 2:
            r = _req(el.table(6), "Missing range")
            out.append(ir.WireRange(_wire_id(r.table(4), "Missing start value in range"),
                                    _wire_id(r.table(6), "Missing end value in range")))
        else:
            raise FbsError("Unknown type in WireListElement")
    return out


def _iterexpr(t):
    ty = t.u8(4)
    if ty == 0 or ty >= len(ITER):
        raise FbsError("Unknown Iterator Expression type")
    v = _req(t.table(6), "Missing iterator expression value")
    k = ITER[ty]
    if k == "Const":
        return ("Const", v.u64(4))
    if k == "Name":
        return ("Name", _req(v.string(4), "IterExpr: No name given"))
    if k in ("Add", "Sub", "Mul"):
        return (k, _iterexpr(_req(v.table(4), "Missing left operand")),
                _iterexpr(_req(v.table(6), "Missing right operand")))
    return ("DivConst", _iterexpr(_req(v.table(4), "Missing numerator")), v.u64(6))


def _iterexpr_list(t):
    out = []
    for el in _req(t.table_vec(4), "Missing iterexpr elements"):
        ty = el.u8(4)
        if ty == 1:
            out.append(("Single", _iterexpr(_req(el.table(6), "Missing element"))))
        elif ty == 2:
            r = _req(el.table(6), "Missing element")
            out.append(("Range", _iterexpr(_req(r.table(4), "Missing first value of range")),
                        _iterexpr(_req(r.table(6), "Missing last value of range"))))
        else:
            raise FbsError("Unknown type in IterExprWireListElement")
    return out


def _case_invoke(t):
    ty = t.u8(4)
    if ty == 1:
        c = _req(t.table(6), "Missing invocation")
        return ("AbstractGateCall", _req(c.string(4), "Missing function name."),
                _wirelist(_req(c.table(6), "Missing inputs")))
    if ty == 2:
        c = _req(t.table(6), "Missing invocation")
        sub = _req(c.table_vec(10), "Missing implementation")
        return ("AbstractAnonCall", _wirelist(_req(c.table(4), "Missing inputs")),
                c.u64(6), c.u64(8), [_gate(g) for g in sub])
    raise FbsError("No directive type")


def _gate(d):
    ty = d.u8(4)
    if ty == 0 or ty >= len(DS):
        raise FbsError("No gate type")
    k = DS[ty]
    g = _req(d.table(6), "Missing directive")
    if k == "Constant":
        return (k, _wire_id(g.table(4), "Missing output"), _req(g.bytes_vec(6), "Missing constant"))
    if k == "AssertZero":
        return (k, _wire_id(g.table(4), "Missing input"))
    if k in ("Copy", "Not"):
        return (k, _wire_id(g.table(4), "Missing output"), _wire_id(g.table(6), "Missing input"))
    if k in ("Add", "Mul", "And", "Xor"):
        return (k, _wire_id(g.table(4), "Missing output"), _wire_id(g.table(6), "Missing left input"),
                _wire_id(g.table(8), "Missing right input"))
    if k in ("AddConstant", "MulConstant"):
        return (k, _wire_id(g.table(4), "Missing output"), _wire_id(g.table(6), "Missing input"),
                _req(g.bytes_vec(8), "Missing constant"))
    if k in ("Instance", "Witness"):
        return (k, _wire_id(g.table(4), "Missing output"))
    if k == "Free":
        last = g.table(6)
        return (k, _wire_id(g.table(4), "Missing first wire"), None if last is None else last.u64(4))
    if k == "Call":
        return (k, _req(g.string(4), "Missing function name."),
                _wirelist(_req(g.table(6), "Missing outputs")),
                _wirelist(_req(g.table(8), "Missing inputs")))
    if k == "AnonCall":
        inner = _req(g.table(6), "Missing inner AbstractAnonCall")
        return (k, _wirelist(_req(g.table(4), "Missing output wires")),
                _wirelist(_req(inner.table(4), "Missing input wires")),
                inner.u64(6), inner.u64(8),
                [_gate(x) for x in _req(inner.table_vec(10), "Missing subcircuit")])
    if k == "Switch":
        cases = [_value(c) for c in _req(g.table_vec(8), "Missing cases values")]
        return (k, _wire_id(g.table(4), "Missing condition wire."),
                _wirelist(_req(g.table(6), "Missing output wires")), cases,
                [_case_invoke(b) for b in _req(g.table_vec(10), "Missing branches")])
    if k == "For":
        outs = _wirelist(_req(g.table(4), "missing output list"))
        bt = g.u8(12)
        if bt == 1:
            b = _req(g.table(14), "Missing body")
            body = ("IterExprCall", _req(b.string(4), "Missing function in function name"),
                    _iterexpr_list(_req(b.table(6), "missing output list")),
                    _iterexpr_list(_req(b.table(8), "missing input list")))
        elif bt == 2:
            b = _req(g.table(14), "Missing body")
            body = ("IterExprAnonCall", _iterexpr_list(_req(b.table(4), "missing output list")),
                    _iterexpr_list(_req(b.table(6), "missing input list")),
                    b.u64(8), b.u64(10), [_gate(x) for x in _req(b.table_vec(12), "Missing body")])
        else:
            raise FbsError("Unknown body type")
        return (k, _req(g.string(6), "Missing iterator name"), g.u64(8), g.u64(10), outs, body)
    raise FbsError("No gate type")


def _header(t):
    t = _req(t, "Missing header")
    return ir.Header(
        version=_req(t.string(4), "Missing version"),
        field_characteristic=_value(_req(t.table(6), "Missing field characteristic")),
        field_degree=t.u32(8))


def _function(t):
    body = _req(t.table_vec(14), "Missing reference implementation")
    return ir.Function(name=_req(t.string(4), "Missing name"), output_count=t.u64(6),
                       input_count=t.u64(8), instance_count=t.u64(10), witness_count=t.u64(12),
                       body=[_gate(g) for g in body])


def read_message(buf: bytes):
    """Parse ONE size-prefixed message (`Message::try_from`, message.rs:15-36)."""
    b = memoryview(buf)
    root = _T(b, 4 + struct.unpack_from("<I", b, 4)[0])
    mt = root.u8(4)
    m = root.table(6)
    if mt == MSG_INSTANCE:
        m = _req(m, "Missing message")
        return ir.Instance(header=_header(m.table(4)),
                           common_inputs=[_value(v) for v in _req(m.table_vec(6), "Missing common_input")])
    if mt == MSG_WITNESS:
        m = _req(m, "Missing message")
        return ir.Witness(header=_header(m.table(4)),
                          short_witness=[_value(v) for v in _req(m.table_vec(6), "Missing short_witness")])
    if mt == MSG_RELATION:
        m = _req(m, "Missing message")
        dirs = _req(m.table_vec(12), "Missing directives")
        fns = m.table_vec(10)
        functions = [_function(f) for f in fns] if fns is not None else []
        return ir.Relation(
            header=_header(m.table(4)),
            gate_mask=ir.parse_gate_set(_req(m.string(6), "Missing gateset description")),
            feat_mask=ir.parse_feature_toggle(_req(m.string(8), "Missing feature toggles")),
            functions=functions, gates=[_gate(d) for d in dirs])
    raise FbsError("Invalid message type")


def split_messages(buf: bytes) -> List[bytes]:
    """Size-prefixed framing: rust/src/consumers/utils.rs:6-41."""
    out = []
    pos = 0
    while pos + 4 <= len(buf):
        size = struct.unpack_from("<I", buf, pos)[0]
        if size == 0:
            break
        out.append(bytes(buf[pos:pos + 4 + size]))
        pos += 4 + size
    return out


def read_messages(buf: bytes):
    return [read_message(m) for m in split_messages(buf)]


# =====================================================================
#                               WRITER
# =====================================================================
# A forward writer: a table is emitted (vtable first, then the table with
# placeholder uoffsets), then its children are appended at higher addresses and
# the placeholders are patched.  uoffsets are unsigned and relative to the
# slot, so children must lie above their parent — which this order guarantees.
class _W:
    def __init__(self):
        self.b = bytearray()

    def align(self, n):
        while len(self.b) % n:
            self.b.append(0)

    # fields: list of (slot, kind, value); kind in {"u8","u32","u64","ref"}
    # for "ref", value is a thunk(writer)->absolute position of the child.
    def table(self, fields):
        fields = [f for f in fields if f is not None]
        size = {"u8": 1, "u32": 4, "u64": 8, "ref": 4}
        # lay out: soffset (4), then fields largest-first
        order = sorted(fields, key=lambda f: -size[f[1]])
        has8 = any(f[1] == "u64" for f in fields)
        off = 4
        offs = {}
        for slot, kind, _ in order:
            s = size[kind]
            off = (off + s - 1) // s * s
            offs[slot] = off
            off += s
        tbl_len = (off + 3) // 4 * 4
        max_slot = max([f[0] for f in fields], default=2)
        vt_len = max_slot + 2
        # place vtable so that the table lands 4- (or 8-) aligned
        al = 8 if has8 else 4
        self.align(2)
        while (len(self.b) + vt_len) % al:
            self.b.append(0)
        vt_pos = len(self.b)
        vt = bytearray(vt_len)
        struct.pack_into("<HH", vt, 0, vt_len, tbl_len)
        for slot, _, _ in fields:
            struct.pack_into("<H", vt, slot, offs[slot])
        self.b += vt
        t_pos = len(self.b)
        self.b += bytes(tbl_len)
        struct.pack_into("<i", self.b, t_pos, t_pos - vt_pos)
        refs = []
        for slot, kind, val in fields:
            p = t_pos + offs[slot]
            if kind == "u8":
                self.b[p] = val
            elif kind == "u32":
                struct.pack_into("<I", self.b, p, val)
            elif kind == "u64":
                struct.pack_into("<Q", self.b, p, val)
            else:
                refs.append((p, val))
        for p, thunk in refs:
            child = thunk(self)
            struct.pack_into("<I", self.b, p, child - p)
        return t_pos

    def string(self, s: str):
        data = s.encode("utf-8")
        self.align(4)
        pos = len(self.b)
        self.b += struct.pack("<I", len(data)) + data + b"\0"
        return pos

    def bytes_vec(self, data: bytes):
        self.align(4)
        pos = len(self.b)
        self.b += struct.pack("<I", len(data)) + bytes(data)
        return pos

    def table_vec(self, thunks):
        self.align(4)
        pos = len(self.b)
        self.b += struct.pack("<I", len(thunks)) + bytes(4 * len(thunks))
        for i, th in enumerate(thunks):
            p = pos + 4 + 4 * i
            child = th(self)
            struct.pack_into("<I", self.b, p, child - p)
        return pos


def _w_wire(i):
    # flatc omits scalars equal to their default (id 0) — same here.
    return lambda w: w.table([(4, "u64", i)] if i != 0 else [])


def _w_value(v):
    return lambda w: w.table([(4, "ref", lambda w2: w2.bytes_vec(v))])


def _w_str(s):
    return lambda w: w.string(s)


def _w_bytes(b):
    return lambda w: w.bytes_vec(b)


def _w_wirelist(wl):
    def el(e):
        if e[0] == "Wire":
            return lambda w: w.table([(4, "u8", 1), (6, "ref", _w_wire(e[1]))])
        return lambda w: w.table([(4, "u8", 2), (6, "ref", lambda w2: w2.table(
            [(4, "ref", _w_wire(e[1])), (6, "ref", _w_wire(e[2]))]))])
    return lambda w: w.table([(4, "ref", lambda w2: w2.table_vec([el(e) for e in wl]))])


def _w_iterexpr(e):
    k = e[0]
    if k == "Const":
        inner = lambda w: w.table([(4, "u64", e[1])] if e[1] else [])
    elif k == "Name":
        inner = lambda w: w.table([(4, "ref", _w_str(e[1]))])
    elif k in ("Add", "Sub", "Mul"):
        inner = lambda w: w.table([(4, "ref", _w_iterexpr(e[1])), (6, "ref", _w_iterexpr(e[2]))])
    else:
        inner = lambda w: w.table([(4, "ref", _w_iterexpr(e[1]))] + ([(6, "u64", e[2])] if e[2] else []))
    return lambda w: w.table([(4, "u8", ITER_ID[k]), (6, "ref", inner)])


def _w_iterexpr_list(lst):
    def el(e):
        if e[0] == "Single":
            return lambda w: w.table([(4, "u8", 1), (6, "ref", _w_iterexpr(e[1]))])
        return lambda w: w.table([(4, "u8", 2), (6, "ref", lambda w2: w2.table(
            [(4, "ref", _w_iterexpr(e[1])), (6, "ref", _w_iterexpr(e[2]))]))])
    return lambda w: w.table([(4, "ref", lambda w2: w2.table_vec([el(e) for e in lst]))])


def _u64f(slot, v):
    return (slot, "u64", v) if v else None


def _w_gates(gates):
    return lambda w: w.table_vec([_w_gate(g) for g in gates])


def _w_case(c):
    if c[0] == "AbstractGateCall":
        inner = lambda w: w.table([(4, "ref", _w_str(c[1])), (6, "ref", _w_wirelist(c[2]))])
        return lambda w: w.table([(4, "u8", 1), (6, "ref", inner)])
    inner = lambda w: w.table([(4, "ref", _w_wirelist(c[1])), _u64f(6, c[2]), _u64f(8, c[3]),
                               (10, "ref", _w_gates(c[4]))])
    return lambda w: w.table([(4, "u8", 2), (6, "ref", inner)])


def _w_gate(g):
    k = g[0]
    if k == "Constant":
        inner = [(4, "ref", _w_wire(g[1])), (6, "ref", _w_bytes(g[2]))]
    elif k == "AssertZero":
        inner = [(4, "ref", _w_wire(g[1]))]
    elif k in ("Copy", "Not"):
        inner = [(4, "ref", _w_wire(g[1])), (6, "ref", _w_wire(g[2]))]
    elif k in ("Add", "Mul", "And", "Xor"):
        inner = [(4, "ref", _w_wire(g[1])), (6, "ref", _w_wire(g[2])), (8, "ref", _w_wire(g[3]))]
    elif k in ("AddConstant", "MulConstant"):
        inner = [(4, "ref", _w_wire(g[1])), (6, "ref", _w_wire(g[2])), (8, "ref", _w_bytes(g[3]))]
    elif k in ("Instance", "Witness"):
        inner = [(4, "ref", _w_wire(g[1]))]
    elif k == "Free":
        inner = [(4, "ref", _w_wire(g[1]))] + ([(6, "ref", _w_wire(g[2]))] if g[2] is not None else [])
    elif k == "Call":
        inner = [(4, "ref", _w_str(g[1])), (6, "ref", _w_wirelist(g[2])), (8, "ref", _w_wirelist(g[3]))]
    elif k == "AnonCall":
        sub = lambda w: w.table([(4, "ref", _w_wirelist(g[2])), _u64f(6, g[3]), _u64f(8, g[4]),
                                 (10, "ref", _w_gates(g[5]))])
        inner = [(4, "ref", _w_wirelist(g[1])), (6, "ref", sub)]
    elif k == "Switch":
        inner = [(4, "ref", _w_wire(g[1])), (6, "ref", _w_wirelist(g[2])),
                 (8, "ref", lambda w: w.table_vec([_w_value(c) for c in g[3]])),
                 (10, "ref", lambda w: w.table_vec([_w_case(c) for c in g[4]]))]
    elif k == "For":
        body = g[5]
        if body[0] == "IterExprCall":
            bt = 1
            b = lambda w: w.table([(4, "ref", _w_str(body[1])), (6, "ref", _w_iterexpr_list(body[2])),
                                   (8, "ref", _w_iterexpr_list(body[3]))])
        else:
            bt = 2
            b = lambda w: w.table([(4, "ref", _w_iterexpr_list(body[1])), (6, "ref", _w_iterexpr_list(body[2])),
                                   _u64f(8, body[3]), _u64f(10, body[4]), (12, "ref", _w_gates(body[5]))])
        inner = [(4, "ref", _w_wirelist(g[4])), (6, "ref", _w_str(g[1])), _u64f(8, g[2]), _u64f(10, g[3]),
                 (12, "u8", bt), (14, "ref", b)]
    else:
        raise FbsError(f"cannot serialise gate {k}")
    return lambda w: w.table([(4, "u8", DS_ID[k]), (6, "ref", lambda w2: w2.table(inner))])


def _w_header(h: ir.Header):
    return lambda w: w.table([(4, "ref", _w_str(h.version)), (6, "ref", _w_value(h.field_characteristic)),
                              (8, "u32", h.field_degree) if h.field_degree else None])


def _w_function(f: ir.Function):
    return lambda w: w.table([(4, "ref", _w_str(f.name)), _u64f(6, f.output_count), _u64f(8, f.input_count),
                              _u64f(10, f.instance_count), _u64f(12, f.witness_count),
                              (14, "ref", _w_gates(f.body))])


def write_message(msg) -> bytes:
    """Serialise one message with its 4-byte size prefix and the "siev" identifier."""
    w = _W()
    w.b += bytes(12)  # [size][root uoffset][identifier]
    w.b[8:12] = b"siev"
    if isinstance(msg, ir.Instance):
        mt = MSG_INSTANCE
        body = lambda w2: w2.table([(4, "ref", _w_header(msg.header)),
                                    (6, "ref", lambda w3: w3.table_vec([_w_value(v) for v in msg.common_inputs]))])
    elif isinstance(msg, ir.Witness):
        mt = MSG_WITNESS
        body = lambda w2: w2.table([(4, "ref", _w_header(msg.header)),
                                    (6, "ref", lambda w3: w3.table_vec([_w_value(v) for v in msg.short_witness]))])
    elif isinstance(msg, ir.Relation):
        mt = MSG_RELATION
        body = lambda w2: w2.table([
            (4, "ref", _w_header(msg.header)),
            (6, "ref", _w_str(ir.create_gateset_string(msg.gate_mask))),
            (8, "ref", _w_str(ir.create_feature_string(msg.feat_mask))),
            (10, "ref", lambda w3: w3.table_vec([_w_function(f) for f in msg.functions])),
            (12, "ref", _w_gates(msg.gates))])
    else:
        raise FbsError("unknown message")
    root = w.table([(4, "u8", mt), (6, "ref", body)])
    w.align(4)
    struct.pack_into("<I", w.b, 4, root - 4)
    struct.pack_into("<I", w.b, 0, len(w.b) - 4)
    return bytes(w.b)


def write_messages(msgs) -> bytes:
    return b"".join(write_message(m) for m in msgs)
