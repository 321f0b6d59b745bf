"""CPU restatement of `rust/src/consumers/flattening.rs` + the slice of `producers/builder.rs` it drives
(oracle, test infrastructure).

`IRFlattener` is a ZKBackend that does not evaluate: every callback appends one SIMPLE gate through a
GateBuilder (`create_gate`, builder.rs:251-290: output wire ids allocated 0, 1, 2, ... for every gate that has
an output), instance / witness values are pushed to their own messages, and the MessageBuilder cuts messages
every `max_len` = 100 000 gates or values (builder.rs:45-49, 79-135).  Pinned on the reference's own tests
`test_validate_flattening` / `test_evaluate_flattening` (flattening.rs:200-252) in tests/test_flatten.py.
"""
from __future__ import annotations

from . import ir

MAX_LEN = 100 * 1000           # builder.rs:69


def to_bytes_le(v: int) -> bytes:
    """num_bigint BigUint::to_bytes_le: minimal length, zero is [0]"""
    return v.to_bytes(max(1, (v.bit_length() + 7) // 8), "little")


class IRFlattener:
    def __init__(self):
        self.header = None
        self.modulus = 0
        self.gateset = None
        self.free_id = 0
        self.instance_msgs, self.witness_msgs, self.relation_msgs = [], [], []
        self._inst, self._wit, self._gates = [], [], []

    # ---- GateBuilder / MessageBuilder ---------------------------------------------------------
    def _need_builder(self):
        if self.header is None:
            raise ir.OraclePanic("Builder has not been properly initialized.")   # flattening.rs:84-86

    def _push_gate(self, gate):                      # builder.rs:93-98
        self._gates.append(gate)
        if len(self._gates) >= MAX_LEN:
            self._flush_relation()

    def _flush_relation(self):                       # builder.rs:118-123
        self.relation_msgs.append(ir.Relation(self.header, self.gateset, ir.SIMPLE, [], self._gates))
        self._gates = []

    def _create(self, name, *args, has_output=True):  # builder.rs:251-290
        self._need_builder()
        if not has_output:
            self._push_gate((name,) + args)
            return None
        out = self.free_id
        self.free_id += 1
        self._push_gate((name, out) + args)
        return out

    def finish(self):                                # builder.rs:124-135
        if self._inst:
            self.instance_msgs.append(ir.Instance(self.header, self._inst))
            self._inst = []
        if self._wit:
            self.witness_msgs.append(ir.Witness(self.header, self._wit))
            self._wit = []
        if self._gates:
            self._flush_relation()
        return self.instance_msgs + self.witness_msgs + self.relation_msgs   # MemorySink -> Source order

    # ---- ZKBackend, flattening.rs:42-191 ----------------------------------------------------------
    @staticmethod
    def from_bytes_le(val: bytes) -> int:
        return int.from_bytes(val, "little")

    def set_field(self, modulus: bytes, degree: int, is_boolean: bool):   # :50-67
        if self.header is None:
            self.header = ir.Header(bytes(modulus), ir.IR_VERSION, degree)
            self.modulus = int.from_bytes(modulus, "little")
            self.gateset = ir.BOOL if is_boolean else ir.ARITH

    def one(self):
        return 1

    def minus_one(self):
        if self.modulus == 0:
            raise ValueError("Modulus is not initiated, used `set_field()` before calling.")
        return self.modulus - 1

    def zero(self):
        return 0

    def copy(self, w):
        return self._create("Copy", w)

    def constant(self, v):
        return self._create("Constant", to_bytes_le(v))

    def assert_zero(self, w):
        self._create("AssertZero", w, has_output=False)

    def add(self, a, b):
        return self._create("Add", a, b)

    def multiply(self, a, b):
        return self._create("Mul", a, b)

    def add_constant(self, a, b):
        return self._create("AddConstant", a, to_bytes_le(b))

    def mul_constant(self, a, b):
        return self._create("MulConstant", a, to_bytes_le(b))

    def and_(self, a, b):
        return self._create("And", a, b)

    def xor(self, a, b):
        return self._create("Xor", a, b)

    def not_(self, a):
        return self._create("Not", a)

    def instance(self, v):                           # builder.rs:258-260: the value goes to the Instance message
        self._need_builder()
        self._inst.append(to_bytes_le(v))
        if len(self._inst) == MAX_LEN:
            self.instance_msgs.append(ir.Instance(self.header, self._inst))
            self._inst = []
        return self._create("Instance")

    def witness(self, v):                            # builder.rs:261-263 (None: no value pushed)
        self._need_builder()
        if v is not None:
            self._wit.append(to_bytes_le(v))
            if len(self._wit) == MAX_LEN:
                self.witness_msgs.append(ir.Witness(self.header, self._wit))
                self._wit = []
        return self._create("Witness")


class ExpandDefinable:
    """`rust/src/consumers/exp_definable.rs:24-139`: a ZKBackend that rewrites the gates outside `gate_mask` and forwards
    everything to an inner IRFlattener.  The `panic!`s become OraclePanic."""

    def __init__(self, gate_mask: int):
        self.inner = IRFlattener()
        self.gate_mask = gate_mask

    def finish(self):
        return self.inner.finish()

    from_bytes_le = staticmethod(IRFlattener.from_bytes_le)

    def set_field(self, modulus, degree, is_boolean):
        self.inner.set_field(modulus, degree, is_boolean)

    def one(self):
        return self.inner.one()

    def minus_one(self):
        return self.inner.minus_one()

    def zero(self):
        return self.inner.zero()

    def copy(self, w):
        return self.inner.copy(w)

    def constant(self, v):
        return self.inner.constant(v)

    def assert_zero(self, w):
        self.inner.assert_zero(w)

    def _has(self, bit):
        return ir.contains_feature(self.gate_mask, bit)

    def add(self, a, b):                              # :58-67
        if not self._has(ir.ADD):
            if not self._has(ir.XOR):
                raise ir.OraclePanic("Cannot replace ADD by XOR if XOR is not supported.")
            return self.inner.xor(a, b)
        return self.inner.add(a, b)

    def multiply(self, a, b):                         # :69-78
        if not self._has(ir.MUL):
            if not self._has(ir.AND):
                raise ir.OraclePanic("Cannot replace MUL by AND if AND is not supported.")
            return self.inner.and_(a, b)
        return self.inner.multiply(a, b)

    def add_constant(self, a, b):                     # :80-87
        if not self._has(ir.ADDC):
            return self.add(a, self.constant(b))
        return self.inner.add_constant(a, b)

    def mul_constant(self, a, b):                     # :89-96
        if not self._has(ir.MULC):
            return self.multiply(a, self.constant(b))
        return self.inner.mul_constant(a, b)

    def and_(self, a, b):                             # :98-107
        if not self._has(ir.AND):
            if not self._has(ir.MUL):
                raise ir.OraclePanic("Cannot replace AND by MUL if MUL is not supported.")
            return self.multiply(a, b)
        return self.inner.and_(a, b)

    def xor(self, a, b):                              # :109-118
        if not self._has(ir.XOR):
            if not self._has(ir.ADD):
                raise ir.OraclePanic("Cannot replace XOR by ADD if ADD is not supported.")
            return self.add(a, b)
        return self.inner.xor(a, b)

    def not_(self, a):                                # :120-129
        if not self._has(ir.NOT):
            if not self._has(ir.ADD):
                raise ir.OraclePanic("Cannot replace NOT by ADD if ADD is not supported.")
            return self.add_constant(a, self.one())
        return self.inner.not_(a)

    def instance(self, v):
        return self.inner.instance(v)

    def witness(self, v):
        return self.inner.witness(v)
