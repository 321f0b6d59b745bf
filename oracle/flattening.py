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
