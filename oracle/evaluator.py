"""CPU restatement of `rust/src/consumers/evaluator.rs` (oracle, test infrastructure).

Every function cites the reference lines it follows.  Field arithmetic in the
reference is `num_bigint::BigUint` (`+ * % & ^`, crate num-bigint 0.3.0, not
vendored); these are exact integer operations, so Python `int` is an exact
stand-in.  Pinned by the reference's golden vectors in
`tests/test_oracle_golden.py`.

Conditions on which the Rust code PANICS raise `OraclePanic`.
"""
from __future__ import annotations

from collections import deque
from typing import Dict, List, Optional

from . import ir
from .ir import OraclePanic, expand_wirelist, evaluate_iterexpr_list


class EvalError(Exception):
    """`Err(Box<dyn Error>)` of the reference; str(e) is the message."""


# ------------------------------------------------------------------------
# ZKBackend implementations
# ------------------------------------------------------------------------
class PlaintextBackend:
    """evaluator.rs:848-947.  Wire = FieldElement = unbounded unsigned integer."""

    def __init__(self):
        self.m = 0

    @staticmethod
    def from_bytes_le(val: bytes) -> int:           # :862-864 (NO reduction)
        return int.from_bytes(val, "little")

    def set_field(self, modulus: bytes, degree: int, is_boolean: bool):   # :866-875
        self.m = int.from_bytes(modulus, "little")
        if self.m == 0:
            raise EvalError("Modulus cannot be zero.")
        if degree != 1:
            raise EvalError("Field should be of degree 1")

    def one(self):                                  # :877-879
        return 1

    def minus_one(self):                            # :881-886
        if self.m == 0:
            raise EvalError("Modulus is not initiated, used `set_field()` before calling.")
        return self.m - 1

    def zero(self):                                 # :888-890
        return 0

    def copy(self, w):                              # :892-894
        return w

    def constant(self, v):                          # :896-898 (kept RAW)
        return v

    def assert_zero(self, w):                       # :900-906 (raw integer test)
        if w != 0:
            raise EvalError("AssertZero failed")

    def add(self, a, b):                            # :908-910
        return (a + b) % self.m

    def multiply(self, a, b):                       # :912-914
        return (a * b) % self.m

    def add_constant(self, a, b):                   # :916-918
        return (a + b) % self.m

    def mul_constant(self, a, b):                   # :920-922
        return (a * b) % self.m

    def and_(self, a, b):                           # :924-926
        return (a & b) % self.m

    def xor(self, a, b):                            # :928-930
        return (a ^ b) % self.m

    def not_(self, a):                              # :932-938
        return 1 if a == 0 else 0

    def instance(self, v):                          # :940-942
        return self.constant(v)

    def witness(self, v):                           # :944-946
        if v is None:
            raise OraclePanic("Missing witness value for PlaintextBackend")
        return self.constant(v)


class TracingBackend(PlaintextBackend):
    """PlaintextBackend that also records every callback, in call order.

    Same idea as the reference's `IRFlattener` (consumers/flattening.rs:83-190):
    each callback yields a fresh sequential id.  A wire is `(ssa_id, value)`;
    `trace[i] = (kind, operand_ssa_ids, value)`.  Used by the tests to compare
    the product's flattened program op-for-op and value-for-value.
    """

    def __init__(self):
        super().__init__()
        self.trace: List[tuple] = []
        self.asserts: List[tuple] = []   # (trace position, ssa id, ok)

    def _new(self, kind, ops, value):
        self.trace.append((kind, tuple(ops), value))
        return (len(self.trace) - 1, value)

    def copy(self, w):
        return self._new("copy", [w[0]], w[1])

    def constant(self, v):
        return self._new("constant", [], v)

    def assert_zero(self, w):
        ok = w[1] == 0
        self.asserts.append((len(self.trace), w[0], ok))
        if not ok:
            raise EvalError("AssertZero failed")

    def add(self, a, b):
        return self._new("add", [a[0], b[0]], (a[1] + b[1]) % self.m)

    def multiply(self, a, b):
        return self._new("mul", [a[0], b[0]], (a[1] * b[1]) % self.m)

    def add_constant(self, a, b):
        return self._new("addc", [a[0]], (a[1] + b) % self.m)

    def mul_constant(self, a, b):
        return self._new("mulc", [a[0]], (a[1] * b) % self.m)

    def and_(self, a, b):
        return self._new("and", [a[0], b[0]], (a[1] & b[1]) % self.m)

    def xor(self, a, b):
        return self._new("xor", [a[0], b[0]], (a[1] ^ b[1]) % self.m)

    def not_(self, a):
        return self._new("not", [a[0]], 1 if a[1] == 0 else 0)

    def instance(self, v):
        return self._new("instance", [], v)

    def witness(self, v):
        if v is None:
            raise OraclePanic("Missing witness value for PlaintextBackend")
        return self._new("witness", [], v)

    def counts(self) -> Dict[str, int]:
        c: Dict[str, int] = {}
        for k, _, _ in self.trace:
            c[k] = c.get(k, 0) + 1
        c["assert_zero"] = len(self.asserts)
        return c


# ------------------------------------------------------------------------
# helpers evaluator.rs:78-126
# ------------------------------------------------------------------------
def as_mul(backend, a, b, is_bool):                 # :80-91
    return backend.and_(a, b) if is_bool else backend.multiply(a, b)


def as_add(backend, a, b, is_bool):                 # :95-106
    return backend.xor(a, b) if is_bool else backend.add(a, b)


def as_negate(backend, w, is_bool):                 # :109-116
    return backend.copy(w) if is_bool else backend.mul_constant(w, backend.minus_one())


def as_add_one(backend, w, is_bool):                # :119-126
    return backend.not_(w) if is_bool else backend.add_constant(w, backend.one())


def exp(backend, base, exponent: int, modulus: int, is_bool):   # :801-820
    if exponent == 1:
        return backend.copy(base)
    if exponent == 0:
        # the Rust recursion never terminates for exponent 0 (modulus 1)
        raise OraclePanic("exp: exponent 0 recurses forever in the reference")
    previous = exp(backend, base, exponent >> 1, modulus, is_bool)
    ret = as_mul(backend, previous, previous, is_bool)
    if exponent & 1:
        return as_mul(backend, ret, base, is_bool)
    return ret


def compute_weight(backend, case: bytes, condition, modulus: int, is_bool):   # :823-839
    case_wire = backend.constant(backend.from_bytes_le(case))
    exponent = modulus - 1
    minus_cond = as_negate(backend, condition, is_bool)
    base = as_add(backend, case_wire, minus_cond, is_bool)
    base_to_exp = exp(backend, base, exponent, modulus, is_bool)
    right = as_negate(backend, base_to_exp, is_bool)
    return as_add_one(backend, right, is_bool)


def _get(scope, wid):                               # :787-791
    if wid not in scope:
        raise EvalError(f"No value given for wire_{wid}")
    return scope[wid]


def _set(scope, wid, wire):                         # :775-785 (inserts even on error)
    had = wid in scope
    scope[wid] = wire
    if had:
        raise EvalError(f"Wire_{wid} already has a value in this scope.")


def _remove(scope, wid):                            # :793-797
    if wid not in scope:
        raise EvalError(f"No value given for wire_{wid}")
    return scope.pop(wid)


class _FunctionDeclaration:                         # :130-136
    __slots__ = ("subcircuit", "instance_nbr", "witness_nbr", "output_count", "input_count")

    def __init__(self, f: ir.Function):
        self.subcircuit = f.body
        self.instance_nbr = f.instance_count
        self.witness_nbr = f.witness_count
        self.output_count = f.output_count
        self.input_count = f.input_count


class Evaluator:
    """evaluator.rs:158-753"""

    def __init__(self):                             # :172-185
        self.values: Dict[int, object] = {}
        self.modulus = 0
        self.instance_queue = deque()
        self.witness_queue = deque()
        self.is_boolean = False
        self.known_functions: Dict[str, _FunctionDeclaration] = {}
        self.verified_at_least_one_gate = False
        self.found_error: Optional[str] = None

    @classmethod
    def from_messages(cls, messages, backend):      # :191-195
        ev = cls()
        for m in messages:
            ev.ingest_message(m, backend)
        return ev

    def get_violations(self) -> List[str]:          # :199-208
        v = []
        if not self.verified_at_least_one_gate:
            v.append("Did not receive any gate to verify.")
        if self.found_error is not None:
            v.append(self.found_error)
        return v

    def ingest_message(self, msg, backend):         # :213-230
        if self.found_error is not None:
            return
        try:
            if isinstance(msg, ir.Instance):
                self.ingest_instance(msg, backend)
            elif isinstance(msg, ir.Witness):
                self.ingest_witness(msg, backend)
            else:
                self.ingest_relation(msg, backend)
        except EvalError as e:
            self.found_error = str(e)

    def ingest_header(self, header: ir.Header):     # :232-235
        self.modulus = int.from_bytes(header.field_characteristic, "little")

    def ingest_instance(self, instance: ir.Instance, backend=PlaintextBackend):   # :239-246
        self.ingest_header(instance.header)
        for v in instance.common_inputs:
            self.instance_queue.append(backend.from_bytes_le(v))

    def ingest_witness(self, witness: ir.Witness, backend=PlaintextBackend):      # :250-257
        self.ingest_header(witness.header)
        for v in witness.short_witness:
            self.witness_queue.append(backend.from_bytes_le(v))

    def ingest_relation(self, relation: ir.Relation, backend):                    # :260-303
        self.ingest_header(relation.header)
        self.is_boolean = ir.contains_feature(relation.gate_mask, ir.BOOL)
        backend.set_field(relation.header.field_characteristic, relation.header.field_degree,
                          self.is_boolean)
        if len(relation.gates) > 0:
            self.verified_at_least_one_gate = True
        for f in relation.functions:
            self.known_functions[f.name] = _FunctionDeclaration(f)
        known_iterators: Dict[str, int] = {}
        for gate in relation.gates:
            self._ingest_gate(gate, backend, self.values, known_iterators,
                              self.instance_queue, self.witness_queue, None)

    def get(self, wid):                             # :750-752
        return _get(self.values, wid)

    # -- :318-691 ---------------------------------------------------------
    def _ingest_gate(self, gate, backend, scope, known_iterators, instances, witnesses, weight):
        k = gate[0]
        is_boolean = self.is_boolean
        modulus = self.modulus
        kf = self.known_functions

        if k == "Constant":                         # :345-348
            _, out, value = gate
            _set(scope, out, backend.constant(backend.from_bytes_le(value)))

        elif k == "AssertZero":                     # :350-364
            inp = gate[1]
            w = _get(scope, inp)
            z = as_mul(backend, weight, w, is_boolean) if weight is not None else backend.copy(w)
            try:
                backend.assert_zero(z)
            except EvalError:
                raise EvalError(f"Wire_{inp} (may be weighted) should be 0, while it is not")

        elif k == "Copy":                           # :366-370
            _, out, inp = gate
            _set(scope, out, backend.copy(_get(scope, inp)))

        elif k in ("Add", "Mul", "And", "Xor"):     # :372-384, 400-412
            _, out, left, right = gate
            l = _get(scope, left)
            r = _get(scope, right)
            fn = {"Add": backend.add, "Mul": backend.multiply, "And": backend.and_, "Xor": backend.xor}[k]
            _set(scope, out, fn(l, r))

        elif k in ("AddConstant", "MulConstant"):   # :386-398
            _, out, inp, constant = gate
            l = _get(scope, inp)
            r = backend.from_bytes_le(constant)
            fn = backend.add_constant if k == "AddConstant" else backend.mul_constant
            _set(scope, out, fn(l, r))

        elif k == "Not":                            # :414-418
            _, out, inp = gate
            _set(scope, out, backend.not_(_get(scope, inp)))

        elif k == "Instance":                       # :420-427
            if not instances:
                raise EvalError("Not enough instance to consume")
            _set(scope, gate[1], backend.instance(instances.popleft()))

        elif k == "Witness":                        # :429-432
            val = witnesses.popleft() if witnesses else None
            _set(scope, gate[1], backend.witness(val))

        elif k == "Free":                           # :434-439
            _, first, last = gate
            last_value = first if last is None else last
            for cur in range(first, last_value + 1):
                _remove(scope, cur)

        elif k == "Call":                           # :441-471
            _, name, output_wires, input_wires = gate
            if name not in kf:
                raise EvalError("Unknown function")
            fn = kf[name]
            eo = _expand(output_wires)
            ei = _expand(input_wires)
            _check_arity(name, fn, eo, ei)
            self._ingest_subcircuit(fn.subcircuit, backend, eo, ei, scope, {}, instances, witnesses, weight)

        elif k == "AnonCall":                       # :473-491
            _, output_wires, input_wires, _ic, _wc, subcircuit = gate
            eo = _expand(output_wires)
            ei = _expand(input_wires)
            self._ingest_subcircuit(subcircuit, backend, eo, ei, scope, known_iterators,
                                    instances, witnesses, weight)

        elif k == "For":                            # :495-559
            _, it_name, start, end, _global_outputs, body = gate
            for i in range(start, end + 1):
                known_iterators[it_name] = i
                if body[0] == "IterExprCall":
                    _, name, outputs, inputs = body
                    if name not in kf:
                        raise EvalError("Unknown function")
                    fn = kf[name]
                    eo = evaluate_iterexpr_list(outputs, known_iterators)
                    ei = evaluate_iterexpr_list(inputs, known_iterators)
                    _check_arity(name, fn, eo, ei)
                    self._ingest_subcircuit(fn.subcircuit, backend, eo, ei, scope, {},
                                            instances, witnesses, weight)
                else:
                    _, outputs, inputs, _ic, _wc, subcircuit = body
                    eo = evaluate_iterexpr_list(outputs, known_iterators)
                    ei = evaluate_iterexpr_list(inputs, known_iterators)
                    self._ingest_subcircuit(subcircuit, backend, eo, ei, scope, known_iterators,
                                            instances, witnesses, weight)
            known_iterators.pop(it_name, None)

        elif k == "Switch":                         # :563-688
            _, condition, output_wires, cases, branches = gate
            max_i = 0
            max_w = 0
            for br in branches:                     # :565-581
                if br[0] == "AbstractGateCall":
                    if br[1] not in kf:
                        raise EvalError("Unknown function")
                    ic, wc = kf[br[1]].instance_nbr, kf[br[1]].witness_nbr
                else:
                    ic, wc = br[2], br[3]
                max_i = max(max_i, ic)
                max_w = max(max_w, wc)
            # :586-591 — the first `max` values go to the branches, the rest stay
            ni = min(len(instances), max_i)
            nw = min(len(witnesses), max_w)
            new_instances = [instances.popleft() for _ in range(ni)]
            new_witnesses = [witnesses.popleft() for _ in range(nw)]

            branches_scope = []
            eo = _expand(output_wires)              # :597
            weights = []
            for case, br in zip(cases, branches):   # :600-670
                bw = compute_weight(backend, case, _get(scope, condition), modulus, is_boolean)
                wbw = as_mul(backend, weight, bw, is_boolean) if weight is not None else bw
                branch_scope = {}
                if br[0] == "AbstractGateCall":
                    _, name, input_wires = br
                    if name not in kf:
                        raise EvalError(f"Unknown function: {name}")
                    fn = kf[name]
                    ei = _expand(input_wires)
                    _check_arity(name, fn, eo, ei)
                    for w in ei:
                        branch_scope[w] = backend.copy(_get(scope, w))
                    self._ingest_subcircuit(fn.subcircuit, backend, eo, ei, branch_scope, {},
                                            deque(new_instances), deque(new_witnesses), wbw)
                else:
                    _, input_wires, _ic, _wc, subcircuit = br
                    ei = _expand(input_wires)
                    for w in ei:
                        branch_scope[w] = backend.copy(_get(scope, w))
                    self._ingest_subcircuit(subcircuit, backend, eo, ei, branch_scope, known_iterators,
                                            deque(new_instances), deque(new_witnesses), wbw)
                weights.append(wbw)
                branches_scope.append(branch_scope)
            for ow in eo:                           # :673-687
                acc = backend.constant(backend.zero())
                for bs, bw in zip(branches_scope, weights):
                    ww = as_mul(backend, _get(bs, ow), bw, is_boolean)
                    acc = as_add(backend, acc, ww, is_boolean)
                _set(scope, ow, acc)
        else:
            raise EvalError(f"unknown gate {k}")

    # -- :698-746 ---------------------------------------------------------
    def _ingest_subcircuit(self, subcircuit, backend, output_list, input_list, scope,
                           known_iterators, instances, witnesses, weight):
        new_scope: Dict[int, object] = {}
        n_out = len(output_list)
        for idx, inp in enumerate(input_list):
            _set(new_scope, idx + n_out, backend.copy(_get(scope, inp)))
        for gate in subcircuit:
            self._ingest_gate(gate, backend, new_scope, known_iterators, instances, witnesses, weight)
        for idx, out in enumerate(output_list):
            _set(scope, out, backend.copy(_get(new_scope, idx)))


def _expand(wl):
    try:
        return expand_wirelist(wl)
    except ValueError as e:
        raise EvalError(str(e))


def _check_arity(name, fn, eo, ei):                 # :449-454 etc.
    if len(eo) != fn.output_count:
        raise EvalError(f"Wrong number of output variables in call to function {name} "
                        f"(Expected {fn.output_count} / Got {len(eo)}).")
    if len(ei) != fn.input_count:
        raise EvalError(f"Wrong number of input variables in call to function {name} "
                        f"(Expected {fn.input_count} / Got {len(ei)}).")


def evaluate(messages, backend=None):
    """`main_evaluate` (rust/src/cli.rs:315-320): returns the violation list."""
    backend = backend or PlaintextBackend()
    return Evaluator.from_messages(messages, backend).get_violations()
