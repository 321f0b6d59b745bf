"""CPU restatement of `rust/src/consumers/validator.rs` (oracle, test infrastructure).

Every method cites the reference lines it follows; violation texts are the reference's, character for
character.  Pinned on the reference's own tests (validator.rs:831-960: `test_validator`,
`test_validator_as_verifier`, `test_validator_violations`, `test_validator_free_violations`;
flattening.rs:200-225 `test_validate_flattening`; cli.rs:602-624 validate verb on both examples) in
tests/test_validator.py.

Third-party pieces: `num_bigint_dig::prime::probably_prime(n, 10)` (crate num-bigint-dig, not vendored) is a
probabilistic primality test (Miller-Rabin rounds + Lucas); any correct primality test gives the same answer
except with negligible probability, so a Miller-Rabin over fixed bases stands in.  `regex` 1.x: the two patterns
are restated with Python's `re` (`\\d`, `\\w` are Unicode classes in both).
"""
from __future__ import annotations

import re
from typing import Dict, List, Set

from . import ir
from .ir import OraclePanic, contains_feature, evaluate_iterexpr_list, expand_wirelist

VERSION_REGEX = r"^\d+.\d+.\d+$"                                                   # validator.rs:23
NAMES_REGEX = r"^[a-zA-Z_][\w]*(?:(?:\.|:{2})[a-zA-Z_][\w]*)*$"                    # validator.rs:25

_VERSION_RE = re.compile(VERSION_REGEX[:-1] + r"\Z")      # Rust's `$` matches only at the very end of the text
_NAMES_RE = re.compile(NAMES_REGEX[:-1] + r"\Z")

_SMALL_PRIMES = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89]


def is_probably_prime(value: bytes) -> bool:
    """structs/value.rs:53-56"""
    n = int.from_bytes(value, "little")
    if n < 2:
        return False
    for p in _SMALL_PRIMES:
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in _SMALL_PRIMES:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def _rust_debug_bytes(v: bytes) -> str:
    """`{:?}` of a Vec<u8>"""
    return "[" + ", ".join(str(b) for b in v) + "]"


class Validator:
    """validator.rs:68-87 (state), :107-829"""

    def __init__(self, as_prover: bool = False):
        self.as_prover = as_prover
        self.instance_queue_len = 0
        self.witness_queue_len = 0
        self.live_wires: Set[int] = set()
        self.got_header = False
        self.gate_set = 0
        self.features = 0
        self.header_version = ""
        self.field_characteristic = 0
        self.field_degree = 0
        self.known_functions: Dict[str, tuple] = {}      # shared with sub-validators (Rc<RefCell<..>>)
        self.known_iterators: Dict[str, int] = {}
        self.violations: List[str] = []

    @classmethod
    def new_as_verifier(cls):                            # :108-110
        return cls(False)

    @classmethod
    def new_as_prover(cls):                              # :112-117
        return cls(True)

    def get_violations(self) -> List[str]:               # :136-144
        self.ensure_all_instance_values_consumed()
        self.ensure_all_witness_values_consumed()
        return self.violations

    def ingest_message(self, msg):                       # :154-160
        if isinstance(msg, ir.Instance):
            self.ingest_instance(msg)
        elif isinstance(msg, ir.Witness):
            self.ingest_witness(msg)
        else:
            self.ingest_relation(msg)

    def ingest_header(self, header):                     # :162-202
        fc = int.from_bytes(header.field_characteristic, "little")
        if self.got_header:
            if self.field_characteristic != fc:
                self.violate("The field_characteristic field is not consistent across headers.")
            if self.field_degree != header.field_degree:
                self.violate("The field_degree is not consistent across headers.")
            if self.header_version != header.version:
                self.violate("The profile version is not consistent across headers.")
        else:
            self.got_header = True
            self.field_characteristic = fc
            if not fc > 1:
                self.violate("The field_characteristic should be > 1")
            if not is_probably_prime(header.field_characteristic):
                self.violate("The field_characteristic should be a prime.")
            self.field_degree = header.field_degree
            if self.field_degree != 1:
                self.violate("field_degree must be = 1")
            if not _VERSION_RE.search(header.version.strip()):
                self.violate("The profile version should match the following format <major>.<minor>.<patch>.")
            self.header_version = header.version

    def ingest_instance(self, instance):                 # :204-213
        self.ingest_header(instance.header)
        for value in instance.common_inputs:
            self.ensure_value_in_field(value, lambda: f"instance value {_rust_debug_bytes(value)}")
        self.instance_queue_len += len(instance.common_inputs)

    def ingest_witness(self, witness):                   # :215-227
        if not self.as_prover:
            self.violate("As verifier, got an unexpected Witness message.")
        self.ingest_header(witness.header)
        for value in witness.short_witness:
            self.ensure_value_in_field(value, lambda: f"witness value {_rust_debug_bytes(value)}")
        self.witness_queue_len += len(witness.short_witness)

    def ingest_relation(self, relation):                 # :229-289
        self.ingest_header(relation.header)
        self.gate_set = relation.gate_mask
        if contains_feature(self.gate_set, ir.BOOL) and contains_feature(self.gate_set, ir.ARITH):
            self.violate("Cannot mix arithmetic and boolean gates")
        if contains_feature(self.gate_set, ir.BOOL):
            if self.field_characteristic != 2:
                self.violate("With boolean profile the field characteristic can only be 2.")
        self.features = relation.feat_mask
        for f in relation.functions:
            self.ensure_allowed_feature("@function", ir.FUNCTION)
            if not _NAMES_RE.search(f.name.strip()):
                self.violate(f"The function name ({f.name}) should match the proper format ({NAMES_REGEX}).")
            if f.name in self.known_functions:
                self.violate(f"A function with the name '{f.name}' already exists")
                continue
            self.known_functions[f.name] = (f.output_count, f.input_count, f.instance_count, f.witness_count)
            self.ingest_subcircuit(f.body, f.output_count, f.input_count, f.instance_count, f.witness_count, False)
        for gate in relation.gates:
            self.ingest_gate(gate)

    def _expand(self, wl) -> List[int]:
        try:
            return expand_wirelist(wl)
        except ValueError as err:
            self.violate(str(err))
            return []

    def _iterexprs(self, lst) -> List[int]:
        return evaluate_iterexpr_list(lst, self.known_iterators)   # errors are PANICS (iterators.rs:400)

    def ingest_gate(self, gate):                         # :291-642
        k = gate[0]
        if k == "Constant":
            self.ensure_value_in_field(gate[2], lambda: "Gate::Constant constant")
            self.ensure_undefined_and_set(gate[1])
        elif k == "AssertZero":
            self.ensure_defined_and_set(gate[1])
        elif k == "Copy":
            self.ensure_defined_and_set(gate[2])
            self.ensure_undefined_and_set(gate[1])
        elif k in ("Add", "Mul", "And", "Xor"):
            name, mask = {"Add": ("@add", ir.ADD), "Mul": ("@mul", ir.MUL), "And": ("@and", ir.AND),
                          "Xor": ("@xor", ir.XOR)}[k]
            self.ensure_allowed_gate(name, mask)
            self.ensure_defined_and_set(gate[2])
            self.ensure_defined_and_set(gate[3])
            self.ensure_undefined_and_set(gate[1])
        elif k in ("AddConstant", "MulConstant"):
            name, mask = ("@addc", ir.ADDC) if k == "AddConstant" else ("@mulc", ir.MULC)
            self.ensure_allowed_gate(name, mask)
            self.ensure_value_in_field(gate[3], lambda: f"Gate::{k}_{gate[1]}")
            self.ensure_defined_and_set(gate[2])
            self.ensure_undefined_and_set(gate[1])
        elif k == "Not":
            self.ensure_allowed_gate("@not", ir.NOT)
            self.ensure_defined_and_set(gate[2])
            self.ensure_undefined_and_set(gate[1])
        elif k == "Instance":
            self.declare(gate[1])
            self.consume_instance(1)
        elif k == "Witness":
            self.declare(gate[1])
            self.consume_witness(1)
        elif k == "Free":
            first, last = gate[1], gate[2]
            if last is not None and last <= first:
                self.violate(f"For Free gates, last WireId ({last}) must be strictly greater than first WireId ({first}).")
            for w in range(first, (first if last is None else last) + 1):
                self.ensure_defined_and_set(w)
                self.remove(w)
        elif k == "AnonCall":
            _, outs, ins, icount, wcount, sub = gate
            self.ensure_allowed_feature("@anoncall", ir.FUNCTION)
            eo = self._expand(outs)
            ei = self._expand(ins)
            for w in ei:
                self.ensure_defined_and_set(w)
            self.ingest_subcircuit(sub, len(eo), len(ei), icount, wcount, True)
            self.consume_instance(icount)
            self.consume_witness(wcount)
            for w in eo:
                self.ensure_undefined_and_set(w)
        elif k == "Call":
            _, name, outs, ins = gate
            self.ensure_allowed_feature("@call", ir.FUNCTION)
            eo = self._expand(outs)
            ei = self._expand(ins)
            for w in ei:
                self.ensure_defined_and_set(w)
            icount, wcount = self.ingest_call(name, eo, ei) or (0, 0)
            self.consume_instance(icount)
            self.consume_witness(wcount)
            for w in eo:
                self.ensure_undefined_and_set(w)
        elif k == "Switch":
            _, cond, outs, cases, branches = gate
            self.ensure_allowed_feature("@switch", ir.SWITCH)
            self.ensure_defined_and_set(cond)
            if len(cases) != len(branches):
                self.violate("Gate::Switch: The number of cases value does not match the number of branches.")
            if len(cases) == 0:
                if len(outs) != 0:
                    self.violate("Switch: no case given while non-empty list of output wires.")
                return
            cases_set = set()
            for case in cases:
                self.ensure_value_in_field(case, lambda: f"Gate::Switch case value: {int.from_bytes(case, 'little')}")
                cases_set.add(int.from_bytes(case, "little"))
            if len(cases_set) != len(cases):
                self.violate("Gate::Switch: The cases values contain duplicates.")
            max_i = max_w = 0
            eo = self._expand(outs)
            for br in branches:
                if br[0] == "AbstractGateCall":
                    ei = self._expand(br[2])
                    for w in ei:
                        self.ensure_defined_and_set(w)
                    icount, wcount = self.ingest_call(br[1], eo, ei) or (0, 0)
                else:
                    _, ins, icount, wcount, sub = br
                    ei = self._expand(ins)
                    for w in ei:
                        self.ensure_defined_and_set(w)
                    self.ingest_subcircuit(sub, len(eo), len(ei), icount, wcount, True)
                max_i = max(max_i, icount)
                max_w = max(max_w, wcount)
            self.consume_instance(max_i)
            self.consume_witness(max_w)
            for w in eo:
                self.ensure_undefined_and_set(w)
        elif k == "For":
            _, it_name, start, end, global_outs, body = gate
            self.ensure_allowed_feature("@for", ir.FOR)
            if end < start:
                self.violate(f"In a For loop, the end value ({end}) must be strictly greater than the start value ({start}).")
                return
            if it_name in self.known_iterators:
                self.violate("Iterator already used in this context.")
                return
            if not _NAMES_RE.search(it_name):
                self.violate(f"The iterator name ({it_name}) should match the following format ({NAMES_REGEX}).")
            for i in range(start, end + 1):
                self.known_iterators[it_name] = i
                if body[0] == "IterExprCall":
                    _, name, outs, ins = body
                    eo = self._iterexprs(outs)
                    ei = self._iterexprs(ins)
                    for w in ei:
                        self.ensure_defined_and_set(w)
                    icount, wcount = self.ingest_call(name, eo, ei) or (0, 0)
                    for w in eo:
                        self.ensure_undefined_and_set(w)
                    self.consume_instance(icount)
                    self.consume_witness(wcount)
                else:
                    _, outs, ins, icount, wcount, sub = body
                    eo = self._iterexprs(outs)
                    ei = self._iterexprs(ins)
                    for w in ei:
                        self.ensure_defined_and_set(w)
                    self.ingest_subcircuit(sub, len(eo), len(ei), icount, wcount, True)
                    for w in eo:
                        self.ensure_undefined_and_set(w)
                    self.consume_instance(icount)
                    self.consume_witness(wcount)
            self.known_iterators.pop(it_name, None)
            for w in self._expand(global_outs):
                self.ensure_defined_and_set(w)
        else:
            raise ValueError(f"unknown gate {k}")

    def ingest_call(self, name, output_wires, input_wires):   # :649-673 (None stands for Err)
        if name not in self.known_functions:
            self.violate(f"Unknown Function gate {name}")
            return None
        oc, ic, instance_count, witness_count = self.known_functions[name]
        if oc != len(output_wires):
            self.violate("Call: number of output wires mismatch.")
        if ic != len(input_wires):
            self.violate("Call: number of input wires mismatch.")
        return instance_count, witness_count

    def ingest_subcircuit(self, subcircuit, output_count, input_count, instance_count, witness_count,
                          use_same_scope):               # :684-738
        cur = Validator(self.as_prover)
        cur.instance_queue_len = instance_count
        cur.witness_queue_len = witness_count if self.as_prover else 0
        cur.got_header = self.got_header
        cur.gate_set = self.gate_set
        cur.features = self.features
        cur.header_version = self.header_version
        cur.field_characteristic = self.field_characteristic
        cur.field_degree = self.field_degree
        cur.known_functions = self.known_functions
        cur.known_iterators = self.known_iterators if use_same_scope else {}
        for wire in range(output_count, output_count + input_count):
            cur.live_wires.add(wire)
        for g in subcircuit:
            cur.ingest_gate(g)
        for w in range(output_count):
            cur.ensure_defined_and_set(w)
        self.violations.extend(cur.violations)
        if cur.instance_queue_len != 0:
            self.violate("The subcircuit has not consumed all the instance variables it should have.")
        if cur.witness_queue_len != 0:
            self.violate("The subcircuit has not consumed all the witness variables it should have.")

    # ---- helpers, :740-829 ---------------------------------------------------------------------------
    def is_defined(self, w):
        return w in self.live_wires

    def declare(self, w):
        self.live_wires.add(w)

    def remove(self, w):
        if w in self.live_wires:
            self.live_wires.remove(w)
        else:
            self.violate(f"The variable {w} is being freed, but was not defined previously, or has been already freed")

    def consume_instance(self, how_many):
        if self.instance_queue_len >= how_many:
            self.instance_queue_len -= how_many
        else:
            self.instance_queue_len = 0
            self.violate("Not enough Instance value to consume.")

    def consume_witness(self, how_many):
        if self.as_prover:
            if self.witness_queue_len >= how_many:
                self.witness_queue_len -= how_many
            else:
                self.witness_queue_len = 0
                self.violate("Not enough Witness value to consume.")

    def ensure_defined_and_set(self, w):
        if not self.is_defined(w):
            if self.as_prover:
                self.violate(f"The wire {w} is used but was not assigned a value, or has been freed already.")
            self.declare(w)

    def ensure_undefined(self, w):
        if self.is_defined(w):
            self.violate(f"The wire {w} has already been initialized before. This violates the SSA property.")

    def ensure_undefined_and_set(self, w):
        self.ensure_undefined(w)
        self.declare(w)

    def ensure_value_in_field(self, value: bytes, name):
        if len(value) == 0:
            self.violate(f"The {name()} is empty.")
        v = int.from_bytes(value, "little")
        if v >= self.field_characteristic:
            self.violate(f"The {name()} cannot be represented in the field specified in Header "
                         f"({v} >= {self.field_characteristic}).")

    def ensure_allowed_gate(self, gate_name, gate_mask):
        if not contains_feature(self.gate_set, gate_mask):
            self.violate(f"The gate {gate_name} is not allowed in this circuit.")

    def ensure_allowed_feature(self, gate_name, feature_mask):
        if not contains_feature(self.features, feature_mask):
            self.violate(f"The feature {gate_name} is not allowed in this circuit.")

    def ensure_all_instance_values_consumed(self):
        if self.instance_queue_len > 0:
            self.violate(f"Too many Instance values ({self.instance_queue_len} not consumed)")

    def ensure_all_witness_values_consumed(self):
        if self.as_prover and self.witness_queue_len > 0:
            self.violate(f"Too many Witness values ({self.witness_queue_len} not consumed)")

    def violate(self, msg):
        self.violations.append(msg)


def validate(messages, as_prover=True) -> List[str]:
    v = Validator(as_prover)
    for m in messages:
        v.ingest_message(m)
    return v.get_violations()
