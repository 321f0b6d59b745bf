"""Owned data model of SIEVE IR messages (oracle side, test infrastructure).

Mirrors the reference's owned structs:
  Gate            rust/src/structs/gates.rs:18-55
  WireListElement rust/src/structs/wire.rs:11-14
  IterExpr*       rust/src/structs/iterators.rs:17-30, 262-269 (list elements)
  Function        rust/src/structs/function.rs:19-26
  CaseInvoke      rust/src/structs/function.rs:121-126
  ForLoopBody     rust/src/structs/function.rs:269-274
  Header          rust/src/structs/header.rs:12-16
  Relation        rust/src/structs/relation.rs:35-41
  Instance        rust/src/structs/instance.rs:13-16
  Witness         rust/src/structs/witness.rs:13-16

Gates are plain tuples whose first element is the variant name, in the same
argument order as the Rust enum, e.g. ("Add", out, left, right).  Values are
`bytes`, little-endian, any length (rust/src/structs/value.rs:11).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

IR_VERSION = "1.0.0"  # rust/src/structs/mod.rs:36

# gate-set / feature masks, rust/src/structs/relation.rs:15-32
ADD, ADDC, MUL, MULC = 0x0001, 0x0002, 0x0004, 0x0008
ARITH = ADD | ADDC | MUL | MULC
XOR, AND, NOT = 0x0100, 0x0200, 0x0400
BOOL = XOR | AND | NOT
FUNCTION, FOR, SWITCH = 0x1000, 0x2000, 0x4000
FOR_FUNCTION_SWITCH = FOR | FUNCTION | SWITCH
SIMPLE = 0x0000


def contains_feature(feature_set: int, feature: int) -> bool:
    """rust/src/structs/relation.rs:284-286"""
    return (feature_set & feature) == feature


def parse_gate_set(s: str) -> int:
    """rust/src/structs/relation.rs:144-167"""
    ret = 0
    table = {"@add": ADD, "@addc": ADDC, "@mul": MUL, "@mulc": MULC,
             "@xor": XOR, "@not": NOT, "@and": AND}
    for sub in s.split(","):
        sub = sub.replace(" ", "")
        if sub == "arithmetic":
            return ARITH
        if sub == "boolean":
            return BOOL
        if sub == "":
            continue
        if sub not in table:
            raise ValueError(f"Unable to parse the following gateset: {s}")
        ret |= table[sub]
    return ret


def parse_feature_toggle(s: str) -> int:
    """rust/src/structs/relation.rs:229-244"""
    ret = 0
    table = {"@function": FUNCTION, "@for": FOR, "@switch": SWITCH}
    for sub in s.split(","):
        sub = sub.replace(" ", "")
        if sub == "simple":
            return SIMPLE
        if sub == "":
            continue
        if sub not in table:
            raise ValueError(f"Unable to parse following feature toggles {sub}")
        ret |= table[sub]
    return ret


def create_gateset_string(mask: int) -> str:
    """rust/src/structs/relation.rs:181-225"""
    ret = ""
    val = mask
    while val != 0:
        if contains_feature(val, ARITH):
            return "arithmetic"
        if contains_feature(val, BOOL):
            return "boolean"
        for m, name in ((ADD, "@add,"), (ADDC, "@addc,"), (MUL, "@mul,"), (MULC, "@mulc,"),
                        (XOR, "@xor,"), (NOT, "@not,"), (AND, "@and,")):
            if contains_feature(val, m):
                ret += name
                val ^= m
                break
        else:
            break
    return ret


def create_feature_string(mask: int) -> str:
    """rust/src/structs/relation.rs:257-280"""
    if (mask & FOR_FUNCTION_SWITCH) == 0:
        return "simple"
    ret = ""
    val = mask & FOR_FUNCTION_SWITCH
    for m, name in ((FOR, "@for,"), (SWITCH, "@switch,"), (FUNCTION, "@function,")):
        if contains_feature(val, m):
            ret += name
    return ret


# ---- wire lists ---------------------------------------------------------
def Wire(i):
    return ("Wire", i)


def WireRange(first, last):
    return ("WireRange", first, last)


def wirelist(*ids):
    """`wirelist![a, b, c]`, rust/src/lib.rs:74-81"""
    return [Wire(i) for i in ids]


def wirelist_rep(elem, n):
    """`wirelist![elem; n]`, rust/src/lib.rs:75-77"""
    return [Wire(elem) for _ in range(n)]


def expand_wirelist(wl) -> List[int]:
    """rust/src/structs/wire.rs:179-203 — a range needs last > first STRICTLY."""
    out: List[int] = []
    for el in wl:
        if el[0] == "Wire":
            out.append(el[1])
        else:
            first, last = el[1], el[2]
            if last <= first:
                raise ValueError(
                    f"In WireRange, last WireId ({last}) must be strictly greater than first WireId ({first}).")
            out.extend(range(first, last + 1))
    return out


def wirelist_len(wl) -> int:
    """rust/src/structs/wire.rs:221-229"""
    return sum(1 if el[0] == "Wire" else el[2] - el[1] + 1 for el in wl)


# ---- iterator expressions ----------------------------------------------
class OraclePanic(Exception):
    """Conditions on which the Rust reference panics (process abort)."""


U64 = (1 << 64) - 1


def evaluate_iterexpr(e, known) -> int:
    """rust/src/structs/iterators.rs:349-371 — u64 arithmetic (release build: wrapping)."""
    k = e[0]
    if k == "Const":
        return e[1]
    if k == "Name":
        if e[1] not in known:
            raise ValueError(f"Unknown iterator name {e[1]}")
        return known[e[1]]
    if k == "Add":
        return (evaluate_iterexpr(e[1], known) + evaluate_iterexpr(e[2], known)) & U64
    if k == "Sub":
        return (evaluate_iterexpr(e[1], known) - evaluate_iterexpr(e[2], known)) & U64
    if k == "Mul":
        return (evaluate_iterexpr(e[1], known) * evaluate_iterexpr(e[2], known)) & U64
    if k == "DivConst":
        if e[2] == 0:
            raise OraclePanic("attempt to divide by zero")
        return evaluate_iterexpr(e[1], known) // e[2]
    raise ValueError("Unknown Iterator Expression type")


def evaluate_iterexpr_list(lst, known) -> List[int]:
    """rust/src/structs/iterators.rs:374-403 — errors become PANICS here."""
    out: List[int] = []
    for el in lst:
        try:
            if el[0] == "Single":
                out.append(evaluate_iterexpr(el[1], known))
            else:
                a = evaluate_iterexpr(el[1], known)
                b = evaluate_iterexpr(el[2], known)
                out.extend(range(a, b + 1))
        except ValueError as exc:
            raise OraclePanic(str(exc))
    return out


# ---- messages ------------------------------------------------------------
@dataclass
class Header:
    field_characteristic: bytes = b""
    version: str = IR_VERSION
    field_degree: int = 1


@dataclass
class Function:
    name: str
    output_count: int
    input_count: int
    instance_count: int
    witness_count: int
    body: list


@dataclass
class Relation:
    header: Header
    gate_mask: int
    feat_mask: int
    functions: List[Function] = field(default_factory=list)
    gates: list = field(default_factory=list)


@dataclass
class Instance:
    header: Header
    common_inputs: List[bytes] = field(default_factory=list)


@dataclass
class Witness:
    header: Header
    short_witness: List[bytes] = field(default_factory=list)


def le_bytes(v: int, n: Optional[int] = None) -> bytes:
    if n is None:
        n = max(1, (v.bit_length() + 7) // 8)
    return int(v).to_bytes(n, "little")


def literal32(v: int) -> bytes:
    """rust/src/producers/examples.rs:216-224"""
    return int(v).to_bytes(4, "little")
