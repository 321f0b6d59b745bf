"""ctypes wrapper of oracle/plaintext_flat.c (test infrastructure / CPU baseline only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libflatoracle.so")

GATE_DTYPE = np.dtype([("op", "u1"), ("pad", "u1", (3,)), ("out", "<u4"), ("a", "<u4"), ("b", "<u4")])
RESULT_DTYPE = np.dtype([("status", "<i4"), ("pad", "<u4"), ("fail_assert_seq", "<u8"), ("fail_wire", "<u8"),
                         ("gates_done", "<u8")])
EV_TRUE, EV_ASSERT_FAILED, EV_NO_VALUE, EV_ALREADY_SET, EV_NOT_ENOUGH_INSTANCE, EV_MISSING_WITNESS_PANIC, EV_BAD_GATE = range(7)


def build():
    src = os.path.join(_HERE, "plaintext_flat.c")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libflatoracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def violation_text(res) -> list:
    """the reference's violation strings for a flat_result (evaluator.rs:199-208, 357-362, 424, 775-797)"""
    s = int(res["status"])
    if s == EV_TRUE:
        return []
    if s == EV_ASSERT_FAILED:
        return [f"Wire_{int(res['fail_wire'])} (may be weighted) should be 0, while it is not"]
    if s == EV_NO_VALUE:
        return [f"No value given for wire_{int(res['fail_wire'])}"]
    if s == EV_ALREADY_SET:
        return [f"Wire_{int(res['fail_wire'])} already has a value in this scope."]
    if s == EV_NOT_ENOUGH_INSTANCE:
        return ["Not enough instance to consume"]
    raise RuntimeError(f"reference would panic / bad gate (status {s})")


def eval_batch(gates, const_pool, modulus_le: bytes, instances, witnesses, n_batch, n_threads=1):
    """instances / witnesses: uint8 [n_batch, n_vals, stride] (or [n_vals, stride] shared, instances only)."""
    gates = np.ascontiguousarray(gates, dtype=GATE_DTYPE)
    const_pool = np.ascontiguousarray(const_pool if const_pool is not None else np.zeros((0, 1), np.uint8), dtype=np.uint8)
    mod = np.frombuffer(modulus_le, dtype=np.uint8)
    dummy = np.zeros((1, 1), np.uint8)

    def prep(x):
        if x is None:
            return dummy, 0, 0, 1
        x = np.ascontiguousarray(x, dtype=np.uint8)
        if x.ndim == 2:
            return x, 0, x.shape[0], x.shape[1]
        return x, x.shape[1] * x.shape[2], x.shape[1], x.shape[2]

    inst, iss, n_inst, istr = prep(instances)
    wit, wss, n_wit, wstr = prep(witnesses)
    vstride = wstr if witnesses is not None else istr
    res = np.zeros(n_batch, dtype=RESULT_DTYPE)
    lib().flat_eval_batch(_p(gates), C.c_uint64(len(gates)), _p(const_pool), C.c_size_t(const_pool.shape[1]), _p(mod),
                          C.c_size_t(len(mod)), _p(inst), C.c_uint64(iss), C.c_uint64(n_inst), _p(wit), C.c_uint64(wss),
                          C.c_uint64(n_wit), C.c_size_t(vstride), C.c_uint32(n_batch), C.c_int(n_threads), _p(res))
    return res


def eval_dump(gates, const_pool, modulus_le: bytes, instance, witness, n_wires, stride=32):
    """single statement; returns (result, values[int or None per wire id < n_wires])"""
    gates = np.ascontiguousarray(gates, dtype=GATE_DTYPE)
    const_pool = np.ascontiguousarray(const_pool if const_pool is not None else np.zeros((0, 1), np.uint8), dtype=np.uint8)
    mod = np.frombuffer(modulus_le, dtype=np.uint8)
    dummy = np.zeros((1, 1), np.uint8)
    inst = np.ascontiguousarray(instance, dtype=np.uint8) if instance is not None else dummy
    wit = np.ascontiguousarray(witness, dtype=np.uint8) if witness is not None else dummy
    vstride = wit.shape[1] if witness is not None else inst.shape[1]
    res = np.zeros(1, dtype=RESULT_DTYPE)
    dump = np.zeros((n_wires, stride), dtype=np.uint8)
    lib().flat_eval_dump(_p(gates), C.c_uint64(len(gates)), _p(const_pool), C.c_size_t(const_pool.shape[1]), _p(mod),
                         C.c_size_t(len(mod)), _p(inst), C.c_uint64(inst.shape[0] if instance is not None else 0), _p(wit),
                         C.c_uint64(wit.shape[0] if witness is not None else 0), C.c_size_t(vstride), _p(res), _p(dump),
                         C.c_size_t(stride), C.c_uint64(n_wires))
    return res[0], dump


def mulmod(a: int, b: int, m: int) -> int:
    def le(v):
        return np.frombuffer(int(v).to_bytes(max(1, (int(v).bit_length() + 7) // 8), "little"), dtype=np.uint8)
    out = np.zeros(((m.bit_length() + 31) // 32) * 4 + 4, dtype=np.uint8)
    A, B, M = le(a), le(b), le(m)
    lib().flat_bn_mulmod(_p(A), C.c_size_t(len(A)), _p(B), C.c_size_t(len(B)), _p(M), C.c_size_t(len(M)), _p(out),
                         C.c_size_t(len(out)))
    return int.from_bytes(out.tobytes(), "little")
