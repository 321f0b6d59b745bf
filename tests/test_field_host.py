"""The kernels' field arithmetic (csrc/field.cuh, __host__ __device__) compiled for the host and checked
against Python integers, for every limb count, including full-width moduli and inputs >= p on the
unreduced operand of a Montgomery product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.util import FIELDS, ROOT

SRC = os.path.join(ROOT, "tests", "native", "field_host.cpp")
LIB = os.path.join(ROOT, "tests", "native", "libfieldhost.so")


@pytest.fixture(scope="module")
def lib():
    deps = [SRC, os.path.join(ROOT, "zkinterface-ir_b200", "csrc", "field.cuh"),
            os.path.join(ROOT, "zkinterface-ir_b200", "csrc", "program.cpp")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I/usr/local/cuda/include", "-o", LIB, SRC,
                               os.path.join(ROOT, "zkinterface-ir_b200", "csrc", "program.cpp")])
    return C.CDLL(LIB)


def limbs(vals, n):
    out = np.zeros((len(vals), n), dtype=np.uint32)
    for i, v in enumerate(vals):
        for k in range(n):
            out[i, k] = (v >> (32 * k)) & 0xFFFFFFFF
    return out


def unlimbs(arr):
    return [sum(int(arr[i, k]) << (32 * k) for k in range(arr.shape[1])) for i in range(arr.shape[0])]


@pytest.mark.parametrize("name", list(FIELDS))
def test_field_ops_match_python(lib, name):
    p = FIELDS[name]
    fp = (C.c_uint32 * 28)()
    mod = p.to_bytes((p.bit_length() + 7) // 8, "little")
    n = lib.field_host_params(mod, len(mod), fp)
    assert n in (1, 2, 4, 8)
    R = 1 << (32 * n)
    assert fp[3 * 8] == (-pow(p, -1, 1 << 32)) % (1 << 32)            # n0inv
    assert unlimbs(np.array([list(fp[8:16])], dtype=np.uint32)[:, :n])[0] == R * R % p   # r2
    assert unlimbs(np.array([list(fp[16:24])], dtype=np.uint32)[:, :n])[0] == R % p      # one
    rng = np.random.default_rng(7)
    cnt = 2000
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2]
    a = edge + [int.from_bytes(rng.bytes(4 * n), "little") % p for _ in range(cnt)]
    b = list(reversed(edge)) + [int.from_bytes(rng.bytes(4 * n), "little") % p for _ in range(cnt)]
    A, B = limbs(a, n), limbs(b, n)
    out = np.zeros_like(A)
    call = lambda op, X, Y: (lib.field_host_batch(op, fp, X.ctypes.data_as(C.c_void_p), Y.ctypes.data_as(C.c_void_p),
                                                  out.ctypes.data_as(C.c_void_p), C.c_size_t(len(X))), unlimbs(out))[1]
    assert call(0, A, B) == [(x + y) % p for x, y in zip(a, b)]
    Rinv = pow(R, -1, p)
    assert call(1, A, B) == [x * y * Rinv % p for x, y in zip(a, b)]
    assert call(2, A, B) == [x * y % p for x, y in zip(a, b)]
    assert call(3, A, B) == [(x & y) % p for x, y in zip(a, b)]
    assert call(4, A, B) == [(x ^ y) % p for x, y in zip(a, b)]
    # unreduced first operand (any value < R) times a reduced one: what k_load_inputs relies on
    u = [R - 1, p, p + 1] + [int.from_bytes(rng.bytes(4 * n), "little") for _ in range(500)]
    r2 = [R * R % p] * len(u)
    U, R2 = limbs(u, n), limbs(r2, n)
    out = np.zeros_like(U)
    assert call(1, U, R2) == [x * R % p for x in u]
