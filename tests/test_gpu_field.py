"""Device field arithmetic vs Python integers: the PTX carry-chain Montgomery product (field_ptx.cuh) and the
portable CIOS product must agree bit for bit with each other and with the integers, for every limb count."""
import numpy as np
import pytest

from tests.test_field_host import limbs, unlimbs
from tests.util import FIELDS, zkb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(FIELDS))
def test_device_field_ops(name):
    z = zkb()
    p = FIELDS[name]
    b = z.GpuBackend(0)
    b.set_field(p)
    n = b.stats()["nlimb"]
    R = 1 << (32 * n)
    rng = np.random.default_rng(5)
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, R % p, (R * R) % p]
    xs = edge + [int.from_bytes(rng.bytes(4 * n), "little") % p for _ in range(20000)]
    ys = list(reversed(edge)) + [int.from_bytes(rng.bytes(4 * n), "little") % p for _ in range(20000)]
    A, B = limbs(xs, n), limbs(ys, n)
    Rinv = pow(R, -1, p)
    assert unlimbs(b.debug_field_ops(0, A, B)) == [(x + y) % p for x, y in zip(xs, ys)]
    want = [x * y * Rinv % p for x, y in zip(xs, ys)]
    assert unlimbs(b.debug_field_ops(1, A, B)) == want       # what the kernels run (PTX chains for N = 4, 8)
    assert unlimbs(b.debug_field_ops(2, A, B)) == want       # portable CIOS
    # an unreduced first operand (any N-limb integer) times a reduced one, as k_load_inputs does with R^2
    us = [R - 1, p, p + 1] + [int.from_bytes(rng.bytes(4 * n), "little") for _ in range(5000)]
    r2 = [(R * R) % p] * len(us)
    assert unlimbs(b.debug_field_ops(1, limbs(us, n), limbs(r2, n))) == [u * R % p for u in us]


@pytest.mark.parametrize("name", [k for k in FIELDS if FIELDS[k] > (1 << 64)])
def test_lazy_reduction_primitives(name):
    """the R1CS check's accumulator (field_ptx.cuh: fe_lazy_mad / fe_lazy_add_one / fe_lazy_finish): plain products summed
    over 2N + 1 limbs, one Montgomery reduction at the end — against Python integers, including moduli with the top bit set
    (the sum overflows 2N limbs) and the edge values"""
    z = zkb()
    p = FIELDS[name]
    b = z.GpuBackend(0)
    b.set_field(p)
    n = b.stats()["nlimb"]
    R = 1 << (32 * n)
    rng = np.random.default_rng(6)
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, R % p, (R * R) % p]
    xs = edge + [p - 1] * 9 + [int.from_bytes(rng.bytes(4 * n), "little") % p for _ in range(20000)]
    ys = list(reversed(edge)) + [p - 1 - k for k in range(9)] + [int.from_bytes(rng.bytes(4 * n), "little") % p for _ in range(20000)]
    A, B = limbs(xs, n), limbs(ys, n)
    Rinv = pow(R, -1, p)
    assert unlimbs(b.debug_field_ops(3, A, B)) == [(3 * x * y * Rinv + y) % p for x, y in zip(xs, ys)]
    assert unlimbs(b.debug_field_ops(4, A, B)) == [(x + y) % p for x, y in zip(xs, ys)]


def test_field_throughput_reports_sane_numbers():
    z = zkb()
    b = z.GpuBackend(0)
    b.set_field(FIELDS["bls381"])
    mul = b.debug_field_throughput(1, 500)
    add = b.debug_field_throughput(0, 500)
    assert mul > 1e10 and add > mul


def test_barrier_cost_microbenchmark():
    """the three barriers the all-levels kernel can use: cooperative_groups' grid.sync, the counter barrier, one cluster"""
    z = zkb()
    b = z.GpuBackend(0)
    for kind in (0, 1, 2):
        us = b.debug_barrier_cost(kind, 0, 50)
        assert 0.0 < us < 200.0, (kind, us)
    # the counter keeps counting across launches: a second measurement starts where the first ended
    assert 0.0 < b.debug_barrier_cost(1, 296, 20) < 200.0
