"""include/zkb.h section 7 without a device: the communicator entry points exist, validate their arguments and refuse
host-only contexts (there is no CPU fallback for evaluation, sharded or not)."""
import ctypes as C

import pytest

from tests.util import zkb


def test_comm_needs_devices():
    z = zkb()
    a, b = z.GpuBackend(-1), z.GpuBackend(-1)
    arr = (C.c_void_p * 2)(a._c, b._c)
    with pytest.raises(z.ZkbError) as e:
        a._chk(z._lib.zkb_comm_init(arr, 2))
    assert e.value.code == z.ZKB_E_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(z.ZkbError) as e:
        a.comm_init_rank(bytes(128), 1, 0)
    assert e.value.code == z.ZKB_E_CUDA
    with pytest.raises(z.ZkbError) as e:
        a.comm_init_rank(bytes(128), 2, 2)          # rank out of range
    assert e.value.code == z.ZKB_E_ARG
    for call in (lambda: a.comm_run(0, 1), lambda: a.comm_broadcast_program(0), lambda: a.comm_info()):
        with pytest.raises(z.ZkbError) as e:
            call()
        assert e.value.code == z.ZKB_E_ARG and "communicator" in str(e.value)
    assert z._lib.zkb_comm_init(None, 0) == z.ZKB_E_ARG
    a.close()
    b.close()
