"""GPU test of the NVLink mailbox all-reduce ABI (csrc/om_mailbox.cu) in a single process: a world of one rank is the
identity, repeated rounds alternate the parity slots, sizes up to 128 values.  The multi-rank check (bit-identical sums on
every rank, equal to NCCL, late and timed-out ranks) is tests/test_gpu_mailbox_multi.py (torchrun / two GPUs); bench.py
checks the mailbox against NCCL bit for bit at N > 1 (`collective_check`)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_mailbox_world_of_one_is_identity_over_many_rounds():
    import torch
    from olympics_mujoco_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    handle = (C.c_ubyte * 64)()
    _lib.check(lib.om_mailbox_create(1, 0, C.byref(h), C.cast(handle, C.c_void_p)))
    try:
        _lib.check(lib.om_mailbox_connect(h, C.cast(handle, C.c_void_p)))
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        g = torch.Generator(device="cuda").manual_seed(0)
        for it in range(9):
            n = [68, 1, 128, 3, 65][it % 5]
            x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
            y = torch.empty_like(x)
            _lib.check(lib.om_mailbox_allreduce(h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), n, st))
            z = x.clone()
            _lib.check(lib.om_mailbox_allreduce(h, C.c_void_p(z.data_ptr()), C.c_void_p(z.data_ptr()), n, st))   # in place
            torch.cuda.synchronize()
            assert torch.equal(x, y) and torch.equal(x, z)
        flag = C.c_int(-1)
        _lib.check(lib.om_mailbox_timed_out(h, C.byref(flag)))
        assert flag.value == 0
        assert lib.om_mailbox_allreduce(h, None, None, 129, st) != 0          # more than 128 values: refused
        _lib.check(lib.om_mailbox_set_timeout_ms(h, 0.0))                     # 0 = wait for ever
        _lib.check(lib.om_mailbox_allreduce(h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), n, st))
        torch.cuda.synchronize()
        assert torch.equal(x, y)
        assert lib.om_mailbox_set_timeout_ms(h, -1.0) != 0
    finally:
        lib.om_mailbox_destroy(h)
