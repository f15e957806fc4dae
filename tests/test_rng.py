"""The glibc rand() models (oracle's sequential one, the product's jump-ahead one) against this
container's libc, plus the exact multiply-shift modulo used by the kernels."""
import ctypes as C

import numpy as np
import pytest

import helpers

FIRST = [1804289383, 846930886, 1681692777, 1714636915, 1957747793, 424238335, 719885386, 1649760492]


def test_oracle_rng_is_glibc_rand(oracle):
    libc = C.CDLL(None)
    libc.srand(1)
    g = helpers.OracleRng()
    oracle.oracle_rng_seed(C.byref(g), 1)
    got = [oracle.oracle_rng_next(C.byref(g)) for _ in range(5000)]
    want = [libc.rand() for _ in range(5000)]
    assert got[:8] == FIRST
    assert got == want


def test_product_rng_sequential_and_jump(emu):
    emu.emu_rand_at.restype = C.c_uint
    emu.emu_rand_at.argtypes = [C.c_ulonglong]
    out = (C.c_uint * 8)()
    emu.emu_rand_run(C.c_ulonglong(0), 8, out)
    assert list(out) == FIRST
    assert emu.emu_rand_at(1000000) == 771126689
    # jump == sequential at positions around a 1080p VHS field boundary
    libc = C.CDLL(None)
    libc.srand(1)
    seq = np.array([libc.rand() for _ in range(300000)], dtype=np.uint32)
    for pos in (0, 1, 30, 31, 32, 4095, 4096, 4097, 123457, 299999):
        assert emu.emu_rand_at(pos) == int(seq[pos])
    run = (C.c_uint * 100)()
    emu.emu_rand_run(C.c_ulonglong(250000), 100, run)
    assert list(run) == [int(v) for v in seq[250000:250100]]


@pytest.mark.parametrize("m", [3, 5, 9, 11, 13, 33, 39, 45, 101, 201, 999, 65535])
def test_mod_magic_is_exact(emu, m):
    assert emu.emu_mod_magic_mismatches(m) == 0
