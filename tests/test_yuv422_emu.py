"""The 4:2:2 GPU algorithm on the CPU: yuv422_pipeline.cuh + the jump-ahead planner, driven exactly as
the kernel drives them (tests/emu422_harness.cpp), must equal the oracle BIT FOR BIT -- the path is
evaluated in double in the reference's operation order, so there is no tolerance."""
import ctypes as C

import numpy as np
import pytest

import helpers


@pytest.fixture(scope="module")
def orc():
    return helpers.load_oracle422()


@pytest.fixture(scope="module")
def emu():
    return helpers.load_emu422()


CASES = [
    (720, 480, 2, []),
    (720, 480, 3, ["-vhs"]),
    (720, 480, 2, ["-vhs", "-vhs-speed", "lp"]),
    (720, 480, 2, ["-vhs", "-vhs-speed", "ep", "-out-composite-lowpass", "0"]),
    (720, 480, 2, ["-vhs", "-in-composite-lowpass", "0", "-out-composite-lowpass", "0", "-out-composite-lowpass-lite", "0"]),
    (720, 480, 2, ["-comp-catv3", "-chroma-noise", "5"]),
    (720, 480, 2, ["-vhs", "-comp-catv", "-subcarrier-amp", "40"]),
    (720, 480, 2, ["-vhs", "-nocolor-subcarrier"]),
    (720, 480, 2, ["-nocolor-subcarrier-after-yc-sep"]),
    (720, 480, 2, ["-vhs", "-vhs-svideo", "1", "-vhs-chroma-vblend", "0"]),
    (724, 480, 3, ["-vhs", "-comp-phase", "270", "-comp-phase-offset", "1"]),
    (724, 480, 3, ["-comp-phase", "90", "-yc-recomb", "2"]),
    (720, 480, 2, ["-comp-phase", "0", "-comp-phase-offset", "3", "-noise", "0"]),
    (724, 576, 2, ["-tvstd", "pal", "-vhs"]),
    (720, 480, 2, ["-vhs", "-vhs-head-switching-point", "0.9501", "-vhs-head-switching-noise-level", "0.0001"]),
    (720, 480, 2, ["-vhs", "-chroma-dropout", "30000"]),
    (720, 480, 2, ["-vhs", "-noise", "300", "-chroma-noise", "200", "-chroma-phase-noise", "40"]),
    (34, 21, 3, ["-vhs", "-vhs-speed", "ep"]),
    (102, 67, 2, ["-vhs", "-vhs-speed", "lp"]),
    (16, 6, 2, ["-vhs"]),
    (8, 3, 2, []),
    (2, 2, 2, ["-vhs"]),
    (64, 32, 2, ["-vhs"]),
    (24, 10, 3, ["-vhs", "-vhs-speed", "ep"]),
    (16, 40, 2, ["-vhs", "-vhs-speed", "lp", "-chroma-dropout", "2000"]),
    (720, 480, 2, ["-vhs", "-vhs-speed", "ep"]),
    (1920, 1080, 1, ["-vhs", "-vhs-speed", "sp"]),
]


@pytest.mark.parametrize("w,h,n,argv", CASES)
@pytest.mark.parametrize("general", [0, 1, 2, 3])
def test_emulated_kernel_equals_oracle(orc, emu, w, h, n, argv, general):
    if general in (1, 3) and w >= 1920:
        pytest.skip("covered by the kernel's own choice")
    p = helpers.params422(*argv)
    want, g = helpers.run_oracle422(orc, p, w, h, n)
    got, pos = helpers.run_emu422(emu, p, w, h, n, force_general=general)
    assert pos == g.pos
    for k in range(n):
        for pl in range(3):
            d = np.argwhere(want[k][pl] != got[k][pl])
            assert d.size == 0, (k, pl, d[:5], want[k][pl][tuple(d[0])], got[k][pl][tuple(d[0])])


def test_emulated_kernel_tight_rows(orc, emu):
    w, h, n = 64, 32, 2
    p = helpers.params422("-vhs")
    want, _ = helpers.run_oracle422(orc, p, w, h, n, pad=0)
    got, _ = helpers.run_emu422(emu, p, w, h, n, pad=0)
    for k in range(n):
        for pl in range(3):
            assert np.array_equal(want[k][pl], got[k][pl])


def test_emulated_kernel_from_any_stream_position(orc, emu):
    """Sharding: a field computed from a seek position equals the field computed in sequence."""
    w, h = 102, 67
    p = helpers.params422("-vhs", "-vhs-speed", "lp")
    want, g = helpers.run_oracle422(orc, p, w, h, 4)
    pos = sum(orc.oracle422_draws_per_field(C.byref(p), w, h, (k & 1) ^ 1) for k in range(3))
    got, _ = helpers.run_emu422(emu, p, w, h, 1, first=3, rng_pos=pos)
    for pl in range(3):
        assert np.array_equal(want[3][pl], got[0][pl])
