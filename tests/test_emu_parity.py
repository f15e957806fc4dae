"""The product's lane pipeline (csrc/lane_pipeline.cuh) + host planning (field_plan.cpp, glibc_rand.cpp)
executed on the CPU, 32 lanes in lock-step exactly as the kernels run them (tests/emu_harness.cpp),
against the oracle.  fp64 = the reference's arithmetic: must be bit-exact, on all three code variants
(fast interior, edge, general).  fp32 = production arithmetic: within +-1 LSB."""
import numpy as np
import pytest

import helpers

CASES = [
    (720, 480, 2, []),
    (720, 480, 3, ["-vhs", "-vhs-speed", "sp"]),
    (720, 480, 2, ["-vhs", "-vhs-speed", "ep"]),
    (724, 482, 2, ["-vhs", "-vhs-speed", "lp"]),
    (720, 480, 2, ["-vhs", "-out-composite-lowpass-lite", "0"]),
    (720, 480, 2, ["-out-composite-lowpass-lite", "0", "-chroma-noise", "3", "-chroma-phase-noise", "2"]),
    (720, 480, 2, ["-vhs", "-in-composite-lowpass", "0", "-out-composite-lowpass", "0"]),
    (720, 480, 2, ["-vhs", "-comp-catv", "-subcarrier-amp", "40"]),
    (720, 480, 2, ["-vhs", "-nocolor-subcarrier"]),
    (720, 480, 2, ["-vhs", "-vhs-svideo", "1", "-vhs-chroma-vblend", "0"]),
    (720, 480, 3, ["-vhs", "-comp-phase", "270", "-comp-phase-offset", "1"]),
    (720, 480, 2, ["-vhs", "-comp-phase", "90"]),
    (720, 576, 2, ["-tvstd", "pal", "-vhs"]),
    (720, 480, 2, ["-vhs", "-vhs-head-switching-phase", "0.001", "-vhs-head-switching-point", "0.5"]),
    (720, 480, 2, ["-vhs", "-noise", "0", "-chroma-noise", "0", "-chroma-phase-noise", "0", "-chroma-dropout", "0"]),
    (720, 480, 2, ["-vhs", "-chroma-dropout", "30000"]),
    (33, 21, 3, ["-vhs", "-vhs-speed", "ep"]),
    (37, 9, 2, ["-vhs"]),
    (101, 67, 2, ["-vhs", "-vhs-speed", "lp", "-out-composite-lowpass-lite", "0"]),
    (48, 480, 2, ["-vhs"]),
    (32, 4, 2, []),
    (1920, 1080, 1, ["-vhs", "-vhs-speed", "sp"]),
]


@pytest.mark.parametrize("w,h,n,argv", CASES)
def test_lane_pipeline_matches_oracle(oracle, emu, w, h, n, argv):
    p = helpers.params(*argv)
    frames = lambda k: helpers.stream_frame(w, h, k)
    want, g = helpers.run_oracle(oracle, p, frames, n, w, h)
    for general in (0, 1, 2):       # 0 = the kernel's own choice, 1 = general variant everywhere, 2 = edge variant everywhere
        got, pos = helpers.run_emu(emu, p, frames, n, w, h, precision=1, general=general)
        assert np.array_equal(want, got), ("fp64", general, helpers.channel_diff(want, got))
        assert pos == g.pos
    got, _ = helpers.run_emu(emu, p, frames, n, w, h, precision=0)
    mx, nd, n2 = helpers.channel_diff(want, got)
    assert mx <= 1 and n2 == 0, (mx, nd, n2)
    assert nd <= max(8, 0.005 * want.size * 4)


def test_worst_case_content(oracle, emu):
    """Random-noise source image: every pixel is an edge (SURVEY App. C precision probe)."""
    w, h, n = 720, 480, 2
    p = helpers.params("-vhs", "-vhs-speed", "ep")
    frames = lambda k: helpers.noise_frame(w, h, 100 + k)
    want, _ = helpers.run_oracle(oracle, p, frames, n, w, h)
    got, _ = helpers.run_emu(emu, p, frames, n, w, h, precision=1)
    assert np.array_equal(want, got)
    got, _ = helpers.run_emu(emu, p, frames, n, w, h, precision=0)
    mx, nd, n2 = helpers.channel_diff(want, got)
    assert mx <= 1 and n2 == 0


def test_interlaced_source_rows(oracle, emu):
    w, h, n = 160, 121, 3
    p = helpers.params("-vhs")
    frames = lambda k: helpers.noise_frame(w, h, k)
    want, _ = helpers.run_oracle(oracle, p, frames, n, w, h, interlaced=1, tff=1)
    got, _ = helpers.run_emu(emu, p, frames, n, w, h, precision=1, interlaced=1, tff=1)
    assert np.array_equal(want, got)


@pytest.mark.parametrize("name", sorted(helpers.load_golden().keys()))
def test_lane_pipeline_matches_golden(emu, name):
    argv, w, h, n, want = helpers.load_golden()[name]
    got, _ = helpers.run_emu(emu, helpers.params(*argv), lambda k: helpers.stream_frame(w, h, k), n, w, h, precision=1)
    assert np.array_equal(want, got)
