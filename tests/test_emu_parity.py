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


@pytest.mark.parametrize("w,h,n,argv", [(720, 480, 3, ["-vhs", "-vhs-speed", "sp"]), (724, 482, 2, ["-vhs", "-vhs-speed", "ep"]),
                                        (720, 480, 2, []), (101, 67, 2, ["-vhs", "-vhs-speed", "lp"]),
                                        (720, 480, 2, ["-vhs", "-chroma-dropout", "30000"])])
def test_fast_noise_mode_stays_within_one_lsb(oracle, emu, w, h, n, argv):
    """CVS_NOISE_FAST (include/cvs_ntsc.h): the per-pixel noise draws come from counter generators, everything
    else -- per-line phase noise, dropouts, head-switch jitter, the rand() position -- is as in exact mode, so
    the pictures stay within +-1 LSB of the reference and the stream position is unchanged."""
    p = helpers.params(*argv)
    frames = lambda k: helpers.stream_frame(w, h, k)
    want, g = helpers.run_oracle(oracle, p, frames, n, w, h)
    for general in (0, 1, 2):
        got, pos = helpers.run_emu(emu, p, frames, n, w, h, precision=0, general=general, noise_fast=1)
        mx, nd, n2 = helpers.channel_diff(want, got)
        assert mx <= 1 and n2 == 0, (general, mx, nd, n2)
        assert nd <= 0.15 * want.size * 4, nd            # more values move by 1 than in exact mode (< 0.5 %)
        assert pos == g.pos
        if general == 0:
            ref = got
        else:
            assert np.array_equal(ref, got)              # all code variants draw the same fast stream
    exact, _ = helpers.run_emu(emu, p, frames, n, w, h, precision=0)
    assert not np.array_equal(exact, ref)                # (it really is a different noise)


def test_fast_noise_mode_is_ignored_for_large_amplitudes_and_fp64(oracle, emu):
    w, h, n = 160, 120, 2
    frames = lambda k: helpers.stream_frame(w, h, k)
    p = helpers.params("-vhs", "-noise", "100")           # 100 + 2 * 16 > 96: no longer sub-LSB
    a, _ = helpers.run_emu(emu, p, frames, n, w, h, precision=0, noise_fast=1)
    b, _ = helpers.run_emu(emu, p, frames, n, w, h, precision=0, noise_fast=0)
    assert np.array_equal(a, b)
    p = helpers.params("-vhs")
    a, _ = helpers.run_emu(emu, p, frames, n, w, h, precision=1, noise_fast=1)
    want, _ = helpers.run_oracle(oracle, p, frames, n, w, h)
    assert np.array_equal(a, want)


@pytest.mark.parametrize("warm_px", [0, 3])
@pytest.mark.parametrize("w,h,n,argv", [(160, 120, 3, ["-vhs"]), (101, 67, 2, ["-vhs", "-vhs-speed", "ep"]), (64, 300, 2, [])])
def test_short_warmup_takes_the_second_chance_and_stays_exact(oracle, emu, warm_px, w, h, n, argv):
    """The noise state of a row is recovered by running the recurrence from both ends of its range over the draws
    before the row; when the two runs do not merge (p < 2^-58 per row at the production length of 64 pixels) the lane
    walks the generator BACKWARD (rewarm_bracket, lane_pipeline.cuh) and tries again.  A warm-up of 0 or 3 pixels
    makes the first attempt fail on nearly every row: the pictures must still be the reference's, bit for bit."""
    p = helpers.params(*argv)
    frames = lambda k: helpers.noise_frame(w, h, 7 + k)
    want, g = helpers.run_oracle(oracle, p, frames, n, w, h)
    emu.emu_set_warm_px(warm_px)
    try:
        got, pos = helpers.run_emu(emu, p, frames, n, w, h, precision=1)
    finally:
        emu.emu_set_warm_px(64)
    assert pos == g.pos
    assert np.array_equal(want, got)
