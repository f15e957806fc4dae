"""cvs422_params_apply_argv mirrors parse_argv() of ffmpeg_to_composite.cpp (:1292-1650) for the switches
the video path honours: order-dependent side effects, ignored audio/container switches, failures."""
import ctypes as C

import numpy as np
import pytest

import helpers
from composite_video_simulator_b200 import yuv422


def test_defaults_are_the_reference_initialisers():
    p = yuv422.params_default()
    assert p.output_ntsc == 1 and p.video_scanline_phase_shift == 180
    assert (p.composite_in_chroma_lowpass, p.composite_out_chroma_lowpass, p.composite_out_chroma_lowpass_lite) == (1, 1, 1)
    assert p.vhs_head_switching == 0 and p.emulating_vhs == 0 and p.enable_composite_emulation == 1
    assert p.vhs_head_switching_phase == 1.0 - ((4.5 + 0.01) / 262.5)
    assert p.vhs_head_switching_phase_noise == ((1.0 / 300) / 262.5)
    assert p.composite_preemphasis == 0 and p.composite_preemphasis_cut == 1000000


def test_vhs_sets_noise_levels_and_head_switching():
    p = helpers.params422("-vhs")
    assert (p.emulating_vhs, p.vhs_head_switching) == (1, 1)
    assert (p.video_chroma_phase_noise, p.video_chroma_noise, p.video_chroma_loss, p.video_noise) == (4, 16, 4, 4)
    p = helpers.params422("-noise", "9", "-vhs-speed", "ep")           # -vhs-speed overwrites, implies -vhs, not head switching
    assert (p.emulating_vhs, p.vhs_head_switching, p.output_vhs_tape_speed) == (1, 0, 2)
    assert (p.video_chroma_phase_noise, p.video_chroma_noise, p.video_chroma_loss, p.video_noise) == (6, 22, 8, 6)
    p = helpers.params422("-vhs-speed", "lp", "-noise", "9")           # ... and a later -noise wins
    assert p.video_noise == 9 and p.video_chroma_noise == 19


def test_catv_and_amplitude_back_adjustment():
    p = helpers.params422("-comp-catv2")
    assert p.composite_preemphasis == 2.5 and p.composite_preemphasis_cut == float(315000000 // 88 // 2)
    assert p.video_chroma_phase_noise == 4
    assert p.subcarrier_amplitude == 50 and p.subcarrier_amplitude_back == int(50 + (50 * 2.5) / 4)
    p = helpers.params422("-subcarrier-amp", "40", "-comp-pre", "-1.5")
    assert p.subcarrier_amplitude_back == int(40 + (50 * -1.5) / 4)   # int += double truncates the sum


def test_misc_switches():
    p = helpers.params422("-tvstd", "pal", "-yc-recomb", "2.9", "-vhs-head-switching-point", "0.25", "-nocomp",
                          "-vhs-svideo", "1", "-vhs-chroma-vblend", "0", "-comp-phase", "270", "-comp-phase-offset", "-1")
    assert (p.output_ntsc, p.output_height, p.video_yc_recombine) == (0, 576, 2)
    assert p.vhs_head_switching_phase == 0.25 and p.enable_composite_emulation == 0
    assert (p.vhs_svideo_out, p.vhs_chroma_vert_blend, p.video_scanline_phase_shift, p.video_scanline_phase_shift_offset) == (1, 0, 270, -1)
    # audio / container switches are accepted and ignored with their value
    q = helpers.params422("-i", "in.mkv", "-o", "out.mkv", "-audio-hiss", "-60", "-vn", "-422", "-ss", "3", "-vhs-hifi", "0")
    assert q.emulating_vhs == 1 and q.video_noise == 2


@pytest.mark.parametrize("argv,status", [
    (["-bogus"], -2), (["stray"], -2), (["-comp-phase", "45"], -2), (["-vhs-speed", "xp"], -2),
    (["-tvstd", "secam"], -2), (["-width", "16"], -2), (["-noise"], -2), (["-h"], -6),
])
def test_failures(argv, status):
    with pytest.raises(yuv422.Yuv422Error) as e:
        helpers.params422(*argv)
    assert e.value.status == status


def test_golden_fixtures_against_oracle_and_emulation():
    """The committed outputs of the reference's own code (tests/golden/make_golden_yuv422.py)."""
    import glob
    import os
    orc, emu = helpers.load_oracle422(), helpers.load_emu422()
    files = sorted(glob.glob(os.path.join(helpers.GOLDEN_DIR, "yuv422_*.npz")))
    assert len(files) >= 4
    for f in files:
        z = np.load(f)
        argv = [a for a in str(z["argv"]).split(" ") if a]
        w, h, n = int(z["w"]), int(z["h"]), int(z["n"])
        p = helpers.params422(*argv)
        a, _ = helpers.run_oracle422(orc, p, w, h, n)
        b, _ = helpers.run_emu422(emu, p, w, h, n)
        for k in range(n):
            for pl, nm in enumerate("YUV"):
                assert np.array_equal(a[k][pl], z["%s%d" % (nm, k)]), (f, k, nm)
                assert np.array_equal(b[k][pl], z["%s%d" % (nm, k)]), (f, k, nm)
