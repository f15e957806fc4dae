"""cvs_params / cvs_params_apply_argv against the semantics of parse_argv() (ffmpeg_ntsc.cpp:972-1282)."""
import pytest

import composite_video_simulator_b200 as cvs
from composite_video_simulator_b200 import CvsError


def P(*argv):
    return cvs.params_from_argv(list(argv))


def test_defaults():
    p = cvs.default_params()
    assert (p.output_width, p.output_height, p.output_ntsc) == (720, 480, 1)
    assert (p.video_noise, p.video_chroma_noise, p.video_chroma_phase_noise, p.video_chroma_loss) == (2, 0, 0, 0)
    assert (p.subcarrier_amplitude, p.subcarrier_amplitude_back) == (50, 50)
    assert p.composite_in_chroma_lowpass and p.composite_out_chroma_lowpass and p.composite_out_chroma_lowpass_lite
    assert not p.emulating_vhs and not p.vhs_head_switching and p.vhs_chroma_vert_blend
    assert p.vhs_out_sharpen == 1.5 and p.video_scanline_phase_shift == 180
    assert abs(p.vhs_head_switching_point - (1.0 - 4.51 / 262.5)) < 1e-15
    assert abs(p.vhs_head_switching_phase - 0.99 / 262.5) < 1e-15
    assert abs(p.vhs_head_switching_phase_noise - (1.0 / 500) / 262.5) < 1e-18


def test_vhs_presets_and_order_dependence():
    sp = P("-vhs")
    assert (sp.emulating_vhs, sp.vhs_head_switching) == (1, 1)
    assert (sp.video_noise, sp.video_chroma_noise, sp.video_chroma_phase_noise, sp.video_chroma_loss) == (4, 16, 4, 4)
    ep = P("-vhs", "-vhs-speed", "ep")
    assert ep.output_vhs_tape_speed == cvs.VHS_EP
    assert (ep.video_noise, ep.video_chroma_noise, ep.video_chroma_phase_noise, ep.video_chroma_loss) == (6, 22, 6, 8)
    # reversed order: -vhs resets the noise levels to SP values but keeps the EP filters
    rev = P("-vhs-speed", "ep", "-vhs")
    assert rev.output_vhs_tape_speed == cvs.VHS_EP
    assert (rev.video_noise, rev.video_chroma_noise) == (4, 16)
    # -vhs-speed alone implies VHS emulation but NOT head switching (:1160-1189)
    lp = P("-vhs-speed", "lp")
    assert lp.emulating_vhs == 1 and lp.vhs_head_switching == 0
    assert (lp.video_noise, lp.video_chroma_noise, lp.video_chroma_phase_noise, lp.video_chroma_loss) == (5, 19, 5, 6)


def test_catv_presets_and_amplitude_back():
    p = P("-comp-catv")
    assert p.composite_preemphasis == 7 and p.composite_preemphasis_cut == 315000000 // 88
    assert p.video_chroma_phase_noise == 2
    # :1264-1265: back += (50 * pre * (315000000/88)) / (2 * cut), int += double
    assert p.subcarrier_amplitude_back == int(50 + (50 * 7 * (315000000 // 88)) / (2 * (315000000 // 88)))
    p4 = P("-comp-catv4")
    assert p4.composite_preemphasis_cut == (315000000 * 4) // 88 and p4.video_chroma_phase_noise == 6
    q = P("-subcarrier-amp", "40", "-comp-pre", "2", "-comp-cut", "2000000")
    assert q.subcarrier_amplitude == 40
    assert q.subcarrier_amplitude_back == int(40 + (50 * 2.0 * (315000000 // 88)) / (2 * 2000000.0))


def test_misc_switches():
    p = P("--vhs", "---noise", "7")                       # leading dashes stripped greedily (:980)
    assert p.emulating_vhs == 1 and p.video_noise == 7
    p = P("-tvstd", "pal")
    assert (p.output_height, p.output_ntsc) == (576, 0)
    p = P("-tvstd", "pal", "-tvstd", "ntsc", "-width", "1920")
    assert (p.output_width, p.output_height, p.output_ntsc) == (1920, 480, 1)
    p = P("-in-composite-lowpass", "0", "-out-composite-lowpass-lite", "0", "-vhs-svideo", "1", "-vhs-chroma-vblend", "0",
          "-nocolor-subcarrier", "-comp-phase", "270", "-comp-phase-offset", "3", "-chroma-dropout", "9",
          "-vhs-head-switching", "1", "-vhs-head-switching-point", "0.5", "-vhs-head-switching-phase", "0.25",
          "-vhs-head-switching-noise-level", "0")
    assert p.composite_in_chroma_lowpass == 0 and p.composite_out_chroma_lowpass_lite == 0
    assert p.vhs_svideo_out == 1 and p.vhs_chroma_vert_blend == 0 and p.nocolor_subcarrier == 1
    assert (p.video_scanline_phase_shift, p.video_scanline_phase_shift_offset) == (270, 3)
    assert p.video_chroma_loss == 9 and p.vhs_head_switching == 1
    assert (p.vhs_head_switching_point, p.vhs_head_switching_phase, p.vhs_head_switching_phase_noise) == (0.5, 0.25, 0.0)
    # accepted but outside the video path
    P("-i", "in.mp4", "-o", "out.mp4", "-d", "3", "-422", "-420", "-nocomp", "-vhs-hifi", "0", "-preemphasis", "1",
      "-deemphasis", "0", "-audio-hiss", "-60", "-vhs-linear-video-crosstalk", "-40", "-vhs-linear-high-boost", "0.5",
      "-yc-recomb", "2", "-nocolor-subcarrier-after-yc-sep")


@pytest.mark.parametrize("argv,status", [
    (["-bogus"], -2), (["positional"], -2), (["-comp-phase", "45"], -2), (["-width", "16"], -2),
    (["-d", "0"], -2), (["-d", "300"], -2), (["-tvstd", "secam"], -2), (["-vhs-speed", "slp"], -2),
    (["-noise"], -2), (["-h"], -6), (["-help"], -6),
    # help lists these but parse_argv() rejects them ("Unknown switch")
    (["-ss", "1"], -2), (["-t", "1"], -2), (["-a", "0"], -2), (["-bkey-feedback", "1"], -2),
])
def test_errors(argv, status):
    with pytest.raises(CvsError) as e:
        P(*argv)
    assert e.value.status == status


def test_draws_per_field():
    sp = P("-vhs")
    assert cvs.draws_per_field(sp, 1920, 1080, 0) == 3 * 540 * 1920 + 2 * 540 + 4
    assert cvs.draws_per_field(sp, 1920, 1080, 1) == 3111484
    assert cvs.draws_per_field(P(), 720, 480, 1) == 240 * 720
    assert cvs.draws_per_field(sp, 720, 481, 0) == 3 * 241 * 720 + 2 * 241 + 4           # odd height: parity 0 is longer
    assert cvs.draws_per_field(sp, 720, 481, 1) == 3 * 240 * 720 + 2 * 240 + 4
    assert cvs.draws_per_field(P("-vhs", "-vhs-head-switching-noise-level", "0"), 720, 480, 0) == 3 * 240 * 720 + 480
