"""The audio step: cvs_audio_process() == the reference's composite_audio_process()
(ffmpeg_ntsc.cpp:901-970), sample for sample, including the rand() draws of the tape hiss.  CPU only."""
import os

import numpy as np
import pytest

import helpers
import composite_video_simulator_b200 as cvs
from composite_video_simulator_b200.api import Audio

CASES = {
    "default": [],
    "vhs_hifi": ["-vhs"],
    "vhs_linear_sp": ["-vhs", "-vhs-hifi", "0"],
    "vhs_linear_ep_boost": ["-vhs", "-vhs-speed", "ep", "-vhs-hifi", "0", "-vhs-linear-high-boost", "0.5"],
    "pal_linear_lp": ["-tvstd", "pal", "-vhs", "-vhs-speed", "lp", "-vhs-hifi", "0", "-vhs-linear-video-crosstalk", "-30"],
    "loud_hiss": ["-audio-hiss", "-20"],
    "no_emphasis": ["-preemphasis", "0", "-deemphasis", "0"],
}
PACKETS = (1024, 37, 4096, 1, 1500)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "audio_ntsc.npz")


def packets(channels, seed=3):
    sig = helpers.audio_signal(sum(PACKETS), channels, seed)
    out, at = [], 0
    for n in PACKETS:
        out.append(sig[at:at + n])
        at += n
    return out


def run_own(argv, start_pos=0):
    p = helpers.params(*argv)
    with Audio(params=p) as a:
        pos, out = start_pos, []
        for pk in packets(a.channels):
            buf = np.ascontiguousarray(pk.copy())
            pos = a.process(buf, pos)
            out.append(buf)
        return out, pos, a.channels


def known_switch(argv):
    try:
        helpers.params(*argv)
        return True
    except Exception:
        return False


@pytest.mark.parametrize("name", sorted(CASES))
def test_audio_matches_reference_code(emu, name):
    ref = helpers.load_refaudio()
    if ref is None:
        pytest.skip("oracle/_ref/librefaudio.so not built (needs /root/reference)")
    argv = CASES[name]
    if not known_switch(argv):
        pytest.skip("switch not accepted by parse_argv")
    p = helpers.params(*argv)
    got, pos, ch = run_own(argv)
    ref.refaudio_setup(helpers.C.byref(p))
    assert ref.refaudio_channels() == ch
    want, next_rand = helpers.run_refaudio(ref, p, packets(ch))
    for k, (a, b) in enumerate(zip(want, got)):
        assert np.array_equal(a, b), (name, k, np.argwhere(a != b)[:4])
    assert emu.emu_rand_at(helpers.C.c_ulonglong(pos)) == next_rand       # same number of hiss draws


def test_audio_shares_the_stream_position():
    """Starting the hiss at another stream position changes it exactly like seeking the generator would."""
    a0, pos0, ch = run_own(["-vhs", "-vhs-hifi", "0"], 0)
    a1, pos1, _ = run_own(["-vhs", "-vhs-hifi", "0"], 1000)
    assert pos1 - 1000 == pos0 == sum(PACKETS) * ch
    assert any(not np.array_equal(x, y) for x, y in zip(a0, a1))


def test_audio_golden():
    """Fixture produced by the reference's own code (tests/golden/make_golden_audio.py)."""
    g = np.load(GOLD, allow_pickle=False)
    for name in sorted(CASES):
        if not known_switch(CASES[name]):
            continue
        got, pos, ch = run_own(CASES[name])
        assert np.array_equal(np.concatenate(got), g[name]), name
        assert pos == int(g[name + "_draws"])


def test_audio_disabled_and_errors():
    p = helpers.params("-nocomp")                     # :1023: also switches the audio emulation off
    with Audio(params=p) as a:
        buf = helpers.audio_signal(64, a.channels, 1)
        keep = buf.copy()
        assert a.process(buf, 5) == 5
        assert np.array_equal(buf, keep)
    lib = cvs._lib.load()
    assert lib.cvs_audio_process(None, None, 0, None) == -1


@pytest.mark.parametrize("argv", [["-vhs", "-vhs-hifi", "0"], ["-vhs", "-vhs-speed", "ep"], ["-audio-hiss", "-30"]])
def test_interleaved_audio_and_video_share_one_stream(emu, oracle, ref, argv):
    """The reference's loop alternates audio packets and fields on ONE libc rand() stream (ffmpeg_ntsc.cpp:2157-2163,
    :2229).  Here the two engines are separate objects and the stream position is handed back and forth
    (cvs_rng_tell -> cvs_audio_process(..., &pos) -> cvs_rng_seek): the result must be what the reference's own two
    functions produce when they share libc's generator in one process."""
    refaudio = helpers.load_refaudio()
    if refaudio is None:
        pytest.skip("oracle/_ref/librefaudio.so not built (needs /root/reference)")
    w, h, nframes, fpf, packet = 96, 64, 3, 2, 300
    p = helpers.params(*(["-width", str(w)] + argv))
    frames = [helpers.stream_frame(w, h, k) for k in range(nframes)]
    with Audio(params=p) as a:
        ch = a.channels
        pcm = helpers.audio_signal(5000, ch, 9)
        want_pics, want_pcm = helpers.reference_av_loop(ref, refaudio, oracle, p, frames, pcm, w, h, fpf, 1, packet, bob=False)
        got_pcm = np.ascontiguousarray(pcm.copy())
        pos = helpers.C.c_ulonglong(0)
        done = 0
        dst = np.zeros((h, w), dtype=np.uint32)
        for current in range(nframes * fpf):
            while done < len(got_pcm) and helpers.audio_due(done, current):
                n = min(packet, len(got_pcm) - done)
                blk = np.ascontiguousarray(got_pcm[done:done + n])
                pos.value = a.process(blk, pos.value)
                got_pcm[done:done + n] = blk
                done += n
            src = frames[current // fpf]
            rc = emu.emu_composite_layer_ex(helpers.C.byref(p), 1, helpers.C.byref(pos), dst.ctypes.data_as(helpers.C.c_void_p),
                                            4 * w, src.ctypes.data_as(helpers.C.c_void_p), 4 * w, w, h, 0, 0, (current & 1) ^ 1,
                                            helpers.C.c_ulonglong(current), 0, 0)
            assert rc == 0
            assert np.array_equal(dst, want_pics[current]), current
        while done < len(got_pcm):
            n = min(packet, len(got_pcm) - done)
            blk = np.ascontiguousarray(got_pcm[done:done + n])
            pos.value = a.process(blk, pos.value)
            got_pcm[done:done + n] = blk
            done += n
    assert np.array_equal(got_pcm, want_pcm)
