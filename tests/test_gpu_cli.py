"""The raw-frame host (tools/cvs_ntsc_raw.cpp) against the reference's field loop replayed with the oracle:
ring of `delay` zero-initialised pictures, composite_layer + line doubling per output field, one picture
emitted per field (ffmpeg_ntsc.cpp:2202-2282)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

TOOL = os.path.join(helpers.ROOT, "composite_video_simulator_b200", "cvs_ntsc_raw")


def reference_loop(oracle, p, frames, w, h, fields_per_frame, delay):
    g = helpers.OracleRng()
    oracle.oracle_rng_seed(C.byref(g), 1)
    ring = [np.zeros((h, w), dtype=np.uint32) for _ in range(delay)]
    out, idx = [], 0
    for current in range(len(frames) * fields_per_frame):
        src = frames[current // fields_per_frame]
        f = (current & 1) ^ 1
        pic = ring[idx]
        oracle.oracle_composite_layer(C.byref(p), C.byref(g), pic.ctypes.data_as(C.c_void_p), 4 * w,
                                      src.ctypes.data_as(C.c_void_p), 4 * w, w, h, 0, 0, f, C.c_ulonglong(current))
        oracle.oracle_bob(pic.ctypes.data_as(C.c_void_p), 4 * w, w, h, f)
        out.append(pic.copy())
        idx = (idx + 1) % delay
    return np.stack(out)


@pytest.mark.parametrize("w,h,delay,batch,argv", [
    (160, 120, 1, 3, ["-vhs"]),
    (160, 120, 2, 5, ["-vhs", "-vhs-speed", "ep"]),
    (164, 121, 3, 4, []),
])
def test_cli_matches_reference_field_loop(oracle, tmp_path, w, h, delay, batch, argv):
    nframes, fpf = 4, 2
    frames = [helpers.stream_frame(w, h, k) for k in range(nframes)]
    inp, outp = str(tmp_path / "in.bgra"), str(tmp_path / "out.bgra")
    np.stack(frames).tofile(inp)
    cmd = [TOOL, "-i", inp, "-o", outp, "-width", str(w), "-height", str(h), "-d", str(delay), "-batch", str(batch),
           "-fields-per-frame", str(fpf), "-double"] + argv
    subprocess.run(cmd, check=True, stderr=subprocess.DEVNULL)
    got = np.fromfile(outp, dtype=np.uint32).reshape(-1, h, w)
    want = reference_loop(oracle, helpers.params(*(["-width", str(w)] + argv)), frames, w, h, fpf, delay)
    assert got.shape == want.shape
    for k in range(len(want)):
        assert np.array_equal(got[k], want[k]), k


def test_cli_rejects_what_the_reference_rejects():
    r = subprocess.run([TOOL, "-i", "x", "-o", "y", "-bogus"], capture_output=True)
    assert r.returncode == 1
    r = subprocess.run([TOOL, "-vhs"], capture_output=True)
    assert r.returncode == 1 and b"No input files specified" in r.stderr


@pytest.mark.parametrize("w,h,delay,batch,argv", [
    (160, 120, 1, 4, ["-vhs", "-vhs-hifi", "0"]),
    (96, 64, 2, 3, ["-vhs", "-vhs-speed", "ep"]),
])
def test_cli_audio_and_video_share_the_rand_stream(oracle, ref, tmp_path, w, h, delay, batch, argv):
    """`-audio-in/-audio-out`: the host alternates audio packets and fields like the reference's loop
    (ffmpeg_ntsc.cpp:2157-2163, :2229) and hands the rand() position between the two engines; against the reference's
    own composite_layer() and composite_audio_process() sharing libc's generator in one process.  Small audio packets
    force batches to be cut where a packet is due."""
    refaudio = helpers.load_refaudio()
    if refaudio is None:
        pytest.skip("oracle/_ref/librefaudio.so not built (needs /root/reference)")
    nframes, fpf, packet = 5, 2, 500
    p = helpers.params(*(["-width", str(w)] + argv))
    import composite_video_simulator_b200 as cvs
    ch = cvs._lib.load().cvs_audio_channels(C.byref(p))
    frames = [helpers.stream_frame(w, h, k) for k in range(nframes)]
    pcm = helpers.audio_signal(9000, ch, 4)
    want_pics, want_pcm = helpers.reference_av_loop(ref, refaudio, oracle, p, frames, pcm, w, h, fpf, delay, packet)
    inp, outp = str(tmp_path / "in.bgra"), str(tmp_path / "out.bgra")
    ain, aout = str(tmp_path / "in.s16"), str(tmp_path / "out.s16")
    np.stack(frames).tofile(inp)
    pcm.tofile(ain)
    cmd = [TOOL, "-i", inp, "-o", outp, "-width", str(w), "-height", str(h), "-d", str(delay), "-batch", str(batch),
           "-fields-per-frame", str(fpf), "-double", "-audio-in", ain, "-audio-out", aout, "-audio-packet", str(packet)] + argv
    subprocess.run(cmd, check=True, stderr=subprocess.DEVNULL)
    got = np.fromfile(outp, dtype=np.uint32).reshape(-1, h, w)
    got_pcm = np.fromfile(aout, dtype=np.int16).reshape(-1, ch)
    assert got.shape == want_pics.shape
    for k in range(len(want_pics)):
        assert np.array_equal(got[k], want_pics[k]), k
    assert np.array_equal(got_pcm, want_pcm)
