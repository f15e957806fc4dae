"""Pin the oracle: the CPU restatement (oracle/ntsc_oracle.c) against the reference's OWN code
(oracle/_ref/libref.so, extracted at build time from /root/reference/ffmpeg_ntsc.cpp), against the
survey-time known-answer hashes and against the committed golden fixtures."""
import numpy as np
import pytest

import helpers

# FNV-1a-64 of the dst picture after n fields of 75% colour bars, default seed
KAT = [
    (720, 480, "sp", 4, 0xadc46d5e6368519b),
    (720, 480, "sp", 8, 0x1f6f1de0e08b3970),
    (720, 480, "sp", 60, 0x6efec46688da8257),
    (720, 480, "sp", 120, 0x8ff744cad3018bd9),
    (720, 480, "comp", 60, 0xdddc5e14c4d9b4a9),
    (1920, 1080, "sp", 16, 0x6dbd7301d1debb9f),
    (1920, 1080, "sp", 32, 0x69cd6d98c07172ab),
    (1920, 1080, "ep", 16, 0x34b7815ed20c013d),
    (3840, 2160, "comp", 8, 0x950d9b8328d39f81),
]
MODES = {"sp": ["-vhs", "-vhs-speed", "sp"], "ep": ["-vhs", "-vhs-speed", "ep"], "comp": []}


@pytest.mark.parametrize("w,h,mode,n,want", KAT)
def test_known_answer_hashes(oracle, w, h, mode, n, want):
    p = helpers.params(*MODES[mode])
    bars = helpers.bars_frame(w, h)
    dst, g = helpers.run_oracle(oracle, p, lambda k: bars, n, w, h)
    assert helpers.fnv1a64(dst) == want
    assert g.pos == sum(oracle.oracle_draws_per_field(__import__("ctypes").byref(p), w, h, (k & 1) ^ 1) for k in range(n))


SWEEP = [
    (720, 480, 3, []),
    (720, 480, 3, ["-vhs"]),
    (720, 480, 2, ["-vhs", "-vhs-speed", "lp"]),
    (720, 480, 2, ["-vhs", "-vhs-speed", "ep", "-out-composite-lowpass-lite", "0"]),
    (720, 480, 2, ["-vhs", "-in-composite-lowpass", "0", "-out-composite-lowpass", "0"]),
    (720, 480, 2, ["-comp-catv3", "-chroma-noise", "5"]),
    (720, 480, 2, ["-vhs", "-comp-catv", "-subcarrier-amp", "40"]),
    (720, 480, 2, ["-vhs", "-nocolor-subcarrier"]),
    (720, 480, 2, ["-vhs", "-vhs-svideo", "1", "-vhs-chroma-vblend", "0"]),
    (720, 480, 3, ["-vhs", "-comp-phase", "270", "-comp-phase-offset", "1"]),
    (720, 576, 2, ["-tvstd", "pal", "-vhs"]),
    (720, 480, 2, ["-vhs", "-vhs-head-switching-phase", "0.001", "-vhs-head-switching-point", "0.5"]),
    (720, 480, 2, ["-vhs", "-chroma-dropout", "30000"]),
    (33, 21, 3, ["-vhs", "-vhs-speed", "ep"]),
    (101, 67, 2, ["-vhs", "-vhs-speed", "lp"]),
    (1920, 1080, 1, ["-vhs", "-vhs-speed", "sp"]),
]


@pytest.mark.parametrize("w,h,n,argv", SWEEP)
def test_oracle_equals_reference_code(oracle, ref, w, h, n, argv):
    p = helpers.params(*argv)
    frames = lambda k: helpers.stream_frame(w, h, k)
    want = helpers.run_ref(ref, p, frames, n, w, h)
    got, _ = helpers.run_oracle(oracle, p, frames, n, w, h)
    assert np.array_equal(want, got)


def test_oracle_equals_reference_code_interlaced_source(oracle, ref):
    w, h, n = 160, 121, 3                      # odd height: the clamped last source row (:1599)
    p = helpers.params("-vhs")
    frames = lambda k: helpers.noise_frame(w, h, k)
    want = helpers.run_ref(ref, p, frames, n, w, h, interlaced=1, tff=1)
    got, _ = helpers.run_oracle(oracle, p, frames, n, w, h, interlaced=1, tff=1)
    assert np.array_equal(want, got)


@pytest.mark.parametrize("name", sorted(helpers.load_golden().keys()))
def test_oracle_matches_golden_fixture(oracle, name):
    argv, w, h, n, want = helpers.load_golden()[name]
    got, _ = helpers.run_oracle(oracle, helpers.params(*argv), lambda k: helpers.stream_frame(w, h, k), n, w, h)
    assert np.array_equal(want, got)


def test_guards(oracle):
    import ctypes as C
    p = helpers.params()
    g = helpers.OracleRng()
    oracle.oracle_rng_seed(C.byref(g), 1)
    buf = np.zeros((8, 8), dtype=np.uint32)
    # stride < 4*w: the reference returns without touching dst or rand() (:1580)
    rc = oracle.oracle_composite_layer(C.byref(p), C.byref(g), buf.ctypes.data_as(C.c_void_p), 16,
                                       buf.ctypes.data_as(C.c_void_p), 32, 8, 8, 0, 0, 0, C.c_ulonglong(0))
    assert rc != 0 and g.pos == 0
