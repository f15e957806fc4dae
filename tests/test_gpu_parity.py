"""GPU parity: the CUDA path, called through the C ABI, against the oracle on the same inputs.

fp32 (production): max |delta| <= 1 LSB per 8-bit channel (north_star's bar).
fp64 (validation mode): bit-exact with the reference arithmetic.
"""
import numpy as np
import pytest

import helpers
import composite_video_simulator_b200 as cvs

pytestmark = pytest.mark.gpu

CASES = [
    (720, 480, 4, []),
    (720, 480, 4, ["-vhs", "-vhs-speed", "sp"]),
    (720, 480, 3, ["-vhs", "-vhs-speed", "ep"]),
    (724, 482, 3, ["-vhs", "-vhs-speed", "lp"]),
    (101, 67, 2, ["-vhs", "-vhs-speed", "lp", "-out-composite-lowpass-lite", "0"]),
    (720, 480, 2, ["-vhs", "-comp-catv", "-subcarrier-amp", "40"]),
    (720, 480, 2, ["-vhs", "-vhs-svideo", "1", "-comp-phase", "90"]),
    (720, 480, 2, ["-vhs", "-vhs-head-switching-phase", "0.001", "-vhs-head-switching-point", "0.5"]),
    (1920, 1080, 2, ["-vhs", "-vhs-speed", "sp"]),
]


@pytest.mark.parametrize("w,h,n,argv", CASES)
@pytest.mark.parametrize("use_double", [False, True])
def test_seam_matches_oracle(oracle, w, h, n, argv, use_double):
    p = helpers.params(*argv)
    frames = lambda k: helpers.stream_frame(w, h, k)
    want, g = helpers.run_oracle(oracle, p, frames, n, w, h)
    got = np.zeros((h, w), dtype=np.uint32)
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=4) as eng:
        eng.set_precision(use_double)
        for k in range(n):
            eng.composite_layer(got, frames(k), (k & 1) ^ 1, k)
        assert eng.rng_tell() == g.pos
    mx, nd, n2 = helpers.channel_diff(want, got)
    if use_double:
        assert mx == 0, (mx, nd)
    else:
        assert mx <= 1 and n2 == 0, (mx, nd, n2)
        assert nd <= 0.005 * want.size * 4


def test_batch_equals_sequential(oracle):
    w, h, n = 720, 480, 6
    p = helpers.params("-vhs")
    src = np.stack([helpers.stream_frame(w, h, k) for k in range(n)])
    seq = np.zeros((n, h, w), dtype=np.uint32)
    bat = np.zeros((n, h, w), dtype=np.uint32)
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=n) as eng:
        for k in range(n):
            eng.composite_layer(seq[k], src[k], (k & 1) ^ 1, k)
        pos = eng.rng_tell()
        eng.rng_seek(0)
        eng.composite_fields_host(bat, src, 0)
        assert eng.rng_tell() == pos
    assert np.array_equal(seq, bat)
