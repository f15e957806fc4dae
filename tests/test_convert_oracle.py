"""oracle/convert_oracle.c on its own (CPU): sanity properties of the two conversions -- a scaler and a BT.601 matrix
behave as such.  The byte-for-byte pin against libswscale is tests/test_swscale_pin.py; BGRA sources and odd
destination widths go through the repository's own resampler (not pinned), whose arithmetic the first tests spell out."""
import numpy as np
import pytest

import helpers

BGRA, YUV420P, YUV422P, NV12 = 0, 1, 2, 3


def yuv_planes(w, h, fmt, seed):
    rng = np.random.default_rng(seed)
    cw, ch = (w + 1) // 2, (h if fmt == YUV422P else (h + 1) // 2)
    Y = rng.integers(16, 236, size=(h, w), dtype=np.uint8)
    U = rng.integers(16, 241, size=(ch, cw), dtype=np.uint8)
    V = rng.integers(16, 241, size=(ch, cw), dtype=np.uint8)
    if fmt == NV12:
        return [Y, np.ascontiguousarray(np.stack([U, V], axis=2).reshape(ch, 2 * cw))]
    return [Y, U, V]


def test_same_size_bgra_is_the_identity():
    w, h = 37, 23
    src = np.random.default_rng(1).integers(0, 1 << 32, size=(h, w), dtype=np.uint32)
    out = helpers.oracle_scale_to_bgra([src.view(np.uint8).reshape(h, 4 * w)], w, h, BGRA, w, h)
    assert np.array_equal(out, src)


@pytest.mark.parametrize("dw,dh", [(64, 48), (17, 9), (160, 120), (33, 50)])
def test_constant_pictures_stay_constant(dw, dh):
    """The 14-bit weights of every sample sum to exactly 16384: a flat picture stays flat at any ratio."""
    w, h = 40, 30
    flat = np.full((h, w), 0x80C04020, dtype=np.uint32)
    out = helpers.oracle_scale_to_bgra([flat.view(np.uint8).reshape(h, 4 * w)], w, h, BGRA, dw, dh)
    assert (out == 0x81C04020).all()          # colour exact; alpha 128 -> 129: the library widens a to a << 6 | a >> 2
    Y = np.full((h, w), 126, np.uint8)
    C_ = np.full(((h + 1) // 2, (w + 1) // 2), 128, np.uint8)
    out = helpers.oracle_scale_to_bgra([Y, C_, C_], w, h, YUV420P, dw, dh)
    g = int(out[0, 0]) & 0xFF
    assert abs(g - 1.164383 * (126 - 16)) <= 2.5               # (the floor of a luma table that sits ~1.3 codes low)
    assert (out == (0xFF000000 | (g << 16) | (g << 8) | g)).all()


def test_doubling_interpolates_between_neighbours():
    """2x enlargement of a horizontal ramp: destination sample i sits at (i + .5)/2 - .5: weights 3/4, 1/4."""
    w, h = 8, 2
    ramp = (np.arange(w, dtype=np.uint32) * 16)
    src = np.broadcast_to(ramp * 0x01010101, (h, w)).copy()
    out = helpers.oracle_scale_to_bgra([src.view(np.uint8).reshape(h, 4 * w)], w, h, BGRA, 2 * w, h) & 0xFF
    want = [0, 4, 12, 20, 28, 36, 44, 52, 60, 68, 76, 84, 92, 100, 108, 112]
    assert out[0].tolist() == want


def test_halving_averages_with_a_triangle():
    """2x reduction: half-width 2 source samples, taps at distance .5 and 1.5: weights 3/8, 3/8, 1/8, 1/8."""
    w, h = 8, 1
    vals = np.array([0, 0, 0, 0, 160, 160, 160, 160], dtype=np.uint32)
    src = (vals * 0x01010101)[None, :].copy()
    out = helpers.oracle_scale_to_bgra([src.view(np.uint8).reshape(h, 4 * w)], w, h, BGRA, 4, 1) & 0xFF
    assert out[0].tolist() == [0, 20, 140, 160]


@pytest.mark.parametrize("fmt", [YUV420P, YUV422P, NV12])
def test_yuv_to_bgra_is_bt601_limited_range(fmt):
    w, h = 16, 8
    planes = yuv_planes(w, h, fmt, 3)
    for p in planes:
        p[:] = 128
    planes[0][:] = np.array([16, 235] * (w // 2), np.uint8)[None, :]
    out = helpers.oracle_scale_to_bgra(planes, w, h, fmt, w, h)
    assert (out[:, 0::2] == 0xFF000000).all() and (out[:, 1::2] == 0xFFFDFDFD).all()    # black / white (253: the library's table)


def test_nv12_equals_planar_420():
    w, h = 31, 17
    p = yuv_planes(w, h, YUV420P, 5)
    ch, cw = p[1].shape
    nv = [p[0], np.ascontiguousarray(np.stack([p[1], p[2]], axis=2).reshape(ch, 2 * cw))]
    a = helpers.oracle_scale_to_bgra(p, w, h, YUV420P, 50, 40)
    b = helpers.oracle_scale_to_bgra(nv, w, h, NV12, 50, 40)
    assert np.array_equal(a, b)


def test_bgra_to_yuv_round_trip_within_quantisation():
    """BGRA -> YUV 4:2:2 (flat chroma pairs) -> BGRA returns every channel within a few codes."""
    w, h = 64, 16
    rng = np.random.default_rng(9)
    half = rng.integers(0, 1 << 24, size=(h, w // 2), dtype=np.uint32)
    src = np.repeat(half, 2, axis=1)                       # pairs of equal pixels: chroma subsampling loses nothing
    y, u, v = helpers.oracle_bgra_to_yuv(src, False)
    back = helpers.oracle_scale_to_bgra([y, u, v], w, h, YUV422P, w, h)
    # chroma is co-sited with the even columns: compare those
    d = np.abs((back[:, 0::2] & 0xFFFFFF).view(np.uint8).astype(int) - (src[:, 0::2] & 0xFFFFFF).view(np.uint8).astype(int))
    assert d.max() <= 5
