"""The two picture conversions either side of the hot path PINNED against libswscale itself:
  * encoder side (ffmpeg_ntsc.cpp:2118-2131, 2266-2274): sws_getContext(w, h, BGRA -> YUV420P | YUV422P, SWS_BILINEAR) + sws_scale;
  * input side, InputFile::frame_copy_scale (:574-585, 603-610): sws_getContext(sw, sh, YUV420P | YUV422P | NV12 | BGRA -> dw,
    dh, BGRA, SWS_BILINEAR) + sws_scale (even destination widths: the table writers; odd ones and BGRA sources: the
    full-chroma writers).
oracle/convert_oracle.c restates what the library's portable C code does for those calls; here it is compared byte
for byte with the library (libswscale 9.1.100 of this image, tests/swscale_ref.py) where that exists, and with outputs
of the library committed as tests/golden/swscale_*.npz everywhere."""
import os

import numpy as np
import pytest

import helpers
import swscale_ref

HERE = os.path.dirname(os.path.abspath(__file__))
need_lib = pytest.mark.skipif(not swscale_ref.available(), reason="no libswscale on this machine (golden fixtures cover it)")

SIZES = [(64, 48), (720, 480), (66, 50), (100, 47), (101, 48), (101, 67), (33, 21), (8, 6), (1920, 1080), (719, 575)]


def picture(w, h, kind, seed=0):
    if kind == "random":
        return np.random.default_rng(seed).integers(0, 1 << 32, size=(h, w), dtype=np.uint32)
    if kind == "bars":
        return helpers.stream_frame(w, h, 3).astype(np.uint32)
    if kind == "extremes":                      # saturated primaries and their complements, two pixels wide
        cols = np.array([0xFF000000, 0xFFFFFFFF, 0xFFFF0000, 0xFF00FF00, 0xFF0000FF, 0xFFFFFF00, 0xFF00FFFF, 0xFFFF00FF], np.uint32)
        return np.broadcast_to(np.repeat(cols, 2)[np.arange(w) % 16], (h, w)).copy()
    raise ValueError(kind)


def lib_planes(src, fmt, c_code=True, flags=swscale_ref.SWS_BILINEAR):
    h, w = src.shape
    return swscale_ref.scale([src.view(np.uint8).reshape(h, 4 * w)], "bgra", w, h, fmt, w, h, flags=flags, c_code=c_code)


@need_lib
@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("fmt", ["yuv420p", "yuv422p"])
def test_oracle_equals_libswscale_c_code(w, h, fmt):
    for kind in ("random", "bars", "extremes"):
        src = picture(w, h, kind, seed=w * 31 + h)
        want = lib_planes(src, fmt)
        got = helpers.oracle_bgra_to_yuv(src, fmt == "yuv420p")
        for name, a, b in zip("yuv", got, want):
            assert np.array_equal(a, b), (kind, name, int(np.abs(a.astype(int) - b.astype(int)).max()))


@need_lib
def test_bitexact_flags_select_the_same_code_with_cpu_extensions_on():
    """SWS_ACCURATE_RND | SWS_BITEXACT with the CPU's extensions enabled == the C code the oracle is pinned to."""
    src = picture(720, 480, "random", 5)
    for fmt in ("yuv420p", "yuv422p"):
        a = lib_planes(src, fmt, c_code=False,
                       flags=swscale_ref.SWS_BILINEAR | swscale_ref.SWS_ACCURATE_RND | swscale_ref.SWS_BITEXACT)
        b = helpers.oracle_bgra_to_yuv(src, fmt == "yuv420p")
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), fmt


@need_lib
def test_x86_code_of_the_library_stays_within_one_code_of_its_c_code():
    """What the reference gets on an x86 host with plain SWS_BILINEAR: the library's SIMD vertical scaler rounds
    differently from its C code -- luma and 4:2:2 identical, 4:2:0 chroma +-1 on a few percent of the samples."""
    src = picture(720, 480, "random", 6)
    for fmt in ("yuv420p", "yuv422p"):
        simd, ccode = lib_planes(src, fmt, c_code=False), lib_planes(src, fmt, c_code=True)
        assert np.array_equal(simd[0], ccode[0])
        for a, b in zip(simd[1:], ccode[1:]):
            d = np.abs(a.astype(int) - b.astype(int))
            assert d.max() <= 1 and (d > 0).mean() < 0.15
            if fmt == "yuv422p":
                assert d.max() == 0


def test_oracle_equals_golden_outputs_of_libswscale():
    g = np.load(os.path.join(HERE, "golden", "swscale_bgra_yuv.npz"))
    names = sorted({k[:-4] for k in g.files if k.endswith("_src")})
    assert len(names) >= 8
    for n in names:
        src = g[n + "_src"]
        got = helpers.oracle_bgra_to_yuv(src, "yuv420p" in n)
        for pl, a in zip("yuv", got):
            assert np.array_equal(a, g["%s_%s" % (n, pl)]), (n, pl)


@pytest.mark.parametrize("srcn,dstn,one", [(480, 240, 4096), (1080, 540, 4096), (67, 34, 4096), (101, 51, 16384), (5, 3, 4096)])
def test_filter_bank_properties(srcn, dstn, one):
    """Every row of the bilinear filter bank sums to `one`, taps stay inside the picture, positions are monotonic;
    an exact 2:1 reduction is 1/8 3/8 3/8 1/8 away from the borders."""
    pos, coef = helpers.oracle_sws_filter(srcn, dstn, one)
    assert (coef.sum(axis=1) == one).all()
    assert pos.min() >= 0 and (pos + coef.shape[1]).max() <= srcn and (np.diff(pos) >= 0).all()
    if srcn == 2 * dstn:
        assert coef[dstn // 2].tolist() == [one // 8, 3 * one // 8, 3 * one // 8, one // 8]
        assert coef[0].tolist()[:3] == [one // 2, 3 * one // 8, one // 8] and pos[0] == 0


def test_product_filter_bank_equals_the_oracles():
    """The table builder of the product (csrc/sws_filter.cpp, through the C ABI; no GPU involved) against the oracle's
    restatement, which is pinned above: every geometry a conversion can ask for in a sweep of sizes."""
    import ctypes as C
    from composite_video_simulator_b200 import _lib
    lib = _lib.load()
    cases = [(h, (h + 1) // 2, 1 << 12) for h in list(range(3, 80)) + [240, 480, 486, 575, 576, 720, 1080, 1081, 2160]]
    cases += [(w, (w + 1) // 2, 1 << 14) for w in list(range(3, 80, 2)) + [719, 721, 1919, 3839]]
    cases += [(h, h, 1 << 12) for h in (1, 2, 5, 480)]
    for srcn, dstn, one in cases:
        pos = np.zeros(dstn, np.int32)
        coef = np.zeros(dstn * 16, np.int32)
        taps = lib.cvs_sws_bilinear_bank(srcn, dstn, one, pos.ctypes.data_as(C.POINTER(C.c_int32)),
                                         coef.ctypes.data_as(C.POINTER(C.c_int32)), 16)
        assert taps > 0, (srcn, dstn, taps)
        opos, ocoef = helpers.oracle_sws_filter(srcn, dstn, one)
        assert taps == ocoef.shape[1], (srcn, dstn, taps, ocoef.shape)
        assert np.array_equal(pos, opos) and np.array_equal(coef[:dstn * taps].reshape(dstn, taps), ocoef), (srcn, dstn)


# ---- the input side ---------------------------------------------------------------------------------------------
FMT_CODE = {"yuv420p": 1, "yuv422p": 2, "nv12": 3}          # the product's CVS_PIX_* / the oracle's format argument
SCALES = [(720, 480, 720, 480), (720, 481, 720, 481), (640, 480, 720, 480), (352, 288, 720, 480), (1920, 1080, 720, 480),
          (720, 576, 720, 480), (720, 480, 360, 240), (351, 287, 720, 480), (20, 10, 320, 240), (5, 3, 8, 8), (1280, 720, 1920, 1080),
          # odd destination widths: the library's full-chroma-interpolation writers (32-bit arithmetic, no tables)
          (352, 288, 721, 480), (353, 289, 353, 289), (640, 480, 721, 480), (1920, 1080, 721, 480), (720, 576, 719, 480),
          (720, 480, 361, 240), (5, 3, 9, 8)]


def source_planes(fmt, w, h, seed, legal=False):
    rng = np.random.default_rng(seed)
    lo, hi = (16, 236) if legal else (0, 256)
    return [rng.integers(lo, hi, size=s, dtype=np.uint8) for s in swscale_ref.plane_shapes(fmt, w, h)]


@need_lib
@pytest.mark.parametrize("sw,sh,dw,dh", SCALES)
@pytest.mark.parametrize("fmt", sorted(FMT_CODE))
def test_scaler_oracle_equals_libswscale_c_code(fmt, sw, sh, dw, dh):
    for legal in (False, True):
        planes = source_planes(fmt, sw, sh, sw + 3 * dh + legal, legal)
        want = swscale_ref.scale(planes, fmt, sw, sh, "bgra", dw, dh, c_code=True)[0].view(np.uint32).reshape(dh, dw)
        got = helpers.oracle_scale_to_bgra(planes, sw, sh, FMT_CODE[fmt], dw, dh)
        assert np.array_equal(got, want), int((got != want).sum())


@need_lib
def test_colour_tables_on_every_yuv_triple():
    """YUV420P at the same size takes the library's direct converter, a pure function of (Y, U, V): all 2^24 triples."""
    W = H = 4096
    bx, by = np.meshgrid(np.arange(W // 2), np.arange(H // 2))
    U, V = (bx % 256).astype(np.uint8), (by % 256).astype(np.uint8)
    xx, yy = np.meshgrid(np.arange(W), np.arange(H))
    Y = (((xx // 2) // 256) * 32 + ((yy // 2) // 256) * 4 + (yy % 2) * 2 + (xx % 2)).astype(np.uint8)
    assert len(np.unique(Y)) == 256
    want = swscale_ref.scale([Y, U, V], "yuv420p", W, H, "bgra", W, H, c_code=True)[0].view(np.uint32).reshape(H, W)
    got = helpers.oracle_scale_to_bgra([Y, U, V], W, H, 1, W, H)
    assert np.array_equal(got, want)


def test_scaler_oracle_equals_golden_outputs_of_libswscale():
    g = np.load(os.path.join(HERE, "golden", "swscale_to_bgra.npz"))
    names = sorted({k[:-5] for k in g.files if k.endswith("_bgra")})
    assert len(names) >= 18
    for n in names:
        fmt, s, d = n.split("_")
        (sw, sh), (dw, dh) = [tuple(int(v) for v in q.split("x")) for q in (s, d)]
        planes = [g["%s_p%d" % (n, i)] for i in range({"nv12": 2, "bgra": 1}.get(fmt, 3))]
        got = helpers.oracle_scale_to_bgra(planes, sw, sh, {"bgra": 0, **FMT_CODE}[fmt], dw, dh)
        assert np.array_equal(got, g[n + "_bgra"]), n


@need_lib
def test_random_geometries_against_the_library():
    """A seeded sweep over sizes no case above names: 40 scaler geometries (ratios from 1/6 to 6, odd source sizes, tiny
    pictures) and 20 encoder-side sizes, every format, oracle == library."""
    rng = np.random.default_rng(20261017)
    fmts = sorted(FMT_CODE)
    for k in range(40):
        sw, sh = int(rng.integers(3, 400)), int(rng.integers(3, 300))
        dw = 2 * int(rng.integers(max(2, sw // 12), 3 * sw + 2)) + (k % 4 == 3)     # every fourth one odd
        dh = int(rng.integers(max(2, sh // 6), 6 * sh + 1))
        dw, dh = min(dw, 1200), min(dh, 900)
        if sw > 16 * dw or sh > 16 * dh:
            continue
        fmt = fmts[k % 3]
        planes = source_planes(fmt, sw, sh, 1000 + k)
        want = swscale_ref.scale(planes, fmt, sw, sh, "bgra", dw, dh, c_code=True)[0].view(np.uint32).reshape(dh, dw)
        got = helpers.oracle_scale_to_bgra(planes, sw, sh, FMT_CODE[fmt], dw, dh)
        assert np.array_equal(got, want), (fmt, sw, sh, dw, dh, int((got != want).sum()))
    for k in range(20):
        w, h = int(rng.integers(3, 500)), int(rng.integers(3, 400))
        src = picture(w, h, "random", 2000 + k)
        fmt = ("yuv420p", "yuv422p")[k % 2]
        want = lib_planes(src, fmt)
        got = helpers.oracle_bgra_to_yuv(src, fmt == "yuv420p")
        assert all(np.array_equal(a, b) for a, b in zip(got, want)), (fmt, w, h)


# ---- BGRA sources at another size (the library's RGB -> YUV(A) -> RGB route) -------------------------------------------
BGRA_SCALES = [(640, 480, 720, 480), (352, 288, 720, 480), (720, 576, 720, 480), (720, 480, 720, 482), (64, 48, 81, 61),
               (1920, 1080, 720, 480), (720, 480, 360, 240), (720, 480, 400, 480), (720, 480, 359, 240), (100, 67, 50, 40),
               (100, 67, 51, 40), (101, 67, 51, 40), (5, 3, 9, 8), (720, 480, 720, 480)]


@need_lib
@pytest.mark.parametrize("sw,sh,dw,dh", BGRA_SCALES)
def test_bgra_source_oracle_equals_libswscale_c_code(sw, sh, dw, dh):
    """Alpha included: random bytes in all four channels (full chroma when the width shrinks by less than 2, chroma from
    pixel pairs otherwise; every writer variant of the vertical bank; a same-size source is a copy)."""
    img = np.random.default_rng(sw * 3 + dh).integers(0, 256, size=(sh, 4 * sw), dtype=np.uint8)
    want = swscale_ref.scale([img], "bgra", sw, sh, "bgra", dw, dh, c_code=True)[0].view(np.uint32).reshape(dh, dw)
    got = helpers.oracle_scale_to_bgra([img], sw, sh, 0, dw, dh)
    assert np.array_equal(got, want), int((got != want).sum())


@need_lib
@pytest.mark.parametrize("sw,sh,dw,dh", [(101, 67, 50, 40), (33, 21, 7, 5), (721, 480, 360, 240), (101, 67, 51, 40), (100, 67, 50, 40)])
def test_bgra_odd_width_corner(sw, sh, dw, dh):
    """The one geometry the PRODUCT does not take through the library's route yet: a BGRA source of odd width reduced to
    half its width or less.  The library's rule is known -- an odd source width keeps chroma per pixel at every ratio -- and
    restated (oracle_sws_bgra_to_bgra == the library here); the product still uses the repository's resampler for it,
    because the rule was found after the round's last GPU run.  Next step: route it to k_sws_bgra_to_bgra (half = false)."""
    img = np.random.default_rng(sw + dw).integers(0, 256, size=(sh, 4 * sw), dtype=np.uint8)
    want = swscale_ref.scale([img], "bgra", sw, sh, "bgra", dw, dh, c_code=True)[0].view(np.uint32).reshape(dh, dw)
    assert np.array_equal(helpers.oracle_sws_bgra_route(img, sw, sh, dw, dh), want)
