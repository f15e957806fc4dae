"""GPU parity of the 4:2:2 path (include/cvs_yuv422.h) through the C ABI: BIT-EXACT against the oracle
(the kernel evaluates the reference's double arithmetic in the reference's order; no tolerance)."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc():
    return helpers.load_oracle422()


@pytest.fixture(scope="module")
def y422():
    from composite_video_simulator_b200 import yuv422
    return yuv422


def stack(frames):
    return [np.ascontiguousarray(np.stack([f[pl] for f in frames])) for pl in range(3)]


CASES = [
    (720, 480, 2, []),
    (720, 480, 3, ["-vhs"]),
    (720, 480, 2, ["-vhs", "-vhs-speed", "lp"]),
    (720, 480, 2, ["-vhs", "-vhs-speed", "ep", "-out-composite-lowpass", "0"]),
    (720, 480, 2, ["-vhs", "-in-composite-lowpass", "0", "-out-composite-lowpass", "0", "-out-composite-lowpass-lite", "0"]),
    (720, 480, 2, ["-comp-catv3", "-chroma-noise", "5"]),
    (720, 480, 2, ["-vhs", "-comp-catv", "-subcarrier-amp", "40"]),
    (720, 480, 2, ["-vhs", "-nocolor-subcarrier"]),
    (720, 480, 2, ["-nocolor-subcarrier-after-yc-sep"]),
    (720, 480, 2, ["-vhs", "-vhs-svideo", "1", "-vhs-chroma-vblend", "0"]),
    (724, 480, 3, ["-vhs", "-comp-phase", "270", "-comp-phase-offset", "1"]),
    (724, 480, 3, ["-comp-phase", "90", "-yc-recomb", "2"]),
    (724, 576, 2, ["-tvstd", "pal", "-vhs"]),
    (720, 480, 2, ["-vhs", "-vhs-head-switching-point", "0.9501", "-vhs-head-switching-noise-level", "0.0001"]),
    (720, 480, 2, ["-vhs", "-chroma-dropout", "30000"]),
    (720, 480, 2, ["-vhs", "-noise", "300", "-chroma-noise", "200", "-chroma-phase-noise", "40"]),
    (34, 21, 3, ["-vhs", "-vhs-speed", "ep"]),
    (102, 67, 2, ["-vhs", "-vhs-speed", "lp"]),
    (8, 3, 2, []),
    (2, 2, 2, ["-vhs"]),
    (64, 32, 2, ["-vhs"]),
    (24, 10, 3, ["-vhs", "-vhs-speed", "ep"]),
    (16, 40, 2, ["-vhs", "-vhs-speed", "lp", "-chroma-dropout", "2000"]),
    (720, 480, 2, ["-vhs", "-vhs-speed", "ep"]),
    (1920, 1080, 2, ["-vhs", "-vhs-speed", "sp"]),
    (3840, 2160, 1, ["-vhs", "-vhs-speed", "ep"]),
]


@pytest.fixture(params=["fast", "general"])
def kernel(request):
    """Rows of whole blocks with the common switches run k_yuv422_fast; CVS422_GENERAL=1 sends them through the
    general kernel k_yuv422 as well (the library reads the variable at every launch)."""
    if request.param == "general":
        os.environ["CVS422_GENERAL"] = "1"
    yield request.param
    os.environ.pop("CVS422_GENERAL", None)


def fast_kernel_applies(w, argv):
    general = ("-subcarrier-amp", "-nocolor-subcarrier", "-nocolor-subcarrier-after-yc-sep", "-yc-recomb", "-comp-pre", "-comp-catv",
               "-comp-catv2", "-comp-catv3", "-comp-catv4")
    return w % 8 == 0 and not any(a in general for a in argv)


@pytest.mark.parametrize("w,h,n,argv", CASES)
def test_seam_equals_oracle(orc, y422, w, h, n, argv, kernel):
    if kernel == "general" and not fast_kernel_applies(w, argv):
        pytest.skip("already the general kernel")
    p = helpers.params422(*argv)
    want, g = helpers.run_oracle422(orc, p, w, h, n)
    with y422.Yuv422Engine(argv, max_w=w, max_h=h, max_batch=1) as eng:
        for k in range(n):
            Y, U, V = helpers.yuv422_frame(w, h, k, 2)
            eng.composite_video_process(Y, U, V, w, (k & 1) ^ 1, k)
            for pl, got in enumerate((Y, U, V)):
                d = np.argwhere(want[k][pl] != got)
                assert d.size == 0, (k, pl, d[:5])
        assert eng.rng_tell() == g.pos


def test_batch_equals_sequential_and_device_form(orc, y422):
    import torch
    w, h, n = 720, 480, 7
    argv = ["-vhs", "-vhs-speed", "ep"]
    p = helpers.params422(*argv)
    want, _ = helpers.run_oracle422(orc, p, w, h, n, pad=16)
    frames = [helpers.yuv422_frame(w, h, k, 16) for k in range(n)]
    Y, U, V = stack(frames)
    with y422.Yuv422Engine(argv, max_w=w, max_h=h, max_batch=n) as eng:
        eng.process_fields_host(Y, U, V, w, 0)
        for k in range(n):
            for pl, got in enumerate((Y[k], U[k], V[k])):
                assert np.array_equal(want[k][pl], got), (k, pl)
        # device pointers, split 3 + 4, on a torch stream
        Yd, Ud, Vd = [torch.from_numpy(a).cuda() for a in stack(frames)]
        st = torch.cuda.Stream()
        eng.set_stream(st.cuda_stream)
        eng.rng_seek(0)
        with torch.cuda.stream(st):
            eng.process_fields_device(Yd[:3], Ud[:3], Vd[:3], w, 0)
            eng.process_fields_device(Yd[3:], Ud[3:], Vd[3:], w, 3)
        eng.synchronize()
        eng.set_stream(0)
        for k in range(n):
            for pl, got in enumerate((Yd[k], Ud[k], Vd[k])):
                assert np.array_equal(want[k][pl], got.cpu().numpy()), (k, pl)


@pytest.mark.parametrize("w,h,n,argv", [(64, 67, 7, ["-vhs"]),                       # 34 / 33 rows per parity: two fields meet in a warp
                                        (96, 65, 5, ["-vhs", "-vhs-speed", "ep"]),   # 33 / 32 rows
                                        (64, 61, 5, ["-vhs"]),                       # 31 / 30 rows: not packed
                                        (720, 480, 6, [])])
def test_packed_rows_batch_equals_oracle(orc, y422, w, h, n, argv):
    """The batch's rows are cut into warps of 31 across field boundaries (`packed`, halo records per warp of the
    launch): bit-exact against the oracle, and the same bytes as the per-field mapping."""
    import os
    p = helpers.params422(*argv)
    want, _ = helpers.run_oracle422(orc, p, w, h, n, pad=4)
    got = {}
    for packed in ("1", "0"):
        os.environ["CVS_PACKED_ROWS"] = packed
        try:
            Y, U, V = stack([helpers.yuv422_frame(w, h, k, 4) for k in range(n)])
            with y422.Yuv422Engine(argv, max_w=w, max_h=h, max_batch=n) as eng:
                eng.process_fields_host(Y, U, V, w, 0)
            got[packed] = (Y, U, V)
        finally:
            os.environ.pop("CVS_PACKED_ROWS", None)
    for k in range(n):
        for pl in range(3):
            assert np.array_equal(want[k][pl], got["1"][pl][k]), (k, pl)
            assert np.array_equal(got["0"][pl][k], got["1"][pl][k]), (k, pl)


def test_tight_rows_and_unaligned_strides(orc, y422):
    """linesize == width (the two bytes past a row are the next row's) and odd linesizes (byte path)."""
    for w, h, pad in ((64, 32, 0), (102, 31, 3), (720, 64, 5)):
        p = helpers.params422("-vhs")
        want, _ = helpers.run_oracle422(orc, p, w, h, 2, pad=pad)
        with y422.Yuv422Engine(["-vhs"], max_w=w, max_h=h) as eng:
            for k in range(2):
                Y, U, V = helpers.yuv422_frame(w, h, k, pad)
                eng.composite_video_process(Y, U, V, w, (k & 1) ^ 1, k)
                for pl, got in enumerate((Y, U, V)):
                    assert np.array_equal(want[k][pl], got), (w, k, pl)


def test_untouched_rows_and_padding(y422):
    w, h = 128, 40
    with y422.Yuv422Engine(["-vhs"], max_w=w, max_h=h) as eng:
        Y, U, V = helpers.yuv422_frame(w, h, 0, 6)
        Y0, U0, V0 = Y.copy(), U.copy(), V.copy()
        eng.composite_video_process(Y, U, V, w, 1, 0)
        for a, b, ww in ((Y, Y0, w), (U, U0, w // 2), (V, V0, w // 2)):
            assert np.array_equal(a[0::2], b[0::2])            # the other field
            assert np.array_equal(a[:, ww:], b[:, ww:])        # padding columns
        assert not np.array_equal(Y[1::2, :w], Y0[1::2, :w])


def test_invalid_arguments(y422):
    w, h = 64, 16
    with y422.Yuv422Engine([], max_w=w, max_h=h) as eng:
        Y, U, V = helpers.yuv422_frame(w, h, 0, 2)
        with pytest.raises(y422.Yuv422Error) as e:
            eng.composite_video_process(Y, U, V, w - 1, 1, 0)          # odd width
        assert e.value.status == -1
        with pytest.raises(y422.Yuv422Error) as e:
            eng.composite_video_process(Y, U, V, w, 0, 0)              # field must be (fieldno & 1) ^ 1
        assert e.value.status == -1
        with pytest.raises(y422.Yuv422Error) as e:
            eng.composite_video_process(Y, U, V, w + 64, 1, 0)         # beyond the linesize / capacity
        assert e.value.status in (-1, -5)
    p = helpers.params422("-yc-recomb", "9")
    with y422.Yuv422Engine(p, max_w=w, max_h=h) as eng:
        Y, U, V = helpers.yuv422_frame(w, h, 0, 2)
        with pytest.raises(y422.Yuv422Error) as e:
            eng.composite_video_process(Y, U, V, w, 1, 0)
        assert e.value.status == -8


def test_golden_fixtures(y422):
    import glob
    files = sorted(glob.glob(os.path.join(helpers.GOLDEN_DIR, "yuv422_*.npz")))
    assert files
    for f in files:
        z = np.load(f)
        argv = [a for a in str(z["argv"]).split(" ") if a]
        w, h, n = int(z["w"]), int(z["h"]), int(z["n"])
        with y422.Yuv422Engine(argv, max_w=w, max_h=h) as eng:
            for k in range(n):
                Y, U, V = helpers.yuv422_frame(w, h, k, 2)
                eng.composite_video_process(Y, U, V, w, (k & 1) ^ 1, k)
                assert np.array_equal(Y, z["Y%d" % k]) and np.array_equal(U, z["U%d" % k]) and np.array_equal(V, z["V%d" % k]), (f, k)


RENDER = [
    (64, 48, 48, 0, 0, 0, 0),
    (64, 48, 36, 0, 0, 0, 0),
    (64, 48, 100, 1, 0, 0, 0),
    (64, 48, 58, 1, 1, 1, 0),
    (64, 48, 58, 1, 1, 1, 1),
    (64, 48, 58, 0, 1, 0, 0),
    (720, 480, 1080, 1, 1, 1, 1),
    (720, 480, 576, 0, 0, 1, 0),
    (1920, 1080, 2160, 1, 0, 0, 0),
]


@pytest.mark.parametrize("dst_w,dst_h,src_h,is420,il,tff,second", RENDER)
def test_render_field_equals_oracle(orc, y422, dst_w, dst_h, src_h, is420, il, tff, second):
    import torch
    I3, V3 = C.c_int * 3, C.c_void_p * 3
    rs = np.random.RandomState(11)
    ch = src_h // 2 if is420 else src_h
    ls = [dst_w + 16, dst_w // 2 + 8, dst_w // 2 + 8]
    src = [rs.randint(0, 256, size=(src_h, ls[0]), dtype=np.uint8),
           rs.randint(0, 256, size=(ch, ls[1]), dtype=np.uint8), rs.randint(0, 256, size=(ch, ls[2]), dtype=np.uint8)]
    with y422.Yuv422Engine([], max_w=dst_w, max_h=dst_h) as eng:
        for field in (0, 1):
            want = [np.full((dst_h, ls[i]), 7 + i, dtype=np.uint8) for i in range(3)]
            orc.oracle422_render_field(V3(*[a.ctypes.data for a in want]), I3(*ls), C.c_int(dst_h),
                                       V3(*[a.ctypes.data for a in src]), I3(*ls), C.c_int(src_h), I3(*ls), C.c_int(is420),
                                       C.c_int(il), C.c_int(tff), C.c_int(second), C.c_uint(field))
            dst = [torch.full((dst_h, ls[i]), 7 + i, dtype=torch.uint8, device="cuda") for i in range(3)]
            srcd = [torch.from_numpy(a).cuda() for a in src]
            eng.render_field_device(dst, srcd, ls, is420, il, tff, second, field)
            eng.synchronize()
            for pl in range(3):
                assert np.array_equal(want[pl], dst[pl].cpu().numpy()), (field, pl)


def test_1080p_stream_64_fields(orc, y422):
    """BASELINE config 2 geometry on the 4:2:2 path: 64 fields in one batch, every 8th checked against the
    oracle, plus stream-position bookkeeping."""
    w, h, n = 1920, 1080, 64
    argv = ["-vhs", "-vhs-speed", "sp"]
    p = helpers.params422(*argv)
    frames = [helpers.yuv422_frame(w, h, k, 32) for k in range(n)]
    Y, U, V = stack(frames)
    with y422.Yuv422Engine(argv, max_w=w, max_h=h, max_batch=n) as eng:
        eng.process_fields_host(Y, U, V, w, 0)
        pos = 0
        for k in range(n):
            if k % 8 == 0:
                g = helpers.OracleRng()
                orc.oracle_rng_seed(C.byref(g), 1)
                # position the oracle's generator at the field's first draw
                skip = pos
                for _ in range(skip):
                    orc.oracle_rng_next(C.byref(g))
                want, _ = helpers.run_oracle422(orc, p, w, h, 1, first=k, frames=lambda kk: helpers.yuv422_frame(w, h, kk, 32), rng=g)
                for pl, got in enumerate((Y[k], U[k], V[k])):
                    assert np.array_equal(want[0][pl], got), (k, pl)
            pos += orc.oracle422_draws_per_field(C.byref(p), w, h, (k & 1) ^ 1)
        assert eng.rng_tell() == pos
