"""cvs_scale_to_bgra_device (the input side: InputFile::frame_copy_scale(), ffmpeg_ntsc.cpp:544-613):
decoder picture -> BGRA at the output size on the device.  PINNED (planar YUV and BGRA sources, even and odd output widths):
the kernel is compared bit for bit with oracle/convert_oracle.c (which tests/test_swscale_pin.py pins against libswscale
9.1.100 itself), with the library directly when the GPU box has it, and with its committed outputs.  One corner stays on
the repository's own resampler (not pinned; same oracle file): a BGRA source of odd width reduced to half or less."""
import os

import numpy as np
import pytest

import helpers
import swscale_ref
import composite_video_simulator_b200 as cvs

pytestmark = pytest.mark.gpu

BGRA, YUV420P, YUV422P, NV12 = 0, 1, 2, 3


def source(w, h, fmt, seed):
    rng = np.random.default_rng(seed)
    if fmt == BGRA:
        return [rng.integers(0, 256, size=(h, 4 * w), dtype=np.uint8)]
    cw, ch = (w + 1) // 2, (h if fmt == YUV422P else (h + 1) // 2)
    Y = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    U = rng.integers(0, 256, size=(ch, cw), dtype=np.uint8)
    V = rng.integers(0, 256, size=(ch, cw), dtype=np.uint8)
    if fmt == NV12:
        return [Y, np.ascontiguousarray(np.stack([U, V], axis=2).reshape(ch, 2 * cw))]
    return [Y, U, V]


@pytest.mark.parametrize("sw,sh,dw,dh", [(64, 48, 64, 48), (64, 48, 160, 120), (720, 480, 1920, 1080), (1920, 1080, 720, 480),
                                         (101, 67, 33, 200), (352, 288, 720, 576), (33, 21, 7, 5)])
@pytest.mark.parametrize("fmt", [BGRA, YUV420P, YUV422P, NV12])
def test_scaler_matches_the_oracle(sw, sh, dw, dh, fmt):
    import torch
    planes = source(sw, sh, fmt, sw * 7 + dw + fmt)
    want = helpers.oracle_scale_to_bgra(planes, sw, sh, fmt, dw, dh)
    dev = [torch.from_numpy(p).cuda() for p in planes]
    dst = torch.zeros((dh, dw), dtype=torch.int32, device="cuda")
    with cvs.Engine([], max_w=dw, max_h=dh, max_batch=1) as eng:
        eng.scale_to_bgra_device(dst, dw, dh, dev, [p.shape[1] for p in planes], sw, sh, fmt)
        eng.synchronize()
    assert np.array_equal(dst.cpu().numpy().view(np.uint32), want)


def _scale_on_gpu(planes, sw, sh, fmt, dw, dh):
    import torch
    dev = [torch.from_numpy(np.ascontiguousarray(p)).cuda() for p in planes]
    dst = torch.zeros((dh, dw), dtype=torch.int32, device="cuda")
    with cvs.Engine([], max_w=dw, max_h=dh, max_batch=1) as eng:
        eng.scale_to_bgra_device(dst, dw, dh, dev, [p.shape[1] for p in planes], sw, sh, fmt)
        eng.synchronize()
    return dst.cpu().numpy().view(np.uint32)


FMT_NAME = {YUV420P: "yuv420p", YUV422P: "yuv422p", NV12: "nv12"}


@pytest.mark.skipif(not swscale_ref.available(), reason="no libswscale on this machine (golden fixtures cover it)")
@pytest.mark.parametrize("sw,sh,dw,dh", [(720, 480, 720, 480), (720, 481, 720, 481), (640, 480, 720, 480), (352, 288, 720, 480),
                                         (1920, 1080, 720, 480), (720, 576, 720, 480), (1280, 720, 1920, 1080), (351, 287, 720, 480),
                                         (640, 480, 721, 480), (353, 289, 353, 289), (1920, 1080, 721, 481)])   # odd widths
@pytest.mark.parametrize("fmt", [YUV420P, YUV422P, NV12])
def test_scaler_equals_libswscale_itself(sw, sh, dw, dh, fmt):
    """sws_getContext(sw, sh, fmt, dw, dh, BGRA, SWS_BILINEAR, ...) + sws_scale() of the library on this machine (its C
    code) against the kernel, byte for byte: every tap-count combination of the library's vertical dispatch."""
    planes = source(sw, sh, fmt, sw + dh + fmt)
    want = swscale_ref.scale(planes, FMT_NAME[fmt], sw, sh, "bgra", dw, dh, c_code=True)[0].view(np.uint32).reshape(dh, dw)
    got = _scale_on_gpu(planes, sw, sh, fmt, dw, dh)
    assert np.array_equal(got, want), int((got != want).sum())


@pytest.mark.skipif(not swscale_ref.available(), reason="no libswscale on this machine (golden fixtures cover it)")
@pytest.mark.parametrize("sw,sh,dw,dh", [(640, 480, 720, 480), (352, 288, 720, 480), (1920, 1080, 720, 480), (720, 576, 720, 480),
                                         (720, 480, 720, 482), (720, 480, 359, 240), (100, 67, 50, 40), (101, 67, 51, 40), (64, 48, 81, 61)])
def test_bgra_source_equals_libswscale_itself(sw, sh, dw, dh):
    """BGRA -> BGRA at another size: the library's RGB -> YUV(A) -> RGB route (k_sws_bgra_to_bgra), alpha included."""
    planes = source(sw, sh, BGRA, sw + dh)
    want = swscale_ref.scale(planes, "bgra", sw, sh, "bgra", dw, dh, c_code=True)[0].view(np.uint32).reshape(dh, dw)
    got = _scale_on_gpu(planes, sw, sh, BGRA, dw, dh)
    assert np.array_equal(got, want), int((got != want).sum())


def test_scaler_equals_golden_outputs_of_libswscale():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "swscale_to_bgra.npz"))
    names = sorted({k[:-5] for k in g.files if k.endswith("_bgra")})
    assert len(names) >= 18
    code = {v: k for k, v in FMT_NAME.items()}
    code["bgra"] = BGRA
    for nm in names:
        fmt, s_, d_ = nm.split("_")
        (sw, sh), (dw, dh) = [tuple(int(v) for v in q.split("x")) for q in (s_, d_)]
        planes = [g["%s_p%d" % (nm, i)] for i in range({"nv12": 2, "bgra": 1}.get(fmt, 3))]
        assert np.array_equal(_scale_on_gpu(planes, sw, sh, code[fmt], dw, dh), g[nm + "_bgra"]), nm


def test_scaler_batches_strides_and_errors():
    import torch
    sw, sh, dw, dh, n = 90, 50, 120, 80, 3
    pics = [source(sw, sh, YUV420P, 40 + k) for k in range(n)]
    pad = 6
    dev = []
    for i in range(3):
        arr = np.stack([np.pad(p[i], ((0, 0), (0, pad))) for p in pics])           # padded rows, pictures back to back
        dev.append(torch.from_numpy(arr).cuda())
    dst = torch.full((n, dh, dw + 4), 0x55, dtype=torch.int32, device="cuda")
    with cvs.Engine([], max_w=dw, max_h=dh, max_batch=1) as eng:
        eng.scale_to_bgra_device(dst, dw, dh, dev, [d.shape[2] for d in dev], sw, sh, YUV420P, n=n, dst_stride=4 * (dw + 4),
                                 dst_pic_stride=4 * (dw + 4) * dh, pic_strides=[d.shape[1] * d.shape[2] for d in dev])
        eng.synchronize()
        out = dst.cpu().numpy().view(np.uint32)
        for k in range(n):
            assert np.array_equal(out[k][:, :dw], helpers.oracle_scale_to_bgra(pics[k], sw, sh, YUV420P, dw, dh)), k
            assert (out[k][:, dw:] == 0x55).all()                                     # padding untouched
        with pytest.raises(cvs.CvsError) as e:
            eng.scale_to_bgra_device(dst, dw, dh, dev, [sw - 1, 45, 45], sw, sh, YUV420P)   # luma rows shorter than sw
        assert e.value.status == -1
        with pytest.raises(cvs.CvsError) as e:
            eng.scale_to_bgra_device(dst, 4, 2, dev, [d.shape[2] for d in dev], sw, sh, YUV420P)   # > 16x reduction
        assert e.value.status == -5


def test_scaled_pictures_feed_the_field_loop(oracle):
    """decode -> scale -> composite_layer, all on the device: the scaler's BGRA goes straight into
    cvs_composite_fields_device and the result equals the oracle run on the oracle-scaled picture."""
    import torch
    sw, sh, w, h, n = 176, 144, 320, 240, 2
    p = helpers.params("-vhs")
    srcs = [source(sw, sh, NV12, 70 + k) for k in range(n)]
    scaled = [helpers.oracle_scale_to_bgra(s, sw, sh, NV12, w, h) for s in srcs]
    want, _ = helpers.run_oracle(oracle, p, lambda k: scaled[k], n, w, h)
    Y = torch.from_numpy(np.stack([s[0] for s in srcs])).cuda()
    UV = torch.from_numpy(np.stack([s[1] for s in srcs])).cuda()
    bgra = torch.zeros((n, h, w), dtype=torch.int32, device="cuda")
    out = torch.zeros((n, h, w), dtype=torch.int32, device="cuda")
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=n) as eng:
        eng.set_precision(True)
        eng.scale_to_bgra_device(bgra, w, h, [Y, UV], [Y.shape[2], UV.shape[2]], sw, sh, NV12, n=n)
        eng.composite_fields_device(out, bgra, n, h, w, 0)
        eng.synchronize()
    got = np.zeros((h, w), dtype=np.uint32)
    o = out.cpu().numpy().view(np.uint32)
    got[1::2] = o[0][1::2]
    got[0::2] = o[1][0::2]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("w,h,v420", [(320, 240, True), (164, 121, True), (160, 120, False)])
def test_field_loop_host_equals_the_reference_loop_with_the_oracle_conversions(oracle, w, h, v420):
    """cvs_field_loop_host: decoder pictures (NV12, host) -> scale -> composite_layer -> line doubling -> encoder YUV
    (host), in two calls (the ring row travels in the context), against the same chain made of the oracles: the
    reference's field loop (one ring picture, ffmpeg_ntsc.cpp:2190-2282) with oracle/convert_oracle.c on both ends."""
    import ctypes as C
    sw, sh, nsrc, fpf = 176, 144, 4, 2
    n = nsrc * fpf
    p = helpers.params("-vhs")
    srcs = [source(sw, sh, NV12, 90 + k) for k in range(nsrc)]
    # the oracle chain
    g = helpers.OracleRng()
    oracle.oracle_rng_seed(C.byref(g), 1)
    ring = np.zeros((h, w), dtype=np.uint32)
    want = []
    for cur in range(n):
        bgra = helpers.oracle_scale_to_bgra(srcs[cur // fpf], sw, sh, NV12, w, h)
        f = (cur & 1) ^ 1
        oracle.oracle_composite_layer(C.byref(p), C.byref(g), ring.ctypes.data_as(C.c_void_p), 4 * w,
                                      bgra.ctypes.data_as(C.c_void_p), 4 * w, w, h, 0, 0, f, C.c_ulonglong(cur))
        oracle.oracle_bob(ring.ctypes.data_as(C.c_void_p), 4 * w, w, h, f)
        want.append(helpers.oracle_bgra_to_yuv(ring, v420))
    # the engine, two calls of n/2 fields
    Ys = np.stack([s[0] for s in srcs])
    UVs = np.stack([s[1] for s in srcs])
    cw, ch = (w + 1) // 2, ((h + 1) // 2 if v420 else h)
    y = np.zeros((n, h, w), np.uint8)
    u = np.zeros((n, ch, cw), np.uint8)
    v = np.zeros((n, ch, cw), np.uint8)
    half = n // 2
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=n) as eng:
        eng.set_precision(True)
        for c0 in (0, half):
            sel = slice(c0 // fpf, (c0 + half) // fpf)
            eng.field_loop_host([Ys[sel], UVs[sel]], sw, sh, NV12, w, h, half, c0, y[c0:c0 + half], u[c0:c0 + half], v[c0:c0 + half],
                                fmt420=v420, src_of_field=[k // fpf for k in range(half)])
        assert eng.rng_tell() == g.pos
    for k in range(n):
        for name, got, wnt in (("y", y[k], want[k][0]), ("u", u[k], want[k][1]), ("v", v[k], want[k][2])):
            assert np.array_equal(got, wnt), (k, name)
