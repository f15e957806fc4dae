"""cvs_bgra_to_yuv_device: BGRA -> planar YUV 4:2:0 / 4:2:2, the reference's sws_scale() call at
ffmpeg_ntsc.cpp:2266-2274.  PINNED: checked bit for bit against oracle/convert_oracle.c (which tests/test_swscale_pin.py
pins against libswscale 9.1.100 itself), against the library directly when the GPU box has it, and against outputs of
the library committed under tests/golden/.  Luma (and 4:2:2 chroma of even widths) also within +-1 of the real-valued
BT.601 conversion; grey stays exactly grey."""
import os

import numpy as np
import pytest

import composite_video_simulator_b200 as cvs
import swscale_ref

pytestmark = pytest.mark.gpu

def formula(bgra, v420):
    """oracle/convert_oracle.c: the conversion restated from its specification, independently of the kernel."""
    import helpers
    return helpers.oracle_bgra_to_yuv(bgra, v420)


def real_bt601(bgra, v420):
    h, w = bgra.shape
    b, g, r = [((bgra >> s) & 0xFF).astype(np.float64) for s in (0, 8, 16)]
    y = 16 + (0.299 * r + 0.587 * g + 0.114 * b) * 219 / 255
    cb = 128 + (-0.168736 * r - 0.331264 * g + 0.5 * b) * 224 / 255
    cr = 128 + (0.5 * r - 0.418688 * g - 0.081312 * b) * 224 / 255
    wp, hp = w + (w & 1), h + ((h & 1) if v420 else 0)
    pad = lambda a: np.pad(a, ((0, hp - h), (0, wp - w)), mode="edge")
    if v420:
        mean = lambda a: (lambda q: (q[0::2, 0::2] + q[0::2, 1::2] + q[1::2, 0::2] + q[1::2, 1::2]) / 4)(pad(a))
    else:
        mean = lambda a: (lambda q: (q[:, 0::2] + q[:, 1::2]) / 2)(pad(a))
    return y, mean(cb), mean(cr)


# (widths that are multiples of 8 take the streaming kernels -- 4:2:2 direct, 4:2:0 tiled over 8 chroma rows: tile edges, odd
#  heights and pictures smaller than a tile are in the list; the others take the general kernel / the odd-width kernel)
@pytest.mark.parametrize("w,h,n", [(64, 32, 1), (720, 480, 3), (1920, 1080, 2), (101, 67, 2), (8, 2, 1), (1366, 768, 1),
                                   (720, 481, 2), (256, 67, 1), (64, 6, 1), (64, 3, 2), (512, 17, 1), (264, 34, 1)])
@pytest.mark.parametrize("v420", [True, False])
def test_bgra_to_yuv_matches_the_formula(w, h, n, v420):
    import torch
    rng = np.random.default_rng(w * 31 + h)
    src = rng.integers(0, 1 << 24, size=(n, h, w), dtype=np.uint32)
    src[0, : min(h, 4)] = np.array([0x000000, 0xFFFFFF, 0xFF0000, 0x00FF00, 0x0000FF, 0x808080, 0xC0C000, 0x00C0C0],
                                   dtype=np.uint32)[np.arange(w) % 8]
    cw, ch = (w + 1) // 2, ((h + 1) // 2 if v420 else h)
    d = torch.from_numpy(src.view(np.int32)).cuda()
    y = torch.zeros((n, h, w), dtype=torch.uint8, device="cuda")
    u = torch.zeros((n, ch, cw), dtype=torch.uint8, device="cuda")
    v = torch.zeros((n, ch, cw), dtype=torch.uint8, device="cuda")
    with cvs.Engine([], max_w=w, max_h=h, max_batch=1) as eng:
        eng.bgra_to_yuv_device(y, u, v, d, w, h, n, fmt420=v420)
        eng.synchronize()
    for k in range(n):
        wy, wu, wv = formula(src[k], v420)
        assert np.array_equal(y[k].cpu().numpy(), wy), (k, "y")
        assert np.array_equal(u[k].cpu().numpy(), wu), (k, "u")
        assert np.array_equal(v[k].cpu().numpy(), wv), (k, "v")
        fy, fu, fv = real_bt601(src[k], v420)
        assert np.abs(wy - fy).max() <= 1.0
        if not v420 and w % 2 == 0:                 # (4:2:0 and odd widths are filtered, not averaged)
            assert np.abs(wu - fu).max() <= 1.0 and np.abs(wv - fv).max() <= 1.0
    # limited range: black -> (16, 128, 128), white -> (235, 128, 128)
    assert int(y[0, 0, 0]) == 16 and int(y[0, 0, 1]) == 235


def test_grey_stays_grey():
    """Every grey level maps to U = V = 128 exactly (the library's chroma rows sum to -1 in 2^15: below its rounding)."""
    import torch
    w, h = 256, 4
    src = np.broadcast_to((np.arange(w, dtype=np.uint32) * 0x010101)[None, :], (h, w)).copy()
    d = torch.from_numpy(src.view(np.int32)).cuda()
    y = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    u = torch.zeros((h // 2, w // 2), dtype=torch.uint8, device="cuda")
    v = torch.zeros_like(u)
    with cvs.Engine([], max_w=w, max_h=h, max_batch=1) as eng:
        eng.bgra_to_yuv_device(y, u, v, d, w, h, 1, fmt420=True)
        eng.synchronize()
    assert (u.cpu().numpy() == 128).all() and (v.cpu().numpy() == 128).all()


def test_bgra_to_yuv_padded_strides_and_errors():
    import torch
    w, h = 100, 50
    src = np.random.default_rng(5).integers(0, 1 << 24, size=(h, w + 12), dtype=np.uint32)
    d = torch.from_numpy(src.view(np.int32)).cuda()
    y = torch.full((h, w + 7), 7, dtype=torch.uint8, device="cuda")
    u = torch.full((h // 2, w // 2 + 3), 7, dtype=torch.uint8, device="cuda")
    v = torch.full((h // 2, w // 2 + 5), 7, dtype=torch.uint8, device="cuda")
    with cvs.Engine([], max_w=w, max_h=h, max_batch=1) as eng:
        eng.bgra_to_yuv_device(y, u, v, d, w, h, 1, fmt420=True, ly=w + 7, lu=w // 2 + 3, lv=w // 2 + 5, stride=4 * (w + 12))
        eng.synchronize()
        wy, wu, wv = formula(src[:, :w], True)
        assert np.array_equal(y.cpu().numpy()[:, :w], wy) and (y.cpu().numpy()[:, w:] == 7).all()
        assert np.array_equal(u.cpu().numpy()[:, : w // 2], wu) and (u.cpu().numpy()[:, w // 2:] == 7).all()
        assert np.array_equal(v.cpu().numpy()[:, : w // 2], wv) and (v.cpu().numpy()[:, w // 2:] == 7).all()
        with pytest.raises(cvs.CvsError):
            eng.bgra_to_yuv_device(y, u, v, d, w, h, 1, fmt420=True, ly=w - 1)      # luma rows shorter than w


def _run(src, v420):
    import torch
    n, h, w = src.shape
    cw, ch = (w + 1) // 2, ((h + 1) // 2 if v420 else h)
    d = torch.from_numpy(src.view(np.int32)).cuda()
    y = torch.zeros((n, h, w), dtype=torch.uint8, device="cuda")
    u = torch.zeros((n, ch, cw), dtype=torch.uint8, device="cuda")
    v = torch.zeros((n, ch, cw), dtype=torch.uint8, device="cuda")
    with cvs.Engine([], max_w=w, max_h=h, max_batch=1) as eng:
        eng.bgra_to_yuv_device(y, u, v, d, w, h, n, fmt420=v420)
        eng.synchronize()
    return y.cpu().numpy(), u.cpu().numpy(), v.cpu().numpy()


def test_bgra_to_yuv_equals_golden_outputs_of_libswscale():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "swscale_bgra_yuv.npz"))
    names = sorted({k[:-4] for k in g.files if k.endswith("_src")})
    assert len(names) >= 8
    for nm in names:
        got = _run(g[nm + "_src"][None], "yuv420p" in nm)
        for pl, a in zip("yuv", got):
            assert np.array_equal(a[0], g["%s_%s" % (nm, pl)]), (nm, pl)


@pytest.mark.skipif(not swscale_ref.available(), reason="no libswscale on this machine (golden fixtures cover it)")
@pytest.mark.parametrize("w,h", [(720, 480), (1920, 1080), (719, 575), (100, 47), (34, 3)])
@pytest.mark.parametrize("fmt", ["yuv420p", "yuv422p"])
def test_bgra_to_yuv_equals_libswscale_itself(w, h, fmt):
    """The kernel against sws_scale() of the library on this machine (its C code; the x86 SIMD code of the same library
    differs from that by +-1 on 4:2:0 chroma only -- asserted here too)."""
    src = np.random.default_rng(w + 7 * h).integers(0, 1 << 32, size=(1, h, w), dtype=np.uint32)
    got = _run(src, fmt == "yuv420p")
    bgra = src[0].view(np.uint8).reshape(h, 4 * w)
    want = swscale_ref.scale([bgra], "bgra", w, h, fmt, w, h, c_code=True)
    for pl, a, b in zip("yuv", got, want):
        assert np.array_equal(a[0], b), (pl, int(np.abs(a[0].astype(int) - b.astype(int)).max()))
    simd = swscale_ref.scale([bgra], "bgra", w, h, fmt, w, h, c_code=False)
    for pl, a, b in zip("yuv", got, simd):
        assert np.abs(a[0].astype(int) - b.astype(int)).max() <= (1 if (fmt == "yuv420p" and pl != "y") else 0), pl
