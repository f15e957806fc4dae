"""Generates tests/golden/swscale_bgra_yuv.npz and swscale_to_bgra.npz: outputs of the REAL libswscale (the 9.1.100 build of this image,
tests/swscale_ref.py) for the reference's encoder-side call -- sws_getContext(w, h, BGRA, w, h, YUV420P | YUV422P,
SWS_BILINEAR, NULL, NULL, NULL) + sws_scale (ffmpeg_ntsc.cpp:2118-2131, 2266-2274) -- with the library's portable C
code selected (av_force_cpu_flags(0)).  Run from the repo root:  python tests/golden/make_golden_swscale.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import swscale_ref  # noqa: E402

assert swscale_ref.available(), "no libswscale here"
out = {}
rng = np.random.default_rng(20260101)
for w, h in ((64, 48), (101, 67), (100, 47), (33, 21), (160, 120)):
    src = rng.integers(0, 1 << 32, size=(h, w), dtype=np.uint32)
    for fmt in ("yuv420p", "yuv422p"):
        y, u, v = swscale_ref.scale([src.view(np.uint8).reshape(h, 4 * w)], "bgra", w, h, fmt, w, h, c_code=True)
        n = "%s_%dx%d" % (fmt, w, h)
        out[n + "_src"], out[n + "_y"], out[n + "_u"], out[n + "_v"] = src, y, u, v
out["libswscale_version"] = np.array(swscale_ref.version())
np.savez_compressed(os.path.join(HERE, "swscale_bgra_yuv.npz"), **out)
print("wrote", len(out), "arrays, libswscale", swscale_ref.version())

# the input side: sws_getContext(sw, sh, fmt, dw, dh, BGRA, SWS_BILINEAR, ...) + sws_scale (ffmpeg_ntsc.cpp:574-585, 603-610)
out = {}
for fmt in ("yuv420p", "yuv422p", "nv12"):
    for sw, sh, dw, dh in ((64, 48, 64, 48), (64, 47, 64, 47), (40, 30, 64, 48), (176, 144, 64, 48), (90, 72, 64, 48), (33, 21, 80, 60),
                           (40, 30, 65, 48), (64, 48, 63, 48), (176, 144, 65, 47)):       # (odd widths: the full-chroma writers)
        shapes = swscale_ref.plane_shapes(fmt, sw, sh)
        planes = [rng.integers(0, 256, size=s, dtype=np.uint8) for s in shapes]
        bgra = swscale_ref.scale(planes, fmt, sw, sh, "bgra", dw, dh, c_code=True)[0].view(np.uint32).reshape(dh, dw)
        n = "%s_%dx%d_%dx%d" % (fmt, sw, sh, dw, dh)
        for i, pl in enumerate(planes):
            out["%s_p%d" % (n, i)] = pl
        out[n + "_bgra"] = bgra
for sw, sh, dw, dh in ((40, 30, 64, 48), (90, 72, 64, 48), (176, 144, 64, 48), (64, 48, 65, 50), (64, 48, 64, 30)):   # BGRA sources
    img = rng.integers(0, 256, size=(sh, 4 * sw), dtype=np.uint8)
    n = "bgra_%dx%d_%dx%d" % (sw, sh, dw, dh)
    out[n + "_p0"] = img
    out[n + "_bgra"] = swscale_ref.scale([img], "bgra", sw, sh, "bgra", dw, dh, c_code=True)[0].view(np.uint32).reshape(dh, dw)
out["libswscale_version"] = np.array(swscale_ref.version())
np.savez_compressed(os.path.join(HERE, "swscale_to_bgra.npz"), **out)
print("wrote", len(out), "arrays")
