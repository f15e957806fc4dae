"""ctypes binding of a REAL libswscale, when the machine has one (TEST INFRASTRUCTURE).

The reference's two picture conversions are calls into libswscale (ffmpeg_ntsc.cpp:574-585 / 603-610 and
:2118-2131 / 2266-2274).  This image carries libswscale 9.1.100 (FFmpeg 8.0) inside opencv-python-headless
(site-packages/opencv_python_headless.libs/); importing cv2 loads it with its dependencies, after which the same
file can be opened with ctypes.  `available()` is False on machines without it, and the tests that need it skip --
tests/golden/swscale_*.npz carry its outputs there (tests/golden/make_golden_swscale.py).

Only sws_getContext / sws_scale / sws_freeContext and libavutil's av_force_cpu_flags are bound: the calls the
reference makes, plus the switch that selects the library's portable C code (what SWS_ACCURATE_RND | SWS_BITEXACT
mean), which is the definition the oracle (oracle/convert_oracle.c) and the kernels are pinned to.
"""
import ctypes as C
import glob
import os
import sysconfig

import numpy as np

PIX = {"yuv420p": 0, "yuv422p": 4, "nv12": 23, "bgra": 28}          # AVPixelFormat (libavutil/pixfmt.h; ABI-stable values)
SWS_BILINEAR = 2
SWS_ACCURATE_RND = 0x40000
SWS_BITEXACT = 0x80000

_state = {}


def _load():
    if "sws" in _state:
        return _state["sws"]
    _state["sws"] = None
    roots = {sysconfig.get_paths()["purelib"], sysconfig.get_paths()["platlib"]}
    for root in roots:
        libs = glob.glob(os.path.join(root, "opencv_python_headless.libs")) + glob.glob(os.path.join(root, "opencv_python.libs"))
        for d in libs:
            sw = glob.glob(os.path.join(d, "libswscale-*.so*"))
            au = glob.glob(os.path.join(d, "libavutil-*.so*"))
            if not sw or not au:
                continue
            try:
                import cv2  # noqa: F401  (loads the bundled FFmpeg libraries and what they depend on)
                avutil = C.CDLL(au[0], mode=C.RTLD_GLOBAL)
                sws = C.CDLL(sw[0])
            except (ImportError, OSError):
                continue
            sws.sws_getContext.restype = C.c_void_p
            sws.sws_getContext.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            sws.sws_scale.restype = C.c_int
            sws.sws_scale.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
            sws.sws_freeContext.argtypes = [C.c_void_p]
            sws.swscale_version.restype = C.c_uint
            avutil.av_force_cpu_flags.argtypes = [C.c_int]
            _state["sws"], _state["avutil"], _state["path"] = sws, avutil, sw[0]
            return sws
    return None


def available():
    return _load() is not None


def version():
    v = _load().swscale_version()
    return "%d.%d.%d" % (v >> 16, (v >> 8) & 255, v & 255)


def use_c_code(on=True):
    """av_force_cpu_flags(0): the library's portable C code; -1 restores the CPU's extensions.  Takes effect for
    contexts created afterwards."""
    _load()
    _state["avutil"].av_force_cpu_flags(0 if on else -1)


def plane_shapes(fmt, w, h):
    cw, ch = (w + 1) // 2, (h + 1) // 2
    return {"bgra": [(h, 4 * w)], "yuv420p": [(h, w), (ch, cw), (ch, cw)], "yuv422p": [(h, w), (h, cw), (h, cw)],
            "nv12": [(h, w), (ch, 2 * cw)]}[fmt]


def _padded(shapes):
    # rows of 64-byte multiples plus slack, as av_frame_get_buffer(frame, 64) gives the reference (:2105, :557)
    return [np.zeros((r + 2, (c + 63) // 64 * 64 + 64), np.uint8) for r, c in shapes]


def scale(src_planes, src_fmt, sw, sh, dst_fmt, dw, dh, flags=SWS_BILINEAR, c_code=True):
    """The reference's call: sws_getContext(sw, sh, src_fmt, dw, dh, dst_fmt, flags, NULL, NULL, NULL) + sws_scale() over
    the whole picture.  src_planes: tightly shaped uint8 arrays per plane; returns the destination planes likewise."""
    sws = _load()
    use_c_code(c_code)
    ctx = sws.sws_getContext(sw, sh, PIX[src_fmt], dw, dh, PIX[dst_fmt], flags, None, None, None)
    assert ctx, "sws_getContext failed"
    try:
        sshape, dshape = plane_shapes(src_fmt, sw, sh), plane_shapes(dst_fmt, dw, dh)
        sbuf, dbuf = _padded(sshape), _padded(dshape)
        for b, p, (r, c) in zip(sbuf, src_planes, sshape):
            assert p.shape == (r, c) and p.dtype == np.uint8, (p.shape, (r, c))
            b[:r, :c] = p
        sp = (C.c_void_p * 4)(*([b.ctypes.data for b in sbuf] + [None] * (4 - len(sbuf))))
        ss = (C.c_int * 4)(*([b.strides[0] for b in sbuf] + [0] * (4 - len(sbuf))))
        dp = (C.c_void_p * 4)(*([b.ctypes.data for b in dbuf] + [None] * (4 - len(dbuf))))
        ds = (C.c_int * 4)(*([b.strides[0] for b in dbuf] + [0] * (4 - len(dbuf))))
        rows = sws.sws_scale(ctx, sp, ss, 0, sh, dp, ds)
        assert rows == dh, rows
        return [b[:r, :c].copy() for b, (r, c) in zip(dbuf, dshape)]
    finally:
        sws.sws_freeContext(ctx)
        use_c_code(False)
