"""Host-side mirror of the reference's interface for the composite_layer() path.

The reference exposes no API; its seam is the call
``composite_layer(dst, src, inputfile, field, fieldno)`` (ffmpeg_ntsc.cpp:2229) configured by CLI
switches (parse_argv, ffmpeg_ntsc.cpp:972-1282).  ``Engine`` mirrors exactly that: it is built from
an argv list with the reference's switch names, ``composite_layer`` takes the same arguments (BGRA
pictures as ``uint32[h, w]`` arrays instead of AVFrames) and writes only the rows of the requested
field, and the hidden libc ``rand()`` position is an explicit ``rng_seek``/``rng_tell``.

Everything is a thin ctypes call into libcvs_ntsc.so (the CUDA engine); nothing is computed here.
"""
import ctypes as C

import numpy as np

from . import _lib
from .params import CvsParams


class CvsError(RuntimeError):
    def __init__(self, status, what=""):
        self.status = status
        msg = _lib.load().cvs_strerror(status).decode()
        super().__init__("%s%s (status %d)" % (what + ": " if what else "", msg, status))


def _check(status, what=""):
    if status != 0:
        raise CvsError(status, what)


def default_params():
    """preset_NTSC() + the global initialisers of ffmpeg_ntsc.cpp."""
    p = CvsParams()
    _check(_lib.load().cvs_params_default_ntsc(C.byref(p)), "cvs_params_default_ntsc")
    return p


def params_from_argv(argv, base=None):
    """Apply reference CLI switches (e.g. ``["-vhs", "-vhs-speed", "ep"]``) in order."""
    p = base.copy() if base is not None else default_params()
    args = [b"ffmpeg_ntsc"] + [a.encode() if isinstance(a, str) else a for a in argv]
    arr = (C.c_char_p * len(args))(*args)
    _check(_lib.load().cvs_params_apply_argv(C.byref(p), len(args), arr), "cvs_params_apply_argv")
    return p


def draws_per_field(params, w, h, field):
    return int(_lib.load().cvs_draws_per_field(C.byref(params), w, h, field))


def _ptr(a):
    """Raw address of a numpy array, a torch tensor or an int."""
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError("expected a numpy array, a torch tensor or an address")


def _host_pictures(dst, src, ndim, dst_stride=None, src_stride=None):
    """Validate a pair of HOST picture arrays before their addresses cross the C ABI: the library trusts the
    geometry it is given, so a short, narrow or non-contiguous array would let its copies run outside the buffer."""
    for name, a in (("dst", dst), ("src", src)):
        if not isinstance(a, np.ndarray):
            raise TypeError("%s must be a numpy array in host memory (use composite_fields_device for device "
                            "pictures), got %s" % (name, type(a).__name__))
        if a.ndim != ndim or a.dtype.itemsize != 4 or a.dtype.kind not in "ui":
            raise ValueError("%s must be a %d-dimensional array of 32-bit BGRA pixels, got %s %s"
                             % (name, ndim, a.dtype, a.shape))
        if a.strides[-1] != 4 or any(st <= 0 for st in a.strides):
            raise ValueError("%s rows must be contiguous (pixel stride 4 bytes), got strides %s" % (name, a.strides))
    if not dst.flags.writeable:
        raise ValueError("dst is read-only")
    if dst_stride is None and src_stride is None and dst.shape != src.shape:
        raise ValueError("dst %s and src %s must have the same shape" % (dst.shape, src.shape))


class Audio:
    """composite_audio_process() of ffmpeg_ntsc (ffmpeg_ntsc.cpp:901-970) on the CPU: cvs_audio_* of the C ABI.
    `process(pcm, rng_pos)` filters interleaved int16 samples [n, channels] in place and returns the rand() stream
    position after the tape-hiss draws, to be handed back to the video engine (Engine.rng_seek)."""

    def __init__(self, argv=(), params=None):
        self.lib = _lib.load()
        self.params = params.copy() if params is not None else params_from_argv(list(argv))
        self._a = C.c_void_p()
        _check(self.lib.cvs_audio_create(C.byref(self._a), C.byref(self.params)), "cvs_audio_create")
        self.channels = self.lib.cvs_audio_channels(C.byref(self.params))

    def process(self, pcm, rng_pos=0):
        import numpy as np
        assert pcm.dtype == np.int16 and pcm.flags["C_CONTIGUOUS"] and pcm.size % self.channels == 0
        pos = C.c_ulonglong(rng_pos)
        _check(self.lib.cvs_audio_process(self._a, pcm.ctypes.data_as(C.c_void_p), pcm.size // self.channels,
                                          C.byref(pos)), "cvs_audio_process")
        return pos.value

    def close(self):
        if getattr(self, "_a", None) is not None and self._a:
            self.lib.cvs_audio_destroy(self._a)
            self._a = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class FieldLoopDesc(C.Structure):
    """cvs_field_loop of include/cvs_ntsc.h."""
    _fields_ = [("struct_size", C.c_int32), ("src_format", C.c_int32), ("src_w", C.c_int32), ("src_h", C.c_int32),
                ("src", C.c_void_p * 3), ("src_linesize", C.c_int32 * 3), ("nsrc", C.c_int32),
                ("src_pic_stride", C.c_longlong * 3), ("src_of_field", C.POINTER(C.c_int32)),
                ("w", C.c_int32), ("h", C.c_int32), ("out_format", C.c_int32), ("pad_", C.c_int32),
                ("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
                ("ly", C.c_int32), ("lu", C.c_int32), ("lv", C.c_int32), ("pad2_", C.c_int32),
                ("y_pic_stride", C.c_longlong), ("u_pic_stride", C.c_longlong), ("v_pic_stride", C.c_longlong)]


class Engine:
    """One CUDA scanline engine (context of the C ABI) bound to one device."""

    def __init__(self, argv=(), params=None, device=0, max_w=1920, max_h=1080, max_batch=64):
        self.lib = _lib.load()
        self.params = params.copy() if params is not None else params_from_argv(list(argv))
        self._ctx = C.c_void_p()
        _check(self.lib.cvs_create(C.byref(self._ctx), C.byref(self.params), device, max_w, max_h, max_batch),
               "cvs_create")
        self.device = device

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self.lib.cvs_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- configuration -------------------------------------------------------------------------
    def set_params(self, params):
        _check(self.lib.cvs_set_params(self._ctx, C.byref(params)), "cvs_set_params")
        self.params = params.copy()

    def set_precision(self, use_double):
        _check(self.lib.cvs_set_precision(self._ctx, 1 if use_double else 0), "cvs_set_precision")

    def set_bob(self, enable):
        """Fuse the reference's line doubling (ffmpeg_ntsc.cpp:2232-2257) into the field call."""
        _check(self.lib.cvs_set_bob(self._ctx, 1 if enable else 0), "cvs_set_bob")

    def set_noise_mode(self, fast):
        """CVS_NOISE_FAST: per-pixel noise from counter generators (+-1 LSB); default exact rand() replay."""
        _check(self.lib.cvs_set_noise_mode(self._ctx, 1 if fast else 0), "cvs_set_noise_mode")

    def preferred_batch(self, w, h, max_batch):
        """Largest batch <= max_batch that fills whole waves of the device (see cvs_preferred_batch)."""
        n = self.lib.cvs_preferred_batch(self._ctx, w, h, max_batch)
        if n <= 0:
            raise CvsError(n, "cvs_preferred_batch")
        return n

    def rng_seek(self, draws_consumed):
        _check(self.lib.cvs_rng_seek(self._ctx, draws_consumed), "cvs_rng_seek")

    def rng_tell(self):
        v = C.c_ulonglong()
        _check(self.lib.cvs_rng_tell(self._ctx, C.byref(v)), "cvs_rng_tell")
        return v.value

    # -- the seam ---------------------------------------------------------------------------------
    def composite_layer(self, dst, src, field, fieldno, src_interlaced=False, src_top_field_first=False,
                        dst_stride=None, src_stride=None):
        """ffmpeg_ntsc.cpp:2229: dst/src are uint32[h, w] BGRA host pictures; writes rows y = field (mod 2)."""
        _host_pictures(dst, src, 2, dst_stride, src_stride)
        h, w = src.shape[:2]
        _check(self.lib.cvs_composite_layer(self._ctx, _ptr(dst), dst_stride or dst.strides[0], _ptr(src),
                                            src_stride or src.strides[0], w, h, int(src_interlaced),
                                            int(src_top_field_first), field, fieldno), "cvs_composite_layer")

    # -- throughput forms -----------------------------------------------------------------------------
    def composite_fields_host(self, dst, src, first_fieldno, src_interlaced=False, src_top_field_first=False):
        """dst/src: uint32[n, h, w] host arrays (pinned for full PCIe speed)."""
        _host_pictures(dst, src, 3)
        n, h, w = src.shape
        _check(self.lib.cvs_composite_fields_host(self._ctx, _ptr(dst), dst.strides[0], dst.strides[1], _ptr(src),
                                                  src.strides[0], src.strides[1], w, h, int(src_interlaced),
                                                  int(src_top_field_first), n, first_fieldno),
               "cvs_composite_fields_host")

    def composite_fields_host_async(self, dst, src, first_fieldno, src_interlaced=False, src_top_field_first=False):
        """As composite_fields_host but only queues the work; dst is complete after synchronize().
        Consecutive calls overlap; do not reuse dst/src before the synchronize (use two buffer sets)."""
        _host_pictures(dst, src, 3)
        n, h, w = src.shape
        _check(self.lib.cvs_composite_fields_host_async(self._ctx, _ptr(dst), dst.strides[0], dst.strides[1], _ptr(src),
                                                        src.strides[0], src.strides[1], w, h, int(src_interlaced),
                                                        int(src_top_field_first), n, first_fieldno),
               "cvs_composite_fields_host_async")

    def composite_fields_device(self, dst, src, n, h, w, first_fieldno, dst_pic_stride=None, dst_stride=None,
                                src_pic_stride=None, src_stride=None, src_interlaced=False,
                                src_top_field_first=False):
        """dst/src: device addresses (or torch CUDA tensors) of n packed BGRA pictures; asynchronous."""
        ss = src_stride or 4 * w
        ds = dst_stride or 4 * w
        _check(self.lib.cvs_composite_fields_device(self._ctx, _ptr(dst), dst_pic_stride or ds * h, ds, _ptr(src),
                                                    src_pic_stride or ss * h, ss, w, h, int(src_interlaced),
                                                    int(src_top_field_first), n, first_fieldno),
               "cvs_composite_fields_device")

    def bgra_to_yuv_device(self, y, u, v, bgra, w, h, n=1, fmt420=True, ly=None, lu=None, lv=None, stride=None):
        """BGRA -> planar YUV 4:2:0 / 4:2:2 (BT.601 limited range) for n packed pictures on the device; asynchronous.
        Planes are tightly packed per picture unless strides are given (cvs_bgra_to_yuv_device)."""
        cw, ch = (w + 1) // 2, ((h + 1) // 2 if fmt420 else h)
        ly, lu, lv, stride = ly or w, lu or cw, lv or cw, stride or 4 * w
        _check(self.lib.cvs_bgra_to_yuv_device(self._ctx, _ptr(y), ly, ly * h, _ptr(u), lu, lu * ch, _ptr(v), lv, lv * ch,
                                               _ptr(bgra), stride, stride * h, w, h, n, 0 if fmt420 else 1),
               "cvs_bgra_to_yuv_device")

    def scale_to_bgra_device(self, dst, dw, dh, planes, linesizes, sw, sh, fmt, n=1, dst_stride=None, dst_pic_stride=None,
                             pic_strides=None):
        """Decoder pictures -> BGRA at dw x dh on the device (frame_copy_scale, ffmpeg_ntsc.cpp:544-613); asynchronous.
        planes: device addresses / CUDA tensors of the source planes (1 for BGRA, 2 for NV12, 3 for planar YUV);
        fmt: 0 BGRA, 1 YUV420P, 2 YUV422P, 3 NV12 (CVS_PIX_*)."""
        ch = sh if fmt == 2 else (sh + 1) // 2
        rows = [sh, ch, ch]
        ptrs = (C.c_void_p * 3)(*([_ptr(p) for p in planes] + [None] * (3 - len(planes))))
        ls = (C.c_int * 3)(*(list(linesizes) + [0] * (3 - len(linesizes))))
        ps = pic_strides or [linesizes[i] * rows[i] for i in range(len(planes))]
        pst = (C.c_longlong * 3)(*(list(ps) + [0] * (3 - len(ps))))
        ds = dst_stride or 4 * dw
        _check(self.lib.cvs_scale_to_bgra_device(self._ctx, _ptr(dst), ds, dst_pic_stride or ds * dh, dw, dh, ptrs, ls, pst,
                                                 sw, sh, fmt, n), "cvs_scale_to_bgra_device")

    def field_loop_host(self, planes, src_w, src_h, src_fmt, w, h, n, first_fieldno, y, u, v, fmt420=True, src_of_field=None):
        """The reference's whole field loop on the device (cvs_field_loop_host): `planes` = host arrays [nsrc, rows,
        linesize] of the decoder pictures (CVS_PIX_* src_fmt), y/u/v = host arrays [n, rows, linesize] that receive the
        n line-doubled output pictures as planar YUV."""
        d = FieldLoopDesc()
        d.struct_size = C.sizeof(FieldLoopDesc)
        d.src_format, d.src_w, d.src_h, d.nsrc = src_fmt, src_w, src_h, planes[0].shape[0]
        for i, p in enumerate(planes):
            d.src[i] = p.ctypes.data
            d.src_linesize[i] = p.strides[1]
            d.src_pic_stride[i] = p.strides[0]
        idx = None
        if src_of_field is not None:
            idx = (C.c_int32 * n)(*[int(i) for i in src_of_field])
            d.src_of_field = C.cast(idx, C.POINTER(C.c_int32))
        d.w, d.h, d.out_format = w, h, (0 if fmt420 else 1)
        d.y, d.u, d.v = y.ctypes.data, u.ctypes.data, v.ctypes.data
        d.ly, d.lu, d.lv = y.strides[1], u.strides[1], v.strides[1]
        d.y_pic_stride, d.u_pic_stride, d.v_pic_stride = y.strides[0], u.strides[0], v.strides[0]
        _check(self.lib.cvs_field_loop_host(self._ctx, C.byref(d), n, first_fieldno), "cvs_field_loop_host")

    def synchronize(self):
        _check(self.lib.cvs_synchronize(self._ctx), "cvs_synchronize")

    def kernel_launches(self):
        return int(self.lib.cvs_kernel_launches(self._ctx))

    def kernel_time_reset(self):
        _check(self.lib.cvs_kernel_time_reset(self._ctx), "cvs_kernel_time_reset")

    def kernel_time_query(self):
        """(sum of k_fields device durations in ms, number of launches) since the last reset."""
        ms, n = C.c_double(), C.c_int()
        _check(self.lib.cvs_kernel_time_query(self._ctx, C.byref(ms), C.byref(n)), "cvs_kernel_time_query")
        return ms.value, n.value

    def set_stream(self, cuda_stream):
        """Run on a caller-owned stream (raw cudaStream_t handle, e.g. torch's ``stream.cuda_stream``)."""
        _check(self.lib.cvs_set_stream(self._ctx, cuda_stream), "cvs_set_stream")
