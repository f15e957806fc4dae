"""Python mirror of include/cvs_yuv422.h: the 4:2:2 sibling path (composite_video_process() and
render_field() of ffmpeg_to_composite.cpp) on the B200 engine.

Same rules as the BGRA path: this is a thin ctypes binding of the C ABI, nothing is computed in
Python, and there is no CPU fallback."""
import ctypes as C

import numpy as np

from . import _lib


class Yuv422Params(C.Structure):
    """``cvs422_params``: field names are the globals of ffmpeg_to_composite.cpp (:266-341)."""
    _fields_ = [
        ("output_ntsc", C.c_int32),
        ("output_width", C.c_int32),
        ("output_height", C.c_int32),
        ("video_scanline_phase_shift", C.c_int32),
        ("video_scanline_phase_shift_offset", C.c_int32),
        ("composite_in_chroma_lowpass", C.c_int32),
        ("composite_out_chroma_lowpass", C.c_int32),
        ("composite_out_chroma_lowpass_lite", C.c_int32),
        ("video_yc_recombine", C.c_int32),
        ("video_noise", C.c_int32),
        ("video_chroma_noise", C.c_int32),
        ("video_chroma_phase_noise", C.c_int32),
        ("video_chroma_loss", C.c_int32),
        ("subcarrier_amplitude", C.c_int32),
        ("subcarrier_amplitude_back", C.c_int32),
        ("emulating_vhs", C.c_int32),
        ("output_vhs_tape_speed", C.c_int32),
        ("vhs_head_switching", C.c_int32),
        ("vhs_chroma_vert_blend", C.c_int32),
        ("vhs_svideo_out", C.c_int32),
        ("nocolor_subcarrier", C.c_int32),
        ("nocolor_subcarrier_after_yc_sep", C.c_int32),
        ("enable_composite_emulation", C.c_int32),
        ("reserved0", C.c_int32),
        ("composite_preemphasis", C.c_double),
        ("composite_preemphasis_cut", C.c_double),
        ("vhs_out_sharpen", C.c_double),
        ("vhs_out_sharpen_chroma", C.c_double),
        ("vhs_head_switching_phase", C.c_double),
        ("vhs_head_switching_phase_noise", C.c_double),
    ]

    def copy(self):
        q = Yuv422Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Yuv422Params))
        return q


_P = C.POINTER(Yuv422Params)
_vp = C.c_void_p
_I3 = C.c_int * 3
_V3 = C.c_void_p * 3

SIGNATURES = {
    "cvs422_params_default": (C.c_int, [_P]),
    "cvs422_params_preset_pal": (C.c_int, [_P]),
    "cvs422_params_preset_ntsc": (C.c_int, [_P]),
    "cvs422_params_apply_argv": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p)]),
    "cvs422_create": (C.c_int, [C.POINTER(_vp), _P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cvs422_destroy": (None, [_vp]),
    "cvs422_set_params": (C.c_int, [_vp, _P]),
    "cvs422_composite_video_process": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int,
                                                 C.c_uint, C.c_ulonglong]),
    "cvs422_process_fields_device": (C.c_int, [_vp, _vp, _vp, _vp, C.c_longlong, C.c_longlong, C.c_longlong,
                                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong]),
    "cvs422_process_fields_host": (C.c_int, [_vp, _vp, _vp, _vp, C.c_longlong, C.c_longlong, C.c_longlong,
                                             C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong]),
    "cvs422_render_field_device": (C.c_int, [_vp, _V3, _I3, C.c_int, _V3, _I3, C.c_int, _I3, C.c_int,
                                             C.c_int, C.c_int, C.c_int, C.c_uint]),
    "cvs422_synchronize": (C.c_int, [_vp]),
    "cvs422_set_stream": (C.c_int, [_vp, _vp]),
    "cvs422_rng_seek": (C.c_int, [_vp, C.c_ulonglong]),
    "cvs422_rng_tell": (C.c_ulonglong, [_vp]),
    "cvs422_draws_per_field": (C.c_ulonglong, [_P, C.c_int, C.c_int, C.c_uint]),
    "cvs422_kernel_launches": (C.c_ulonglong, [_vp]),
    "cvs422_kernel_time_reset": (C.c_int, [_vp]),
    "cvs422_kernel_time_query": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
}
EXPORTED_SYMBOLS = list(SIGNATURES)

_bound = {}


def _bind(names):
    L = _lib.load()
    for name in names:
        if name not in _bound:
            fn = getattr(L, name)        # AttributeError if the ABI lost a symbol
            fn.restype, fn.argtypes = SIGNATURES[name]
            _bound[name] = fn
    return L


def lib():
    """The product library with every cvs422_* signature attached."""
    return _bind(SIGNATURES)


class Yuv422Error(RuntimeError):
    def __init__(self, status):
        self.status = status
        RuntimeError.__init__(self, "cvs422: %s (%d)" % (_lib.load().cvs_strerror(status).decode(), status))


def _check(rc):
    if rc != 0:
        raise Yuv422Error(rc)


def params_default():
    p = Yuv422Params()
    _check(_bind(["cvs422_params_default"]).cvs422_params_default(C.byref(p)))
    return p


def params_from_argv(argv):
    """ffmpeg_to_composite switches (without argv[0]) -> parameter block."""
    p = params_default()
    arr = (C.c_char_p * (len(argv) + 1))(b"ffmpeg_to_composite", *[a.encode() for a in argv])
    _check(_bind(["cvs422_params_apply_argv"]).cvs422_params_apply_argv(C.byref(p), len(argv) + 1, arr))
    return p


def _np_ptr(a):
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.ndim >= 2 and a.strides[-1] == 1):
        raise TypeError("planes must be uint8 numpy arrays with contiguous rows")
    return C.c_void_p(a.ctypes.data)


class Yuv422Engine:
    """One context of the 4:2:2 engine: parameters + rand() position + device resources.

    ``Yuv422Engine(["-vhs", "-vhs-speed", "ep"])`` takes ffmpeg_to_composite's switches;
    ``composite_video_process(Y, U, V, field, fieldno)`` is the reference's call
    (ffmpeg_to_composite.cpp:1790) on numpy planes, in place."""

    def __init__(self, argv_or_params=None, device=0, max_w=1920, max_h=1080, max_batch=1):
        if isinstance(argv_or_params, Yuv422Params):
            self.params = argv_or_params.copy()
        else:
            self.params = params_from_argv(list(argv_or_params or []))
        self._lib = lib()
        self._ctx = _vp()
        _check(self._lib.cvs422_create(C.byref(self._ctx), C.byref(self.params), device, max_w, max_h, max_batch))
        self.max_batch = max_batch

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.cvs422_destroy(self._ctx)
            self._ctx = _vp()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, p):
        _check(self._lib.cvs422_set_params(self._ctx, C.byref(p)))
        self.params = p.copy()

    def composite_video_process(self, Y, U, V, w, field, fieldno):
        h = Y.shape[0]
        _check(self._lib.cvs422_composite_video_process(self._ctx, _np_ptr(Y), Y.strides[0], _np_ptr(U), U.strides[0],
                                                        _np_ptr(V), V.strides[0], w, h, field, fieldno))

    def process_fields_host(self, Y, U, V, w, first_fieldno):
        """Y: [n, h, linesize_y], U, V: [n, h, linesize_c] uint8, processed in place."""
        n, h = Y.shape[0], Y.shape[1]
        _check(self._lib.cvs422_process_fields_host(self._ctx, _np_ptr(Y), _np_ptr(U), _np_ptr(V),
                                                    Y.strides[0], U.strides[0], V.strides[0],
                                                    Y.strides[1], U.strides[1], V.strides[1], w, h, n, first_fieldno))

    def process_fields_device(self, Y, U, V, w, first_fieldno):
        """The same on CUDA tensors (torch.uint8, [n, h, linesize]); asynchronous on the context's stream."""
        n, h = Y.shape[0], Y.shape[1]
        _check(self._lib.cvs422_process_fields_device(self._ctx, Y.data_ptr(), U.data_ptr(), V.data_ptr(),
                                                      Y.stride(0), U.stride(0), V.stride(0),
                                                      Y.stride(1), U.stride(1), V.stride(1), w, h, n, first_fieldno))

    def render_field_device(self, dst, src, row_bytes, src_is_420, interlaced, tff, second_field, field):
        """dst, src: three CUDA uint8 tensors [rows, linesize] each."""
        _check(self._lib.cvs422_render_field_device(
            self._ctx, _V3(*[t.data_ptr() for t in dst]), _I3(*[t.stride(0) for t in dst]), dst[0].shape[0],
            _V3(*[t.data_ptr() for t in src]), _I3(*[t.stride(0) for t in src]), src[0].shape[0],
            _I3(*row_bytes), int(src_is_420), int(interlaced), int(tff), int(second_field), field))

    def synchronize(self):
        _check(self._lib.cvs422_synchronize(self._ctx))

    def set_stream(self, cuda_stream):
        _check(self._lib.cvs422_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    def rng_seek(self, draws):
        _check(self._lib.cvs422_rng_seek(self._ctx, draws))

    def rng_tell(self):
        return int(self._lib.cvs422_rng_tell(self._ctx))

    def kernel_launches(self):
        return int(self._lib.cvs422_kernel_launches(self._ctx))

    def kernel_time_reset(self):
        _check(self._lib.cvs422_kernel_time_reset(self._ctx))

    def kernel_time_query(self):
        ms, n = C.c_double(0), C.c_int(0)
        _check(self._lib.cvs422_kernel_time_query(self._ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value
