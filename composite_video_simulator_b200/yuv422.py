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
