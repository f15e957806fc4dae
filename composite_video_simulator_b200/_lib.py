"""Loader of the product library ``libcvs_ntsc.so`` (C ABI in include/cvs_ntsc.h).

The library is built in-tree by ``composite_video_simulator_b200.build`` (nvcc, sm_100a).
There is no Python or CPU fallback: if the library is missing, loading fails loudly.
"""
import ctypes as C
import os

from .params import CvsParams

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CVS_NTSC_LIB") or os.path.join(_HERE, "libcvs_ntsc.so")   # override: experiments only

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libcvs_ntsc.so is not built (%s). Run `python -m composite_video_simulator_b200.build` "
            "(needs nvcc); there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER(CvsParams)
    vp, u8p = C.c_void_p, C.c_void_p
    sig = {
        "cvs_abi_version": (C.c_int, []),
        "cvs_strerror": (C.c_char_p, [C.c_int]),
        "cvs_params_default_ntsc": (C.c_int, [P]),
        "cvs_params_preset_pal": (C.c_int, [P]),
        "cvs_params_apply_argv": (C.c_int, [P, C.c_int, C.POINTER(C.c_char_p)]),
        "cvs_draws_per_field": (C.c_ulonglong, [P, C.c_int, C.c_int, C.c_uint]),
        "cvs_create": (C.c_int, [C.POINTER(vp), P, C.c_int, C.c_int, C.c_int, C.c_int]),
        "cvs_destroy": (None, [vp]),
        "cvs_set_params": (C.c_int, [vp, P]),
        "cvs_set_precision": (C.c_int, [vp, C.c_int]),
        "cvs_set_bob": (C.c_int, [vp, C.c_int]),
        "cvs_set_noise_mode": (C.c_int, [vp, C.c_int]),
        "cvs_bgra_to_yuv_device": (C.c_int, [vp, vp, C.c_int, C.c_longlong, vp, C.c_int, C.c_longlong, vp, C.c_int,
                                             C.c_longlong, vp, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int]),
        "cvs_sws_bilinear_bank": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]),
        "cvs_scale_to_bgra_device": (C.c_int, [vp, vp, C.c_int, C.c_longlong, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_int),
                                               C.POINTER(C.c_longlong), C.c_int, C.c_int, C.c_int, C.c_int]),
        "cvs_field_loop_host": (C.c_int, [vp, vp, C.c_int, C.c_ulonglong]),
        "cvs_audio_channels": (C.c_int, [P]),
        "cvs_audio_create": (C.c_int, [C.POINTER(vp), P]),
        "cvs_audio_process": (C.c_int, [vp, C.c_void_p, C.c_uint, C.POINTER(C.c_ulonglong)]),
        "cvs_audio_destroy": (None, [vp]),
        "cvs_preferred_batch": (C.c_int, [vp, C.c_int, C.c_int, C.c_int]),
        "cvs_composite_layer": (C.c_int, [vp, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_uint, C.c_ulonglong]),
        "cvs_composite_fields_device": (C.c_int, [vp, vp, C.c_size_t, C.c_int, vp, C.c_size_t, C.c_int, C.c_int,
                                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong]),
        "cvs_composite_fields_host": (C.c_int, [vp, vp, C.c_size_t, C.c_int, vp, C.c_size_t, C.c_int, C.c_int,
                                                C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong]),
        "cvs_composite_fields_host_async": (C.c_int, [vp, vp, C.c_size_t, C.c_int, vp, C.c_size_t, C.c_int, C.c_int,
                                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong]),
        "cvs_synchronize": (C.c_int, [vp]),
        "cvs_alloc_host": (C.c_int, [vp, C.POINTER(vp), C.c_size_t]),
        "cvs_free_host": (C.c_int, [vp, vp]),
        "cvs_rng_seek": (C.c_int, [vp, C.c_ulonglong]),
        "cvs_rng_tell": (C.c_int, [vp, C.POINTER(C.c_ulonglong)]),
        "cvs_kernel_launches": (C.c_ulonglong, [vp]),
        "cvs_kernel_time_reset": (C.c_int, [vp]),
        "cvs_kernel_time_query": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
        "cvs_set_stream": (C.c_int, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "cvs_abi_version", "cvs_strerror", "cvs_params_default_ntsc", "cvs_params_preset_pal",
    "cvs_params_apply_argv", "cvs_draws_per_field", "cvs_create", "cvs_destroy", "cvs_set_params",
    "cvs_set_precision", "cvs_set_bob", "cvs_set_noise_mode", "cvs_scale_to_bgra_device", "cvs_field_loop_host", "cvs_audio_channels", "cvs_audio_create", "cvs_audio_process",
    "cvs_audio_destroy", "cvs_bgra_to_yuv_device", "cvs_sws_bilinear_bank", "cvs_preferred_batch", "cvs_composite_layer", "cvs_composite_fields_device", "cvs_composite_fields_host",
    "cvs_composite_fields_host_async",
    "cvs_synchronize", "cvs_alloc_host", "cvs_free_host", "cvs_rng_seek", "cvs_rng_tell", "cvs_kernel_launches", "cvs_kernel_time_reset",
    "cvs_kernel_time_query", "cvs_set_stream",
]
