"""B200-native NTSC/VHS composite-video scanline engine.

Drop-in for one path of joncampbell123/composite-video-simulator: the per-field scanline DSP
``composite_layer()`` of ffmpeg_ntsc.cpp, as hand-written sm_100a CUDA behind the C ABI in
include/cvs_ntsc.h.  See DESIGN.md and INTEGRATION.md.
"""
from .params import CvsParams, VHS_SP, VHS_LP, VHS_EP  # noqa: F401
from .api import Engine, CvsError, default_params, params_from_argv, draws_per_field  # noqa: F401
