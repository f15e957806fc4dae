"""ctypes mirror of ``cvs_params`` (include/cvs_ntsc.h).

Field names are the reference's global names (ffmpeg_ntsc.cpp:205-214, 756-809); see the
header for the defining line and CLI switch of each one.
"""
import ctypes as C

VHS_SP, VHS_LP, VHS_EP = 0, 1, 2


class CvsParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("output_ntsc", C.c_int32),
        ("output_width", C.c_int32),
        ("output_height", C.c_int32),
        ("video_scanline_phase_shift", C.c_int32),
        ("video_scanline_phase_shift_offset", C.c_int32),
        ("composite_in_chroma_lowpass", C.c_int32),
        ("composite_out_chroma_lowpass", C.c_int32),
        ("composite_out_chroma_lowpass_lite", C.c_int32),
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
        ("reserved0", C.c_int32),
        ("composite_preemphasis", C.c_double),
        ("composite_preemphasis_cut", C.c_double),
        ("vhs_out_sharpen", C.c_double),
        ("vhs_head_switching_point", C.c_double),
        ("vhs_head_switching_phase", C.c_double),
        ("vhs_head_switching_phase_noise", C.c_double),
        ("use_422_colorspace", C.c_int32),
        ("output_frame_delay", C.c_int32),
        ("enable_composite_emulation", C.c_int32),
        ("enable_audio_emulation", C.c_int32),
        ("emulating_preemphasis", C.c_int32),
        ("emulating_deemphasis", C.c_int32),
        ("output_vhs_hifi", C.c_int32),
        ("output_vhs_linear_audio", C.c_int32),
        ("nocolor_subcarrier_after_yc_sep", C.c_int32),
        ("video_yc_recombine", C.c_int32),
        ("output_audio_hiss_db", C.c_double),
        ("output_audio_linear_buzz", C.c_double),
        ("vhs_linear_high_boost", C.c_double),
    ]

    def copy(self):
        q = CvsParams()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(CvsParams))
        return q

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}
