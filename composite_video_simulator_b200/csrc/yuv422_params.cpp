// yuv422_params.cpp -- parameter block and argv front end of the 4:2:2 scanline path.
//
// Host-side mirror of the configuration surface composite_video_process() reads in
// ffmpeg_to_composite.cpp: the file-scope globals (:266-341) become the POD cvs422_params,
// preset_PAL()/preset_NTSC() (:1252-1270) and parse_argv() (:1292-1650) become
// cvs422_params_preset_* / cvs422_params_apply_argv with the same switch names and the same
// order-dependent side effects.  Differences from ffmpeg_ntsc's surface (cvs_params.cpp) that
// matter: -vhs also turns head switching on and sets video_noise = 4; the -comp-catv* cut-off is
// the integer 315000000/88/2; -vhs-head-switching-point sets the (single) head switching phase;
// -yc-recomb is live.
#include "../../include/cvs_yuv422.h"

#include <cstdlib>
#include <cstring>

namespace {

struct Args {
    int argc;
    const char *const *argv;
    int i;
    const char *next() { return (i < argc) ? argv[i++] : nullptr; }   // a missing value is always CVS_ERR_BAD_SWITCH here
};
bool want_int(Args &c, int32_t &out) {
    const char *v = c.next();
    if (!v) return false;
    out = (int32_t)std::atoi(v);
    return true;
}
bool want_flag(Args &c, int32_t &out) {
    int32_t v;
    if (!want_int(c, v)) return false;
    out = v > 0;
    return true;
}
bool want_double(Args &c, double &out) {
    const char *v = c.next();
    if (!v) return false;
    out = std::atof(v);
    return true;
}

void vhs_levels(cvs422_params *p, int phase, int chroma, int loss, int luma) {
    p->video_chroma_phase_noise = phase;
    p->video_chroma_noise = chroma;
    p->video_chroma_loss = loss;
    p->video_noise = luma;
}

}  // namespace

extern "C" {

int cvs422_params_default(cvs422_params *p) {
    if (!p) return CVS_ERR_INVALID_ARG;
    std::memset(p, 0, sizeof(*p));
    p->output_ntsc = 1;
    p->output_width = 720;
    p->output_height = 480;
    p->video_scanline_phase_shift = 180;
    p->video_scanline_phase_shift_offset = 0;
    p->composite_in_chroma_lowpass = 1;
    p->composite_out_chroma_lowpass = 1;
    p->composite_out_chroma_lowpass_lite = 1;
    p->video_yc_recombine = 0;
    p->video_noise = 2;
    p->video_chroma_noise = 0;
    p->video_chroma_phase_noise = 0;
    p->video_chroma_loss = 0;
    p->subcarrier_amplitude = 50;
    p->subcarrier_amplitude_back = 50;
    p->emulating_vhs = 0;
    p->output_vhs_tape_speed = CVS_VHS_SP;
    p->vhs_head_switching = 0;
    p->vhs_chroma_vert_blend = 1;
    p->vhs_svideo_out = 0;
    p->nocolor_subcarrier = 0;
    p->nocolor_subcarrier_after_yc_sep = 0;
    p->enable_composite_emulation = 1;
    p->composite_preemphasis = 0;
    p->composite_preemphasis_cut = 1000000;
    p->vhs_out_sharpen = 1.5;
    p->vhs_out_sharpen_chroma = 0.85;
    p->vhs_head_switching_phase = 1.0 - ((4.5 + 0.01) / 262.5);
    p->vhs_head_switching_phase_noise = (((1.0 / 300)) / 262.5);
    return CVS_OK;
}

int cvs422_params_preset_pal(cvs422_params *p) {
    if (!p) return CVS_ERR_INVALID_ARG;
    p->output_height = 576;
    p->output_width = 720;
    p->output_ntsc = 0;
    return CVS_OK;
}

int cvs422_params_preset_ntsc(cvs422_params *p) {
    if (!p) return CVS_ERR_INVALID_ARG;
    p->output_height = 480;
    p->output_width = 720;
    p->output_ntsc = 1;
    return CVS_OK;
}

int cvs422_params_apply_argv(cvs422_params *p, int argc, const char *const *argv) {
    if (!p || (argc > 0 && !argv)) return CVS_ERR_INVALID_ARG;
    Args c{argc, argv, 1};
    // switches of the reference that do not reach the video path: swallowed with their value
    static const char *const ignored_with_value[] = {"bkey-feedback", "ss", "se", "t", "a", "v", "i", "o",
        "vhs-linear-high-boost", "vhs-linear-video-crosstalk", "audio-hiss", "preemphasis", "deemphasis", "vhs-hifi"};
    static const char *const ignored_bare[] = {"422", "420", "an", "vn", "vi", "vp"};
    while (c.i < c.argc) {
        const char *a = c.next();
        if (!a) return CVS_ERR_BAD_SWITCH;
        if (*a != '-') return CVS_ERR_BAD_SWITCH;                     // "Unhandled arg"
        while (*a == '-') a++;
        int32_t iv;
        if (!std::strcmp(a, "h") || !std::strcmp(a, "help")) return CVS_ERR_HELP;
        else if (!std::strcmp(a, "width")) {
            const char *v = c.next();
            if (!v) return CVS_ERR_BAD_SWITCH;
            p->output_width = (int32_t)std::strtoul(v, nullptr, 0);
            if (p->output_width < 32) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "comp-phase-offset")) {
            if (!want_int(c, p->video_scanline_phase_shift_offset)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "comp-phase")) {
            if (!want_int(c, iv)) return CVS_ERR_BAD_SWITCH;
            p->video_scanline_phase_shift = iv;
            if (!(iv == 0 || iv == 90 || iv == 180 || iv == 270)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "in-composite-lowpass")) {
            if (!want_flag(c, p->composite_in_chroma_lowpass)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "out-composite-lowpass")) {
            if (!want_flag(c, p->composite_out_chroma_lowpass)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "out-composite-lowpass-lite")) {
            if (!want_flag(c, p->composite_out_chroma_lowpass_lite)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "nocomp")) {
            p->enable_composite_emulation = 0;
        } else if (!std::strcmp(a, "vhs-head-switching-point")) {
            if (!want_double(c, p->vhs_head_switching_phase)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "vhs-head-switching-noise-level")) {
            if (!want_double(c, p->vhs_head_switching_phase_noise)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "vhs-head-switching")) {
            if (!want_flag(c, p->vhs_head_switching)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "comp-pre")) {
            if (!want_double(c, p->composite_preemphasis)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "comp-cut")) {
            if (!want_double(c, p->composite_preemphasis_cut)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "comp-catv") || !std::strcmp(a, "comp-catv2") || !std::strcmp(a, "comp-catv3")) {
            const int lvl = a[9] == '\0' ? 1 : (a[9] - '0');
            p->composite_preemphasis = lvl == 1 ? 1.5 : (lvl == 2 ? 2.5 : 4);
            p->composite_preemphasis_cut = 315000000 / 88 / 2;        // integer division, as in the reference
            p->video_chroma_phase_noise = 2 * lvl;
        } else if (!std::strcmp(a, "chroma-phase-noise")) {
            if (!want_int(c, p->video_chroma_phase_noise)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "yc-recomb")) {
            double d;
            if (!want_double(c, d)) return CVS_ERR_BAD_SWITCH;
            p->video_yc_recombine = (int32_t)d;                       // atof() assigned to an int
        } else if (!std::strcmp(a, "vhs-svideo")) {
            if (!want_flag(c, p->vhs_svideo_out)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "vhs-chroma-vblend")) {
            if (!want_flag(c, p->vhs_chroma_vert_blend)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "chroma-noise")) {
            if (!want_int(c, p->video_chroma_noise)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "noise")) {
            if (!want_int(c, p->video_noise)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "subcarrier-amp")) {
            if (!want_int(c, iv)) return CVS_ERR_BAD_SWITCH;
            p->subcarrier_amplitude = iv;
            p->subcarrier_amplitude_back = iv;
        } else if (!std::strcmp(a, "nocolor-subcarrier")) {
            p->nocolor_subcarrier = 1;
        } else if (!std::strcmp(a, "nocolor-subcarrier-after-yc-sep")) {
            p->nocolor_subcarrier_after_yc_sep = 1;
        } else if (!std::strcmp(a, "chroma-dropout")) {
            if (!want_int(c, p->video_chroma_loss)) return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "vhs")) {
            p->emulating_vhs = 1;
            p->vhs_head_switching = 1;
            vhs_levels(p, 4, 16, 4, 4);
        } else if (!std::strcmp(a, "vhs-speed")) {
            const char *v = c.next();
            if (!v) return CVS_ERR_BAD_SWITCH;
            p->emulating_vhs = 1;
            if (!std::strcmp(v, "ep")) { p->output_vhs_tape_speed = CVS_VHS_EP; vhs_levels(p, 6, 22, 8, 6); }
            else if (!std::strcmp(v, "lp")) { p->output_vhs_tape_speed = CVS_VHS_LP; vhs_levels(p, 5, 19, 6, 5); }
            else if (!std::strcmp(v, "sp")) { p->output_vhs_tape_speed = CVS_VHS_SP; vhs_levels(p, 4, 16, 4, 4); }
            else return CVS_ERR_BAD_SWITCH;
        } else if (!std::strcmp(a, "tvstd")) {
            const char *v = c.next();
            if (!v) return CVS_ERR_BAD_SWITCH;
            if (!std::strcmp(v, "pal")) cvs422_params_preset_pal(p);
            else if (!std::strcmp(v, "ntsc")) cvs422_params_preset_ntsc(p);
            else return CVS_ERR_BAD_SWITCH;
        } else {
            bool known = false;
            for (const char *s : ignored_with_value)
                if (!std::strcmp(a, s)) { known = true; if (!c.next()) return CVS_ERR_BAD_SWITCH; break; }
            if (!known)
                for (const char *s : ignored_bare)
                    if (!std::strcmp(a, s)) { known = true; break; }
            if (!known) return CVS_ERR_BAD_SWITCH;
            if (!std::strcmp(a, "vhs-hifi")) p->emulating_vhs = 1;    // "implies -vhs" (the audio half is out of scope)
        }
    }
    // after the loop, once: the pre-emphasis raises the demodulation gain reference
    if (p->composite_preemphasis != 0)
        p->subcarrier_amplitude_back = (int32_t)((double)p->subcarrier_amplitude_back + (50 * p->composite_preemphasis) / 4);   // int += double
    return CVS_OK;
}

}  // extern "C"
