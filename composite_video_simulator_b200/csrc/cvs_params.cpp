// cvs_params.cpp -- the parameter block of the scanline engine and its argv front end.
//
// Host-side mirror of the reference's configuration surface for the composite_layer()
// path: the file-scope globals of ffmpeg_ntsc.cpp:205-214,756-809 become the POD
// cvs_params, preset_NTSC()/preset_PAL() (:815-831) and parse_argv() (:972-1282) become
// cvs_params_default_ntsc / cvs_params_preset_pal / cvs_params_apply_argv with the same
// switch names, the same order-dependent side effects (-vhs resets the noise levels,
// -vhs-speed then overrides them) and the same failure cases.
#include "../../include/cvs_ntsc.h"

#include <cstdlib>
#include <cstring>

namespace {

struct ArgCursor {
    int argc;
    const char *const *argv;
    int i;
    // parse_argv() reads argv[i++] without a bounds check (argv is NULL-terminated); a
    // missing value is a NULL there and a crash or "return 1" in the reference.  Here it is
    // always CVS_ERR_BAD_SWITCH.
    const char *next() { return (i < argc) ? argv[i++] : nullptr; }
};

bool take_int(ArgCursor &c, int32_t &out) {
    const char *v = c.next();
    if (!v) return false;
    out = (int32_t)std::atoi(v);
    return true;
}
bool take_bool(ArgCursor &c, int32_t &out) {   // "atoi(x) > 0"
    int32_t v;
    if (!take_int(c, v)) return false;
    out = v > 0;
    return true;
}
bool take_double(ArgCursor &c, double &out) {
    const char *v = c.next();
    if (!v) return false;
    out = std::atof(v);
    return true;
}

}  // namespace

extern "C" {

int cvs_abi_version(void) { return CVS_ABI_VERSION; }

const char *cvs_strerror(int status) {
    switch (status) {
    case CVS_OK: return "ok";
    case CVS_ERR_INVALID_ARG: return "invalid argument (null pointer, bad geometry or stride < 4*width)";
    case CVS_ERR_BAD_SWITCH: return "unknown or malformed switch";
    case CVS_ERR_CUDA: return "CUDA failure or no CUDA device (there is no CPU fallback)";
    case CVS_ERR_NOMEM: return "out of memory";
    case CVS_ERR_CAPACITY: return "request exceeds the capacity the context was created with";
    case CVS_ERR_HELP: return "help requested";
    case CVS_ERR_NOISE_SYNC: return "noise warm-up did not converge";
    case CVS_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
    }
}

int cvs_params_default_ntsc(cvs_params *p) {
    if (!p) return CVS_ERR_INVALID_ARG;
    std::memset(p, 0, sizeof(*p));
    p->struct_size = (int32_t)sizeof(*p);
    // preset_NTSC(), ffmpeg_ntsc.cpp:824-831
    p->output_ntsc = 1;
    p->output_width = 720;
    p->output_height = 480;
    // global initialisers, :205-214
    p->use_422_colorspace = 0;
    p->video_scanline_phase_shift = 180;
    p->video_scanline_phase_shift_offset = 0;
    // :756-809
    p->composite_preemphasis = 0;
    p->composite_preemphasis_cut = 1000000;
    p->vhs_out_sharpen = 1.5;
    p->vhs_head_switching = 0;
    p->vhs_head_switching_point = 1.0 - ((4.5 + 0.01) / 262.5);
    p->vhs_head_switching_phase = ((1.0 - 0.01) / 262.5);
    p->vhs_head_switching_phase_noise = (((1.0 / 500)) / 262.5);
    p->composite_in_chroma_lowpass = 1;
    p->composite_out_chroma_lowpass = 1;
    p->composite_out_chroma_lowpass_lite = 1;
    p->video_yc_recombine = 0;
    p->video_chroma_noise = 0;
    p->video_chroma_phase_noise = 0;
    p->video_chroma_loss = 0;
    p->video_noise = 2;
    p->subcarrier_amplitude = 50;
    p->subcarrier_amplitude_back = 50;
    p->output_audio_hiss_db = -72;
    p->output_audio_linear_buzz = -42;
    p->vhs_linear_high_boost = 0.25;
    p->output_vhs_hifi = 1;
    p->output_vhs_linear_audio = 0;
    p->emulating_vhs = 0;
    p->emulating_preemphasis = 1;
    p->emulating_deemphasis = 1;
    p->nocolor_subcarrier = 0;
    p->nocolor_subcarrier_after_yc_sep = 0;
    p->vhs_chroma_vert_blend = 1;
    p->vhs_svideo_out = 0;
    p->enable_composite_emulation = 1;
    p->enable_audio_emulation = 1;
    p->output_vhs_tape_speed = CVS_VHS_SP;
    p->output_frame_delay = 1;
    return CVS_OK;
}

int cvs_params_preset_pal(cvs_params *p) {       // preset_PAL(), :815-822
    if (!p) return CVS_ERR_INVALID_ARG;
    p->output_height = 576;
    p->output_width = 720;
    p->output_ntsc = 0;
    return CVS_OK;
}

static void preset_ntsc_geometry(cvs_params *p) {   // preset_NTSC(), :824-831
    p->output_height = 480;
    p->output_width = 720;
    p->output_ntsc = 1;
}

int cvs_params_apply_argv(cvs_params *p, int argc, const char *const *argv) {
    if (!p || (argc > 0 && !argv)) return CVS_ERR_INVALID_ARG;
    if (p->struct_size != (int32_t)sizeof(*p)) return CVS_ERR_INVALID_ARG;
    ArgCursor c{argc, argv, 1};
    while (c.i < c.argc) {
        const char *a = c.argv[c.i++];
        if (!a) return CVS_ERR_BAD_SWITCH;
        if (*a != '-') return CVS_ERR_BAD_SWITCH;            // "Unhandled arg", :1225-1228
        do { a++; } while (*a == '-');                       // :980
        auto is = [&](const char *s) { return std::strcmp(a, s) == 0; };
        bool ok = true;
        int32_t iv;
        if (is("h") || is("help")) return CVS_ERR_HELP;
        else if (is("comp-phase-offset")) ok = take_int(c, p->video_scanline_phase_shift_offset);
        else if (is("comp-phase")) {
            ok = take_int(c, iv);
            if (ok) {
                p->video_scanline_phase_shift = iv;            // assigned before validation, :989
                if (!(iv == 0 || iv == 90 || iv == 180 || iv == 270)) return CVS_ERR_BAD_SWITCH;
            }
        } else if (is("width")) {
            const char *v = c.next();
            if (!v) return CVS_ERR_BAD_SWITCH;
            p->output_width = (int32_t)std::strtoul(v, nullptr, 0);
            if (p->output_width < 32) return CVS_ERR_BAD_SWITCH;
        } else if (is("d")) {
            const char *v = c.next();
            if (!v) return CVS_ERR_BAD_SWITCH;
            unsigned long d = std::strtoul(v, nullptr, 0);
            p->output_frame_delay = (int32_t)d;
            if (d == 0 || d > 256) return CVS_ERR_BAD_SWITCH;
        } else if (is("i") || is("o")) {
            if (!c.next()) return CVS_ERR_BAD_SWITCH;          // media I/O is outside the path
        } else if (is("422")) p->use_422_colorspace = 1;
        else if (is("420")) p->use_422_colorspace = 0;
        else if (is("tvstd")) {
            const char *v = c.next();
            if (!v) return CVS_ERR_BAD_SWITCH;
            if (!std::strcmp(v, "pal")) cvs_params_preset_pal(p);
            else if (!std::strcmp(v, "ntsc")) preset_ntsc_geometry(p);
            else return CVS_ERR_BAD_SWITCH;
        } else if (is("in-composite-lowpass")) ok = take_bool(c, p->composite_in_chroma_lowpass);
        else if (is("out-composite-lowpass")) ok = take_bool(c, p->composite_out_chroma_lowpass);
        else if (is("out-composite-lowpass-lite")) ok = take_bool(c, p->composite_out_chroma_lowpass_lite);
        else if (is("nocomp")) { p->enable_composite_emulation = 0; p->enable_audio_emulation = 0; }
        else if (is("vhs-head-switching-point")) ok = take_double(c, p->vhs_head_switching_point);
        else if (is("vhs-head-switching-phase")) ok = take_double(c, p->vhs_head_switching_phase);
        else if (is("vhs-head-switching-noise-level")) ok = take_double(c, p->vhs_head_switching_phase_noise);
        else if (is("vhs-head-switching")) ok = take_bool(c, p->vhs_head_switching);
        else if (is("vhs-linear-high-boost")) ok = take_double(c, p->vhs_linear_high_boost);
        else if (is("comp-pre")) ok = take_double(c, p->composite_preemphasis);
        else if (is("comp-cut")) ok = take_double(c, p->composite_preemphasis_cut);
        else if (is("comp-catv")) {                            // :1077-1096 (integer divisions as written)
            p->composite_preemphasis = 7;
            p->composite_preemphasis_cut = 315000000 / 88;
            p->video_chroma_phase_noise = 2;
        } else if (is("comp-catv2")) {
            p->composite_preemphasis = 15;
            p->composite_preemphasis_cut = 315000000 / 88;
            p->video_chroma_phase_noise = 4;
        } else if (is("comp-catv3")) {
            p->composite_preemphasis = 25;
            p->composite_preemphasis_cut = (315000000 * 2) / 88;
            p->video_chroma_phase_noise = 6;
        } else if (is("comp-catv4")) {
            p->composite_preemphasis = 40;
            p->composite_preemphasis_cut = (315000000 * 4) / 88;
            p->video_chroma_phase_noise = 6;
        } else if (is("vhs-linear-video-crosstalk")) ok = take_double(c, p->output_audio_linear_buzz);
        else if (is("chroma-phase-noise")) ok = take_int(c, p->video_chroma_phase_noise);
        else if (is("yc-recomb")) {
            double d;
            ok = take_double(c, d);
            if (ok) p->video_yc_recombine = (int32_t)d;
        } else if (is("audio-hiss")) ok = take_double(c, p->output_audio_hiss_db);
        else if (is("vhs-svideo")) ok = take_bool(c, p->vhs_svideo_out);
        else if (is("vhs-chroma-vblend")) ok = take_bool(c, p->vhs_chroma_vert_blend);
        else if (is("chroma-noise")) ok = take_int(c, p->video_chroma_noise);
        else if (is("noise")) ok = take_int(c, p->video_noise);
        else if (is("subcarrier-amp")) {
            ok = take_int(c, iv);
            if (ok) { p->subcarrier_amplitude = iv; p->subcarrier_amplitude_back = iv; }
        } else if (is("nocolor-subcarrier")) p->nocolor_subcarrier = 1;
        else if (is("nocolor-subcarrier-after-yc-sep")) p->nocolor_subcarrier_after_yc_sep = 1;
        else if (is("chroma-dropout")) ok = take_int(c, p->video_chroma_loss);
        else if (is("vhs")) {                                  // :1141-1151
            p->emulating_vhs = 1;
            p->vhs_head_switching = 1;
            p->emulating_preemphasis = 0;
            p->emulating_deemphasis = 0;
            p->output_audio_hiss_db = -70;
            p->video_chroma_phase_noise = 4;
            p->video_chroma_noise = 16;
            p->video_chroma_loss = 4;
            p->video_noise = 4;
        } else if (is("preemphasis")) ok = take_bool(c, p->emulating_preemphasis);
        else if (is("deemphasis")) ok = take_bool(c, p->emulating_deemphasis);
        else if (is("vhs-speed")) {                            // :1160-1189; does NOT enable head switching
            const char *v = c.next();
            if (!v) return CVS_ERR_BAD_SWITCH;
            p->emulating_vhs = 1;
            if (!std::strcmp(v, "ep")) {
                p->output_vhs_tape_speed = CVS_VHS_EP;
                p->video_chroma_phase_noise = 6; p->video_chroma_noise = 22;
                p->video_chroma_loss = 8; p->video_noise = 6;
            } else if (!std::strcmp(v, "lp")) {
                p->output_vhs_tape_speed = CVS_VHS_LP;
                p->video_chroma_phase_noise = 5; p->video_chroma_noise = 19;
                p->video_chroma_loss = 6; p->video_noise = 5;
            } else if (!std::strcmp(v, "sp")) {
                p->output_vhs_tape_speed = CVS_VHS_SP;
                p->video_chroma_phase_noise = 4; p->video_chroma_noise = 16;
                p->video_chroma_loss = 4; p->video_noise = 4;
            } else return CVS_ERR_BAD_SWITCH;
        } else if (is("vhs-hifi")) {                           // :1190-1203
            ok = take_bool(c, p->output_vhs_hifi);
            if (ok) {
                p->output_vhs_linear_audio = !p->output_vhs_hifi;
                p->emulating_vhs = 1;
                if (p->output_vhs_hifi) {
                    p->emulating_preemphasis = 1;
                    p->emulating_deemphasis = 1;
                    p->output_audio_hiss_db = -70;
                } else {
                    p->output_audio_hiss_db = -42;
                }
            }
        } else return CVS_ERR_BAD_SWITCH;                      // "Unknown switch", :1221-1224
        if (!ok) return CVS_ERR_BAD_SWITCH;
    }
    // post-parse derivation, :1264-1265 (int += double, truncating)
    if (p->composite_preemphasis != 0)
        p->subcarrier_amplitude_back = (int32_t)(
            p->subcarrier_amplitude_back +
            (50 * p->composite_preemphasis * (315000000 / 88)) / (2 * p->composite_preemphasis_cut));
    return CVS_OK;
}

unsigned long long cvs_draws_per_field(const cvs_params *p, int w, int h, unsigned field) {
    if (!p || w <= 0 || h <= 0 || (int)field >= h) return 0;
    const unsigned long long nl = (unsigned long long)((h - (int)field + 1) / 2);
    unsigned long long n = 0;
    if (p->video_noise != 0) n += nl * (unsigned long long)w;                          // :1632-1644
    if (p->vhs_head_switching && p->vhs_head_switching_phase_noise != 0) n += 4;       // :1654-1655
    if (p->video_chroma_noise != 0) n += 2ull * nl * (unsigned long long)w;            // :1719-1735
    if (p->video_chroma_phase_noise != 0) n += nl;                                     // :1736-1764
    if (p->video_chroma_loss != 0) n += nl;                                            // :1891-1901
    return n;
}

}  // extern "C"
