// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin C wrapper around the reference's own composite_layer(), whose source is
// NOT in this repository: oracle/Makefile extracts it at build time by line range
// from /root/reference/ffmpeg_ntsc.cpp into the git-ignored oracle/_ref/ntsc_ref.inc
//  and this file #includes that extract.  The only things written
// here are the 8-line AVFrame shim the extract needs and C entry points that copy a
// cvs_params block onto the reference's file-scope globals.
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <vector>
using namespace std;

struct AVRational { int num, den; };
struct AVFrame { uint8_t *data[8]; int linesize[8]; int width, height; int interlaced_frame, top_field_first; };
class InputFile;

#include "_ref/ntsc_ref.inc"

#include "../include/cvs_ntsc.h"

extern "C" {

void ref_set_params(const cvs_params *p) {
    output_ntsc = p->output_ntsc != 0;
    output_pal = !output_ntsc;
    output_width = p->output_width;
    output_height = p->output_height;
    video_scanline_phase_shift = p->video_scanline_phase_shift;
    video_scanline_phase_shift_offset = p->video_scanline_phase_shift_offset;
    composite_in_chroma_lowpass = p->composite_in_chroma_lowpass != 0;
    composite_out_chroma_lowpass = p->composite_out_chroma_lowpass != 0;
    composite_out_chroma_lowpass_lite = p->composite_out_chroma_lowpass_lite != 0;
    video_noise = p->video_noise;
    video_chroma_noise = p->video_chroma_noise;
    video_chroma_phase_noise = p->video_chroma_phase_noise;
    video_chroma_loss = p->video_chroma_loss;
    subcarrier_amplitude = p->subcarrier_amplitude;
    subcarrier_amplitude_back = p->subcarrier_amplitude_back;
    emulating_vhs = p->emulating_vhs != 0;
    output_vhs_tape_speed = p->output_vhs_tape_speed;
    vhs_head_switching = p->vhs_head_switching != 0;
    vhs_chroma_vert_blend = p->vhs_chroma_vert_blend != 0;
    vhs_svideo_out = p->vhs_svideo_out != 0;
    nocolor_subcarrier = p->nocolor_subcarrier != 0;
    composite_preemphasis = p->composite_preemphasis;
    composite_preemphasis_cut = p->composite_preemphasis_cut;
    vhs_out_sharpen = p->vhs_out_sharpen;
    vhs_head_switching_point = p->vhs_head_switching_point;
    vhs_head_switching_phase = p->vhs_head_switching_phase;
    vhs_head_switching_phase_noise = p->vhs_head_switching_phase_noise;
}

// The BASELINE presets applied to the reference's own globals, for callers that must not load the product
// library (bench.py --impl reference).  What each switch sets is ffmpeg_ntsc.cpp:1141-1151 (-vhs) and :1160-1189
// (-vhs-speed); "comp" leaves the global initialisers (:756-809) as compiled.  Returns 0, or -1 for an unknown name.
int ref_set_preset(const char *name) {
    static bool saved = false;
    static int d_pn, d_cn, d_cl, d_vn, d_speed;
    static bool d_vhs, d_hs;
    if (!saved) {          // remember the initialisers so that presets do not accumulate
        d_pn = video_chroma_phase_noise; d_cn = video_chroma_noise; d_cl = video_chroma_loss; d_vn = video_noise;
        d_speed = output_vhs_tape_speed; d_vhs = emulating_vhs; d_hs = vhs_head_switching;
        saved = true;
    }
    video_chroma_phase_noise = d_pn; video_chroma_noise = d_cn; video_chroma_loss = d_cl; video_noise = d_vn;
    output_vhs_tape_speed = d_speed; emulating_vhs = d_vhs; vhs_head_switching = d_hs;
    if (!strcmp(name, "comp")) return 0;
    int pn, cn, cl, vn, speed;
    if (!strcmp(name, "sp")) { speed = VHS_SP; pn = 4; cn = 16; cl = 4; vn = 4; }
    else if (!strcmp(name, "lp")) { speed = VHS_LP; pn = 5; cn = 19; cl = 6; vn = 5; }
    else if (!strcmp(name, "ep")) { speed = VHS_EP; pn = 6; cn = 22; cl = 8; vn = 6; }
    else return -1;
    emulating_vhs = true;              // -vhs
    vhs_head_switching = true;
    output_vhs_tape_speed = speed;     // -vhs-speed <name>
    video_chroma_phase_noise = pn; video_chroma_noise = cn; video_chroma_loss = cl; video_noise = vn;
    return 0;
}

void ref_srand(unsigned seed) { srand(seed); }
int  ref_rand(void) { return rand(); }

void ref_composite_layer(uint8_t *dst, int dst_stride, const uint8_t *src, int src_stride,
                         int w, int h, int interlaced, int tff,
                         unsigned field, unsigned long long fieldno) {
    AVFrame d, s;
    memset(&d, 0, sizeof(d));
    memset(&s, 0, sizeof(s));
    d.data[0] = dst; d.linesize[0] = dst_stride; d.width = w; d.height = h;
    s.data[0] = (uint8_t *)src; s.linesize[0] = src_stride; s.width = w; s.height = h;
    s.interlaced_frame = interlaced; s.top_field_first = tff;
    composite_layer(&d, &s, *(InputFile *)0, field, fieldno);
}

}  // extern "C"
