// ref422_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin C wrapper around the reference's own composite_video_process() and render_field()
// (ffmpeg_to_composite.cpp), whose source is NOT in this repository: oracle/Makefile extracts them at
// build time by line range from /root/reference/ffmpeg_to_composite.cpp into the git-ignored
// oracle/_ref/yuv422_ref.inc and this file #includes that extract.  Written here: the few type shims
// the extract needs and C entry points that copy a cvs422_params block onto the reference's globals.
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <assert.h>

struct AVRational { int num, den; };
enum { AV_PIX_FMT_YUV420P = 0, AV_PIX_FMT_YUV422P = 4 };
struct AVFrame { uint8_t *data[8]; int linesize[8]; int width, height; int interlaced_frame, top_field_first; int format; };
struct AVCodecContext { int ticks_per_frame; };
static AVFrame *output_avstream_video_input_frame = NULL;
static AVCodecContext *input_avstream_video_codec_context = NULL;

#include "_ref/yuv422_ref.inc"

#include "../include/cvs_yuv422.h"

extern "C" {

void ref422_set_params(const cvs422_params *p) {
    output_ntsc = p->output_ntsc != 0;
    output_pal = !output_ntsc;
    output_width = p->output_width;
    output_height = p->output_height;
    video_scanline_phase_shift = p->video_scanline_phase_shift;
    video_scanline_phase_shift_offset = p->video_scanline_phase_shift_offset;
    composite_in_chroma_lowpass = p->composite_in_chroma_lowpass != 0;
    composite_out_chroma_lowpass = p->composite_out_chroma_lowpass != 0;
    composite_out_chroma_lowpass_lite = p->composite_out_chroma_lowpass_lite != 0;
    video_yc_recombine = p->video_yc_recombine;
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
    nocolor_subcarrier_after_yc_sep = p->nocolor_subcarrier_after_yc_sep != 0;
    enable_composite_emulation = p->enable_composite_emulation != 0;
    composite_preemphasis = p->composite_preemphasis;
    composite_preemphasis_cut = p->composite_preemphasis_cut;
    vhs_out_sharpen = p->vhs_out_sharpen;
    vhs_out_sharpen_chroma = p->vhs_out_sharpen_chroma;
    vhs_head_switching_phase = p->vhs_head_switching_phase;
    vhs_head_switching_phase_noise = p->vhs_head_switching_phase_noise;
}

void ref422_srand(unsigned seed) { srand(seed); }
int  ref422_rand(void) { return rand(); }

void ref422_composite_video_process(uint8_t *y, int ly, uint8_t *u, int lu, uint8_t *v, int lv,
                                    int w, int h, unsigned field, unsigned long long fieldno) {
    AVFrame d;
    memset(&d, 0, sizeof(d));
    d.data[0] = y; d.data[1] = u; d.data[2] = v;
    d.linesize[0] = ly; d.linesize[1] = lu; d.linesize[2] = lv;
    d.width = w; d.height = h; d.format = AV_PIX_FMT_YUV422P;
    composite_video_process(&d, field, fieldno);
}

// render_field(): the reference takes the 4:2:0 / 4:2:2 decision and the field timing from two
// globals; the harness fills them from plain arguments.
void ref422_render_field(uint8_t *const dst[3], const int dst_linesize[3], int dst_w, int dst_h,
                         uint8_t *const src[3], const int src_linesize[3], int src_h, int src_is_420,
                         int interlaced, int tff, int ticks_per_frame,
                         unsigned field, unsigned long long field_number, long long src_pts) {
    AVFrame d, s, fmt;
    AVCodecContext cc;
    memset(&d, 0, sizeof(d)); memset(&s, 0, sizeof(s)); memset(&fmt, 0, sizeof(fmt));
    for (int p = 0; p < 3; p++) {
        d.data[p] = dst[p]; d.linesize[p] = dst_linesize[p];
        s.data[p] = src[p]; s.linesize[p] = src_linesize[p];
    }
    d.width = dst_w; d.height = dst_h; d.format = AV_PIX_FMT_YUV422P;
    s.width = dst_w; s.height = src_h; s.interlaced_frame = interlaced; s.top_field_first = tff;
    fmt.format = src_is_420 ? AV_PIX_FMT_YUV420P : AV_PIX_FMT_YUV422P;
    cc.ticks_per_frame = ticks_per_frame;
    output_avstream_video_input_frame = &fmt;
    input_avstream_video_codec_context = &cc;
    render_field(&d, &s, field, field_number, src_pts);
}

}  // extern "C"
