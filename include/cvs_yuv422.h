/*
 * cvs_yuv422.h -- C ABI of the B200 engine for the reference's SECOND scanline path:
 * composite_video_process() of ffmpeg_to_composite.cpp (8-bit planar 4:2:2, in place), plus the
 * field renderer that feeds it (render_field()).  SURVEY.md section 8(f) rows 2 and 3.
 *
 * Seam replaced (ffmpeg_to_composite.cpp:1783-1790):
 *
 *     render_field(output_avstream_video_frame, output_avstream_video_input_frame,
 *                  (video_field & 1) ^ 1, video_field, tgt_pts);                       // :1784 -> :1001
 *     composite_video_process(output_avstream_video_frame, (video_field & 1) ^ 1, video_field); // :1790 -> :629
 *
 * composite_video_process() works IN PLACE on the rows y == field (mod 2) of a planar YUV 4:2:2
 * picture (AVFrame::data[0..2], linesize[0..2], width, height), reads ~30 file-scope globals
 * (mirrored by cvs422_params, same names) and draws from the libc rand() stream (mirrored by the
 * context: seed 1, never reseeded).  Arithmetic is the reference's own: every filter runs in IEEE
 * double in the reference's operation order on the GPU, so results are BIT-EXACT, not "within
 * tolerance" (tests/test_gpu_yuv422.py).
 *
 * One documented difference: the reference reads two luma bytes past the end of every row
 * (Y[x+2], :496).  This engine reads the same two bytes when they lie inside the plane the caller
 * described (y*linesize + width + 1 < linesize*height) and uses 0 otherwise, where the reference
 * reads whatever follows the buffer.
 *
 * No CPU fallback exists: without a CUDA device cvs422_create() fails with CVS_ERR_CUDA.
 * Status codes and cvs_strerror() are those of cvs_ntsc.h.
 */
#ifndef CVS_YUV422_H
#define CVS_YUV422_H

#include <stdint.h>
#include "cvs_ntsc.h"

#ifdef __cplusplus
extern "C" {
#endif

/* POD copy of the globals of ffmpeg_to_composite.cpp that the video path reads (:266-317).
 * Field names are the reference's. */
typedef struct cvs422_params {
    int32_t output_ntsc;                        /* :299  (output_pal == !output_ntsc for this path)   */
    int32_t output_width, output_height;        /* :297-298                                            */
    int32_t video_scanline_phase_shift;         /* :285  0 / 90 / 180 / 270                            */
    int32_t video_scanline_phase_shift_offset;  /* :286                                                */
    int32_t composite_in_chroma_lowpass;        /* :279                                                */
    int32_t composite_out_chroma_lowpass;       /* :280                                                */
    int32_t composite_out_chroma_lowpass_lite;  /* :281                                                */
    int32_t video_yc_recombine;                 /* :283                                                */
    int32_t video_noise;                        /* :290                                                */
    int32_t video_chroma_noise;                 /* :287                                                */
    int32_t video_chroma_phase_noise;           /* :288                                                */
    int32_t video_chroma_loss;                  /* :289                                                */
    int32_t subcarrier_amplitude;               /* :291                                                */
    int32_t subcarrier_amplitude_back;          /* :292                                                */
    int32_t emulating_vhs;                      /* :319                                                */
    int32_t output_vhs_tape_speed;              /* :341  CVS_VHS_SP / LP / EP                          */
    int32_t vhs_head_switching;                 /* :275                                                */
    int32_t vhs_chroma_vert_blend;              /* :324                                                */
    int32_t vhs_svideo_out;                     /* :325                                                */
    int32_t nocolor_subcarrier;                 /* :322                                                */
    int32_t nocolor_subcarrier_after_yc_sep;    /* :323                                                */
    int32_t enable_composite_emulation;         /* :326  0: composite_video_process is not called     */
    int32_t reserved0;
    double  composite_preemphasis;              /* :268                                                */
    double  composite_preemphasis_cut;          /* :269                                                */
    double  vhs_out_sharpen;                    /* :271                                                */
    double  vhs_out_sharpen_chroma;             /* :272                                                */
    double  vhs_head_switching_phase;           /* :276                                                */
    double  vhs_head_switching_phase_noise;     /* :277                                                */
} cvs422_params;

typedef struct cvs422_ctx cvs422_ctx;

/* The reference's global initialisers (ffmpeg_to_composite.cpp:266-341). */
int cvs422_params_default(cvs422_params *p);
/* preset_PAL() / preset_NTSC() (:1252-1270), as -tvstd applies them. */
int cvs422_params_preset_pal(cvs422_params *p);
int cvs422_params_preset_ntsc(cvs422_params *p);
/* parse_argv() (:1292-1650) restricted to what the video path honours: same switch names, same
 * order dependence (-vhs / -vhs-speed overwrite the noise levels, -comp-catv* set pre-emphasis and
 * phase noise, the final subcarrier_amplitude_back adjustment of :1626-1627).  argv[0] is skipped.
 * Audio / container / stream-selection switches are accepted and ignored (with their argument).
 * Unknown switch: CVS_ERR_BAD_SWITCH; -h: CVS_ERR_HELP. */
int cvs422_params_apply_argv(cvs422_params *p, int argc, const char *const *argv);

int  cvs422_create(cvs422_ctx **out, const cvs422_params *p, int device, int max_w, int max_h, int max_batch);
void cvs422_destroy(cvs422_ctx *ctx);
int  cvs422_set_params(cvs422_ctx *ctx, const cvs422_params *p);

/* Drop-in for composite_video_process(dst, field, fieldno) (:629): host pointers to the three planes
 * of a 4:2:2 picture (chroma planes are width/2 samples wide, full height), processed in place. */
int cvs422_composite_video_process(cvs422_ctx *ctx,
                                   uint8_t *y, int linesize_y, uint8_t *u, int linesize_u, uint8_t *v, int linesize_v,
                                   int w, int h, unsigned field, unsigned long long fieldno);

/* Throughput forms: n consecutive fields; picture k starts pic_stride_* bytes after picture k-1 in each
 * plane, is processed as field ((first_fieldno + k) & 1) ^ 1 with fieldno first_fieldno + k, exactly as
 * n sequential calls would (including the rand() stream).  _device takes device pointers and is
 * asynchronous on the context's stream; _host takes host pointers (pinned recommended) and returns
 * when the pictures are back. */
int cvs422_process_fields_device(cvs422_ctx *ctx, uint8_t *y, uint8_t *u, uint8_t *v,
                                 long long pic_stride_y, long long pic_stride_u, long long pic_stride_v,
                                 int linesize_y, int linesize_u, int linesize_v,
                                 int w, int h, int n, unsigned long long first_fieldno);
int cvs422_process_fields_host(cvs422_ctx *ctx, uint8_t *y, uint8_t *u, uint8_t *v,
                               long long pic_stride_y, long long pic_stride_u, long long pic_stride_v,
                               int linesize_y, int linesize_u, int linesize_v,
                               int w, int h, int n, unsigned long long first_fieldno);

/* Drop-in for render_field(dst, src, field, field_number, src_pts) (:1001): vertical 8.8 fixed-point
 * resampling of a width-scaled source picture (4:2:0 or 4:2:2, src_h rows) onto the rows
 * y == field (mod 2) of a 4:2:2 picture of dst_h rows.  Device pointers, asynchronous on the context's
 * stream.  src_is_420: the source chroma planes have src_h/2 rows (:1006-1009).  src_interlaced /
 * src_top_field_first are the AVFrame flags; second_field is the reference's
 * (field_number - src_pts) >= ticks_per_frame/2 test (:1046-1050), evaluated by the caller.
 * row_bytes[p] bytes are produced per row and plane (the reference copies src->linesize[p]). */
int cvs422_render_field_device(cvs422_ctx *ctx,
                               uint8_t *const dst[3], const int dst_linesize[3], int dst_h,
                               const uint8_t *const src[3], const int src_linesize[3], int src_h,
                               const int row_bytes[3], int src_is_420,
                               int src_interlaced, int src_top_field_first, int second_field, unsigned field);

int cvs422_synchronize(cvs422_ctx *ctx);
int cvs422_set_stream(cvs422_ctx *ctx, void *cuda_stream);

/* rand() position of the context (draws consumed since srand(1)); see cvs_rng_seek in cvs_ntsc.h */
int cvs422_rng_seek(cvs422_ctx *ctx, unsigned long long draws_consumed);
unsigned long long cvs422_rng_tell(const cvs422_ctx *ctx);
/* draws one composite_video_process() call makes for this geometry / field parity */
unsigned long long cvs422_draws_per_field(const cvs422_params *p, int w, int h, unsigned field);

/* kernels launched so far / device time of the scanline kernel since the last reset (CUDA events) */
unsigned long long cvs422_kernel_launches(const cvs422_ctx *ctx);
int cvs422_kernel_time_reset(cvs422_ctx *ctx);
int cvs422_kernel_time_query(cvs422_ctx *ctx, double *total_ms, int *launches);

#ifdef __cplusplus
}
#endif
#endif
