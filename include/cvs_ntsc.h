/*
 * cvs_ntsc.h -- C ABI of the B200-native NTSC/VHS composite-video scanline engine.
 *
 * This header is the drop-in boundary for ONE path of
 * joncampbell123/composite-video-simulator: the per-field scanline DSP
 *
 *     composite_layer(AVFrame *dst, AVFrame *src, InputFile&, unsigned field,
 *                     unsigned long long fieldno)        ffmpeg_ntsc.cpp:1570-1921
 *
 * called from the reference's main loop at ffmpeg_ntsc.cpp:2229, with the ~30
 * file-scope globals it reads (ffmpeg_ntsc.cpp:205-214, 756-809) as an implicit
 * parameter block.  The reference has no plugin/FFI interface; the seam is that
 * direct call, so the ABI below mirrors it 1:1 (cvs_composite_layer) and adds a
 * batched, device-resident form (cvs_composite_fields*) for throughput.
 *
 * Plain C: pointers, sizes, PODs.  No torch / C++ types cross this boundary.
 * Every entry point returns 0 on success or a negative cvs_status; nothing throws.
 * There is NO CPU fallback: without a usable CUDA device every compute call
 * fails with CVS_ERR_CUDA.
 */
#ifndef CVS_NTSC_H
#define CVS_NTSC_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVS_ABI_VERSION 1

typedef enum cvs_status {
    CVS_OK = 0,
    CVS_ERR_INVALID_ARG = -1,   /* NULL pointer, bad geometry, stride < 4*w (ffmpeg_ntsc.cpp:1578-1583) */
    CVS_ERR_BAD_SWITCH = -2,    /* unknown / malformed argv switch (ffmpeg_ntsc.cpp:1221-1228)          */
    CVS_ERR_CUDA = -3,          /* CUDA runtime failure or no device                                    */
    CVS_ERR_NOMEM = -4,
    CVS_ERR_CAPACITY = -5,      /* w/h/batch larger than the context was created for                    */
    CVS_ERR_HELP = -6,          /* -h/-help seen (reference prints help and exits 1, :981-984)          */
    CVS_ERR_NOISE_SYNC = -7,    /* internal: a row's noise warm-up did not converge even after the kernel's
                                   second, 2048-pixel attempt (p < 2^-2000 per row); the batch's pictures are
                                   not the reference's.  Reported by the next synchronising call.          */
    CVS_ERR_UNSUPPORTED = -8
} cvs_status;

enum { CVS_VHS_SP = 0, CVS_VHS_LP = 1, CVS_VHS_EP = 2 };  /* ffmpeg_ntsc.cpp:803-807 */

/*
 * POD copy of the live globals composite_layer() reads.  Field names are the
 * reference's global names; the comment gives the defining line in
 * ffmpeg_ntsc.cpp and the CLI switch that sets it.
 */
typedef struct cvs_params {
    int32_t struct_size;                         /* = sizeof(cvs_params), ABI guard                     */
    int32_t output_ntsc;                         /* :209  -tvstd ntsc|pal (PAL => 0)                    */
    int32_t output_width;                        /* :207  -width                                        */
    int32_t output_height;                       /* :208  480 NTSC / 576 PAL (:815-831)                 */
    int32_t video_scanline_phase_shift;          /* :213  -comp-phase 0|90|180|270                      */
    int32_t video_scanline_phase_shift_offset;   /* :214  -comp-phase-offset                            */
    int32_t composite_in_chroma_lowpass;         /* :766  -in-composite-lowpass                         */
    int32_t composite_out_chroma_lowpass;        /* :767  -out-composite-lowpass                        */
    int32_t composite_out_chroma_lowpass_lite;   /* :768  -out-composite-lowpass-lite                   */
    int32_t video_noise;                         /* :775  -noise                                        */
    int32_t video_chroma_noise;                  /* :772  -chroma-noise                                 */
    int32_t video_chroma_phase_noise;            /* :773  -chroma-phase-noise                           */
    int32_t video_chroma_loss;                   /* :774  -chroma-dropout                               */
    int32_t subcarrier_amplitude;                /* :776  -subcarrier-amp                               */
    int32_t subcarrier_amplitude_back;           /* :777  (+ pre-emphasis term, :1264-1265)             */
    int32_t emulating_vhs;                       /* :791  -vhs / -vhs-speed / -vhs-hifi                 */
    int32_t output_vhs_tape_speed;               /* :809  -vhs-speed sp|lp|ep                           */
    int32_t vhs_head_switching;                  /* :761  -vhs / -vhs-head-switching                    */
    int32_t vhs_chroma_vert_blend;               /* :796  -vhs-chroma-vblend                            */
    int32_t vhs_svideo_out;                      /* :797  -vhs-svideo                                   */
    int32_t nocolor_subcarrier;                  /* :794  -nocolor-subcarrier                           */
    int32_t reserved0;
    double  composite_preemphasis;               /* :756  -comp-pre / -comp-catv*                       */
    double  composite_preemphasis_cut;           /* :757  -comp-cut / -comp-catv*                       */
    double  vhs_out_sharpen;                     /* :759  (no switch)                                   */
    double  vhs_head_switching_point;            /* :762  -vhs-head-switching-point                     */
    double  vhs_head_switching_phase;            /* :763  -vhs-head-switching-phase                     */
    double  vhs_head_switching_phase_noise;      /* :764  -vhs-head-switching-noise-level               */
    /* Parsed by the reference CLI but never read by composite_layer(); kept so that
       cvs_params_apply_argv() accepts exactly the switches parse_argv() accepts. */
    int32_t use_422_colorspace;                  /* :205  -422 / -420                                   */
    int32_t output_frame_delay;                  /* -d <1..256> (:1003-1011)                            */
    int32_t enable_composite_emulation;          /* :798  -nocomp (never read by the video path)        */
    int32_t enable_audio_emulation;              /* :799                                                */
    int32_t emulating_preemphasis;               /* :792  audio                                         */
    int32_t emulating_deemphasis;                /* :793  audio                                         */
    int32_t output_vhs_hifi;                     /* :788  audio                                         */
    int32_t output_vhs_linear_audio;             /* :790  audio                                         */
    int32_t nocolor_subcarrier_after_yc_sep;     /* :795  parsed, unused                                */
    int32_t video_yc_recombine;                  /* :770  parsed, unused                                */
    double  output_audio_hiss_db;                /* :778  audio                                         */
    double  output_audio_linear_buzz;            /* :779  audio                                         */
    double  vhs_linear_high_boost;               /* :782  audio                                         */
} cvs_params;

typedef struct cvs_ctx cvs_ctx;   /* opaque: device buffers, streams, RNG position, tables */

/* ---- parameter block ------------------------------------------------------------------- */

/* == the global initialisers (ffmpeg_ntsc.cpp:205-214,756-809) + preset_NTSC() (:824-831). */
int cvs_params_default_ntsc(cvs_params *p);
/* == preset_PAL() (:815-822) applied on top of *p. */
int cvs_params_preset_pal(cvs_params *p);
/*
 * Same order-dependent semantics as parse_argv() (ffmpeg_ntsc.cpp:972-1282) for every
 * switch it accepts (argv[0] is skipped, leading dashes stripped greedily, :980).
 * -i/-o take and ignore one argument (media I/O is outside the path).  The post-parse
 * derivation of subcarrier_amplitude_back (:1264-1265) is applied once per call.
 * Returns CVS_ERR_BAD_SWITCH for an unknown switch / bad value, CVS_ERR_HELP for -h.
 */
int cvs_params_apply_argv(cvs_params *p, int argc, const char *const *argv);

/* ---- context ---------------------------------------------------------------------------- */

/*
 * Create an engine on CUDA device `device` for pictures up to max_w x max_h and batches of
 * up to max_batch fields.  The RNG stream starts where an un-seeded glibc rand() starts
 * (the reference never calls srand()).
 */
int  cvs_create(cvs_ctx **out, const cvs_params *p, int device, int max_w, int max_h, int max_batch);
void cvs_destroy(cvs_ctx *ctx);
int  cvs_set_params(cvs_ctx *ctx, const cvs_params *p);
/*
 * Arithmetic of the scanline kernels: 0 (default) = fp32 with explicit FMAs, the production
 * path (within +-1 LSB of the reference); 1 = fp64 evaluated operation for operation like the
 * reference's double code (bit-exact with the reference; a validation mode, about 3x slower).
 */
int  cvs_set_precision(cvs_ctx *ctx, int use_double);
/*
 * Fuse the step that follows composite_layer() in the reference's field loop, the "field
 * deinterlace" line doubling (ffmpeg_ntsc.cpp:2232-2257): when enabled, every processed row y >= 1
 * is also written to row y-1 of the same picture (field 1: rows y-1 <- y for odd y; field 0: rows
 * y <- y+1 for y = 1,3,.. with y+1 < h, so for even h the last row is left as it was).  Off by default.
 */
int  cvs_set_bob(cvs_ctx *ctx, int enable);
/*
 * Per-pixel noise source (production fp32 arithmetic only; the fp64 validation mode always replays).
 *   CVS_NOISE_EXACT (default): every rand() draw of the reference is replayed (ffmpeg_ntsc.cpp:1640, 1729-1731).
 *   CVS_NOISE_FAST: the per-PIXEL luma / chroma noise draws (:1640, :1729, :1731) come from per-row counter
 *     generators instead; their amplitudes are sub-LSB in the reference's x256 domain, so the pictures stay within
 *     +-1 LSB of the reference (more values move by 1 than in exact mode).  The per-LINE draws (phase noise :1744,
 *     dropout :1896, head-switch jitter :1655) and the rand() stream position are exactly as in exact mode.
 *     Ignored (exact replay) when video_noise + 2 * video_chroma_noise > 96, where the noise is no longer sub-LSB.
 */
#define CVS_NOISE_EXACT 0
#define CVS_NOISE_FAST  1
int  cvs_set_noise_mode(cvs_ctx *ctx, int mode);

/* ---- the step after the field loop: BGRA -> the encoder's planar YUV ----------
 * ffmpeg_ntsc hands every finished BGRA picture to sws_scale() (ffmpeg_ntsc.cpp:2266-2274; BT.601 / SMPTE170M,
 * MPEG range, :2100-2101) before encoding.  This entry point does that conversion on the device, for n pictures,
 * asynchronously on the context's stream: y is w x h, u and v are ceil(w/2) x ceil(h/2) (CVS_YUV420P) or
 * ceil(w/2) x h (CVS_YUV422P); *_pic_stride are the byte distances between consecutive pictures of a plane.
 * PINNED against libswscale: the bytes are those of the library's own C code (= SWS_ACCURATE_RND | SWS_BITEXACT) for
 * sws_getContext(w, h, BGRA, w, h, YUV420P | YUV422P, SWS_BILINEAR, ...), including its chroma treatment (even widths:
 * chroma from the sum of the two pixels under a sample; 4:2:0: a 1-3-3-1 filter over the rows 2cy-1 .. 2cy+2, folded
 * at the edges; odd widths: full-width chroma through the library's horizontal bilinear bank) and its rounding
 * (csrc/yuv_convert.cuh states the arithmetic; checked against libswscale 9.1.100, tests/test_swscale_pin.py).  The
 * library's x86 SIMD code, which a plain SWS_BILINEAR context uses on x86 hosts, differs from its C code by +-1 on a few
 * percent of the 4:2:0 chroma samples and nowhere else.  CVS_ERR_UNSUPPORTED for pictures too small for the library's
 * filter (fewer than 3 rows / columns on a resampled axis).
 */
enum { CVS_YUV420P = 0, CVS_YUV422P = 1 };
int  cvs_bgra_to_yuv_device(cvs_ctx *ctx, void *y, int ly, long long y_pic_stride, void *u, int lu, long long u_pic_stride,
                            void *v, int lv, long long v_pic_stride, const void *bgra, int stride,
                            long long bgra_pic_stride, int w, int h, int n, int format);
/* The integer filter bank libswscale builds for one bilinear-resampled axis with centred sampling (srcn -> dstn samples;
 * one = 1 << 12 for rows, 1 << 14 for columns), as cvs_bgra_to_yuv_device applies it: pos[dstn] = first source sample,
 * coef[dstn * taps], every row summing to `one`.  Returns taps (<= max_taps) or a negative status.  Host only. */
int  cvs_sws_bilinear_bank(int srcn, int dstn, int one, int32_t *pos, int32_t *coef, int max_taps);

/* ---- the whole field loop around the seam, host pictures in, host pictures out ----------
 * What ffmpeg_ntsc's main loop does per output field (ffmpeg_ntsc.cpp:2190-2282), with only the decoder's pictures
 * going up and only the encoder's planar YUV coming down:
 *     frame_copy_scale() (:544-613)  ->  composite_layer() (:2229)  ->  line doubling (:2232-2257)
 *                                    ->  sws_scale() to the encoder's format (:2266-2274)
 * for n consecutive output fields first_fieldno .. first_fieldno + n - 1.  A 1080p field then costs 3.1 MB of
 * download (4:2:0) plus its share of a source picture instead of the 4.1 MB + 4.1 MB of field rows that
 * cvs_composite_fields_host moves, and no CPU pass touches a pixel.
 *   src*            nsrc decoder pictures in HOST memory (pinned recommended), format CVS_PIX_*, src_w x src_h
 *   src_of_field    n entries: which source picture each output field shows (the reference shows a decoded frame for
 *                   as many fields as its pts lasts); NULL = field k shows picture k * nsrc / n
 *   y, u, v         n output pictures in HOST memory, w x h luma, chroma per out_format (CVS_YUV420P / CVS_YUV422P)
 * The line doubling leaves row h-1 of an even-height picture of field 0 as the frame ring had it (:2247): the context
 * keeps that row from call to call (the default ring of one picture, `-d 1`), starting from the zeroed ring (:2069-2092).
 * Conversions as specified at cvs_scale_to_bgra_device and cvs_bgra_to_yuv_device (libswscale's bytes for planar YUV
 * sources); the composite_layer() step is the pinned hot path.  Synchronous; the rand() position advances as for n
 * cvs_composite_layer() calls.
 */
typedef struct cvs_field_loop {
    int32_t struct_size;                    /* sizeof(cvs_field_loop) */
    int32_t src_format, src_w, src_h;
    const void *src[3];
    int32_t src_linesize[3];
    int32_t nsrc;
    long long src_pic_stride[3];
    const int32_t *src_of_field;
    int32_t w, h;                           /* output_width x output_height */
    int32_t out_format, pad_;
    void *y, *u, *v;
    int32_t ly, lu, lv, pad2_;
    long long y_pic_stride, u_pic_stride, v_pic_stride;
} cvs_field_loop;
int  cvs_field_loop_host(cvs_ctx *ctx, const cvs_field_loop *d, int n, unsigned long long first_fieldno);

/* ---- the audio step of the same program (CPU) ----------------------------------
 * composite_audio_process(int16_t *audio, unsigned samples) (ffmpeg_ntsc.cpp:901-970; called per decoded audio
 * packet from process_audio(), :1284-1290): band limiting, pre/de-emphasis, sync-pulse buzz of linear tracks, tape
 * hiss, high boost -- in place on interleaved signed 16-bit samples at 44.1 kHz (output_audio_rate, :212) with
 * cvs_audio_channels(p) channels (output_audio_channels after parse_argv(), :1227-1262).  It needs no GPU.
 * The hiss draws rand() once per sample and channel (:952) from the SAME stream as the video noise, so a host
 * that interleaves audio packets and fields like the reference's main loop passes the stream position through:
 *     cvs_rng_tell(ctx, &pos); cvs_audio_process(a, pcm, n, &pos); cvs_rng_seek(ctx, pos);
 * Filter state and the sample counter (audio_proc_count, :889) persist in the cvs_audio object.
 */
typedef struct cvs_audio cvs_audio;
int  cvs_audio_channels(const cvs_params *p);
int  cvs_audio_create(cvs_audio **out, const cvs_params *p);          /* == "prepare audio filtering", :2031-2066 */
int  cvs_audio_process(cvs_audio *a, int16_t *audio, unsigned samples, unsigned long long *rng_pos /* in/out */);
void cvs_audio_destroy(cvs_audio *a);

/* ---- the step before the field loop: decoder picture -> BGRA at the output size ----------
 * InputFile::frame_copy_scale() (ffmpeg_ntsc.cpp:544-613) runs every decoded picture through sws_scale() to BGRA at
 * output_width x output_height (SWS_BILINEAR, :574-585) before composite_layer() reads it.  This entry point does
 * that on the device for n pictures, asynchronously on the context's stream, so a decoder that leaves its pictures in
 * device memory (NV12 from NVDEC, planar YUV) feeds the field loop without a CPU pass.
 *   format     CVS_PIX_BGRA (src[0]), CVS_PIX_YUV420P / CVS_PIX_YUV422P (src[0..2] = Y, U, V; chroma planes are
 *              ceil(sw/2) wide and ceil(sh/2) / sh high), CVS_PIX_NV12 (src[0] = Y, src[1] = interleaved UV)
 *   dst        n BGRA pictures, dw x dh, rows dst_stride bytes (multiple of 4), pictures dst_pic_stride bytes apart
 * Planar YUV sources (CVS_PIX_YUV420P / YUV422P / NV12; U and V planes of one layout): PINNED against libswscale -- the
 * bytes of the library's own C code (= SWS_ACCURATE_RND | SWS_BITEXACT) for sws_getContext(sw, sh, fmt, dw, dh, BGRA,
 * SWS_BILINEAR, ...).  Even dw: its bilinear banks, one chroma sample per pair of output pixels, its vertical dispatch
 * by tap counts, its ITU-R 601 colour tables, and its direct converter (no filtering) for YUV420P at the same size with
 * an even height.  Odd dw: its full-chroma-interpolation writers (one chroma sample per pixel, 32-bit integer matrix,
 * including their wrap-around for far-out-of-gamut samples).  Alpha 255.  (csrc/scale_convert.cuh states the arithmetic;
 * checked against libswscale 9.1.100, tests/test_swscale_pin.py; the library's x86 SIMD converter differs from its C code
 * by up to 3 codes.)
 * CVS_PIX_BGRA sources: PINNED likewise.  A source of the output size is copied, as the library does; at another size
 * the library converts packed RGB to YUV(A) at the precision of its 15-bit intermediates and back (chroma per pixel, or
 * from pixel pairs when the width shrinks to half or less; alpha a << 6 | a >> 2 as a fourth plane; the full-chroma
 * writers), which is what this does.  The one geometry left on the repository's own resampler, NOT pinned (triangle
 * kernel with 14-bit weights, channel by channel; specified in csrc/scale_convert.cuh, restated in
 * oracle/convert_oracle.c): a BGRA source of ODD width reduced to half its width or less (the library keeps chroma per
 * pixel there; not routed yet).
 * Shrinking by more than 16x per axis returns CVS_ERR_CAPACITY.
 */
enum { CVS_PIX_BGRA = 0, CVS_PIX_YUV420P = 1, CVS_PIX_YUV422P = 2, CVS_PIX_NV12 = 3 };
int  cvs_scale_to_bgra_device(cvs_ctx *ctx, void *dst, int dst_stride, long long dst_pic_stride, int dw, int dh,
                              const void *const src[3], const int src_linesize[3], const long long src_pic_stride[3],
                              int sw, int sh, int format, int n);

/* ---- the seam: exact analogue of the call at ffmpeg_ntsc.cpp:2229 ------------------------- */

/*
 * HOST pointers, BGRA8 (uint32 = A<<24|R<<16|G<<8|B), strides in bytes.  Reads the field
 * rows of src, writes ONLY dst rows y == field (mod 2) (alpha byte = 0, :1914), leaves the
 * other rows untouched.  Advances the context's RNG position by exactly the number of
 * rand() draws the reference call makes.  Synchronous.  Invalid geometry returns
 * CVS_ERR_INVALID_ARG and leaves dst untouched (the reference returns silently, :1578-1583).
 */
int cvs_composite_layer(cvs_ctx *ctx,
                        uint8_t *dst, int dst_stride,
                        const uint8_t *src, int src_stride,
                        int w, int h, int src_interlaced, int src_top_field_first,
                        unsigned field, unsigned long long fieldno);

/* ---- throughput forms -------------------------------------------------------------------- */

/*
 * n consecutive calls of the seam in one launch: picture k (k = 0..n-1) lives at
 * src + k*src_pic_stride / dst + k*dst_pic_stride, is processed with
 * field = ((first_fieldno + k) & 1) ^ 1 and fieldno = first_fieldno + k  (the schedule of
 * the reference loop, ffmpeg_ntsc.cpp:2229) and consumes the RNG stream in call order, so
 * the result is byte-identical to n sequential cvs_composite_layer() calls.
 *
 * cvs_composite_fields_device: src/dst are DEVICE pointers (any stream-visible memory);
 * the launch is asynchronous on the context stream (see cvs_synchronize).
 * cvs_composite_fields_host: src/dst are HOST pointers (pinned recommended); the call
 * copies in, computes, copies the field rows back and returns when done.
 */
int cvs_composite_fields_device(cvs_ctx *ctx,
                                void *dst, size_t dst_pic_stride, int dst_stride,
                                const void *src, size_t src_pic_stride, int src_stride,
                                int w, int h, int src_interlaced, int src_top_field_first,
                                int n, unsigned long long first_fieldno);
int cvs_composite_fields_host(cvs_ctx *ctx,
                              void *dst, size_t dst_pic_stride, int dst_stride,
                              const void *src, size_t src_pic_stride, int src_stride,
                              int w, int h, int src_interlaced, int src_top_field_first,
                              int n, unsigned long long first_fieldno);
/*
 * cvs_composite_fields_host_async: as _host, but returns as soon as the work is queued; the pictures
 * in dst are complete after cvs_synchronize().  Consecutive asynchronous calls overlap (the upload of
 * one call runs while the previous one computes and downloads), so a stream of batches keeps both PCIe
 * directions busy without the fill/drain gap of a synchronous call.  The caller must not reuse src/dst
 * of a call before the synchronize that follows it (use two buffer sets).  Host memory should be pinned.
 */
int cvs_composite_fields_host_async(cvs_ctx *ctx,
                                    void *dst, size_t dst_pic_stride, int dst_stride,
                                    const void *src, size_t src_pic_stride, int src_stride,
                                    int w, int h, int src_interlaced, int src_top_field_first,
                                    int n, unsigned long long first_fieldno);
int cvs_synchronize(cvs_ctx *ctx);
/*
 * Page-locked host memory for the pictures handed to the _host entry points (what the reference keeps in its
 * AVFrame buffers, ffmpeg_ntsc.cpp:2069-2092): pageable memory halves the transfer rate.  Plain
 * cudaHostAlloc / cudaFreeHost on the context's device.
 */
int cvs_alloc_host(cvs_ctx *ctx, void **out, size_t bytes);
int cvs_free_host(cvs_ctx *ctx, void *p);

/* ---- RNG stream (hidden state of the reference: the libc rand() position) ------------------ */

/* Absolute position: number of rand() draws consumed since program start (seed 1). */
int cvs_rng_seek(cvs_ctx *ctx, unsigned long long draws_consumed);
int cvs_rng_tell(const cvs_ctx *ctx, unsigned long long *draws_consumed);
/* Draws one composite_layer() call consumes for this geometry/params. */
unsigned long long cvs_draws_per_field(const cvs_params *p, int w, int h, unsigned field);

/*
 * Largest batch size <= max_batch (and <= the context's capacity) whose scanline tasks fill a whole
 * number of waves of the device for this geometry and parameter block: every lane processes one
 * scanline, all tasks take the same time, so a partial last wave is pure loss (a 256-field batch of
 * 1080p on a B200 runs 2.6 waves in the time of 3; 296 fields fill 3).  Returns the batch size (> 0)
 * or a negative cvs_status.
 */
int cvs_preferred_batch(cvs_ctx *ctx, int w, int h, int max_batch);

/* ---- introspection -------------------------------------------------------------------------- */

const char *cvs_strerror(int status);
int         cvs_abi_version(void);
/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
unsigned long long cvs_kernel_launches(const cvs_ctx *ctx);
/*
 * Device time of the fused scanline kernel (k_fields), measured with CUDA events recorded on
 * the context stream around every launch since the last reset (up to 4096 launches).  The query
 * synchronises the stream and returns the sum of the per-launch durations and their count.
 */
int cvs_kernel_time_reset(cvs_ctx *ctx);
int cvs_kernel_time_query(cvs_ctx *ctx, double *total_ms, int *launches);
/*
 * Run on a caller-owned CUDA stream (a cudaStream_t / CUstream handle, e.g. torch's
 * torch.cuda.current_stream().cuda_stream) instead of the context's private stream, so the
 * caller's own events and allocations order with the engine's work.
 */
int cvs_set_stream(cvs_ctx *ctx, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* CVS_NTSC_H */
