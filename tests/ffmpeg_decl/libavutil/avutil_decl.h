/* see tests/ffmpeg_decl/README.md -- declarations for a syntax check only */
#ifndef CVS_FFMPEG_DECL_AVUTIL_H
#define CVS_FFMPEG_DECL_AVUTIL_H
#include <stdint.h>
typedef struct AVRational { int num, den; } AVRational;
enum AVPixelFormat { AV_PIX_FMT_NONE = -1, AV_PIX_FMT_YUV420P = 0, AV_PIX_FMT_YUV422P = 4, AV_PIX_FMT_YUVJ420P = 12, AV_PIX_FMT_YUVJ422P = 13,
                     AV_PIX_FMT_NV12 = 23, AV_PIX_FMT_BGRA = 28 };
enum AVSampleFormat { AV_SAMPLE_FMT_NONE = -1, AV_SAMPLE_FMT_U8, AV_SAMPLE_FMT_S16 };
enum AVMediaType { AVMEDIA_TYPE_UNKNOWN = -1, AVMEDIA_TYPE_VIDEO, AVMEDIA_TYPE_AUDIO };
enum AVColorSpace { AVCOL_SPC_SMPTE170M = 6 };
enum AVColorRange { AVCOL_RANGE_MPEG = 1 };
enum AVRounding { AV_ROUND_UP = 3 };
#define AV_NOPTS_VALUE ((int64_t)UINT64_C(0x8000000000000000))
typedef struct AVChannelLayout { int order, nb_channels; uint64_t mask; void *opaque; } AVChannelLayout;
void av_channel_layout_default(AVChannelLayout *ch_layout, int nb_channels);
typedef struct AVFrame {
    uint8_t *data[8];
    int linesize[8];
    uint8_t **extended_data;
    int width, height, nb_samples, format;
    int64_t pts, best_effort_timestamp;
    int sample_rate;
    AVChannelLayout ch_layout;
} AVFrame;
AVFrame *av_frame_alloc(void);
void av_frame_free(AVFrame **frame);
int av_frame_get_buffer(AVFrame *frame, int align);
int64_t av_rescale_q(int64_t a, AVRational bq, AVRational cq);
int64_t av_rescale_rnd(int64_t a, int64_t b, int64_t c, enum AVRounding rnd);
#endif
