/* see tests/ffmpeg_decl/README.md -- declarations for a syntax check only */
#ifndef CVS_FFMPEG_DECL_AVCODEC_H
#define CVS_FFMPEG_DECL_AVCODEC_H
#include "../libavutil/avutil_decl.h"
enum AVCodecID { AV_CODEC_ID_NONE, AV_CODEC_ID_H264 = 27, AV_CODEC_ID_PCM_S16LE = 0x10000 };
#define AV_CODEC_FLAG_GLOBAL_HEADER (1 << 22)
typedef struct AVCodec AVCodec;
typedef struct AVCodecParameters AVCodecParameters;
typedef struct AVDictionary AVDictionary;
typedef struct AVCodecContext {
    int width, height, gop_size, max_b_frames, flags, sample_rate;
    AVRational sample_aspect_ratio, time_base;
    enum AVPixelFormat pix_fmt;
    enum AVSampleFormat sample_fmt;
    enum AVColorSpace colorspace;
    enum AVColorRange color_range;
    AVChannelLayout ch_layout;
} AVCodecContext;
typedef struct AVPacket { int stream_index; int64_t pts, dts; } AVPacket;
const AVCodec *avcodec_find_encoder(enum AVCodecID id);
AVCodecContext *avcodec_alloc_context3(const AVCodec *codec);
void avcodec_free_context(AVCodecContext **avctx);
int avcodec_parameters_to_context(AVCodecContext *codec, const AVCodecParameters *par);
int avcodec_parameters_from_context(AVCodecParameters *par, const AVCodecContext *codec);
int avcodec_open2(AVCodecContext *avctx, const AVCodec *codec, AVDictionary **options);
int avcodec_send_packet(AVCodecContext *avctx, const AVPacket *avpkt);
int avcodec_receive_frame(AVCodecContext *avctx, AVFrame *frame);
int avcodec_send_frame(AVCodecContext *avctx, const AVFrame *frame);
int avcodec_receive_packet(AVCodecContext *avctx, AVPacket *avpkt);
AVPacket *av_packet_alloc(void);
void av_packet_free(AVPacket **pkt);
void av_packet_unref(AVPacket *pkt);
void av_packet_rescale_ts(AVPacket *pkt, AVRational tb_src, AVRational tb_dst);
#endif
