/* see tests/ffmpeg_decl/README.md -- declarations for a syntax check only */
#ifndef CVS_FFMPEG_DECL_AVFORMAT_H
#define CVS_FFMPEG_DECL_AVFORMAT_H
#include "../libavcodec/avcodec.h"
#define AVFMT_NOFILE 0x0001
#define AVFMT_GLOBALHEADER 0x0040
#define AVIO_FLAG_WRITE 2
typedef struct AVIOContext AVIOContext;
typedef struct AVInputFormat AVInputFormat;
typedef struct AVOutputFormat { const char *name; int flags; } AVOutputFormat;
typedef struct AVStream { int index; AVCodecParameters *codecpar; AVRational time_base; } AVStream;
typedef struct AVFormatContext { const AVOutputFormat *oformat; AVIOContext *pb; unsigned nb_streams; AVStream **streams; } AVFormatContext;
int avformat_open_input(AVFormatContext **ps, const char *url, const AVInputFormat *fmt, AVDictionary **options);
int avformat_find_stream_info(AVFormatContext *ic, AVDictionary **options);
int av_find_best_stream(AVFormatContext *ic, enum AVMediaType type, int wanted, int related, const AVCodec **decoder_ret, int flags);
void avformat_close_input(AVFormatContext **s);
int av_read_frame(AVFormatContext *s, AVPacket *pkt);
int avformat_alloc_output_context2(AVFormatContext **ctx, const AVOutputFormat *oformat, const char *format_name, const char *filename);
AVStream *avformat_new_stream(AVFormatContext *s, const AVCodec *c);
int avio_open(AVIOContext **s, const char *url, int flags);
int avio_closep(AVIOContext **s);
int avformat_write_header(AVFormatContext *s, AVDictionary **options);
int av_interleaved_write_frame(AVFormatContext *s, AVPacket *pkt);
int av_write_trailer(AVFormatContext *s);
void avformat_free_context(AVFormatContext *s);
#endif
