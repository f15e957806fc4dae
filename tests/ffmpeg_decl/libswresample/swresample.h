/* see tests/ffmpeg_decl/README.md -- declarations for a syntax check only */
#ifndef CVS_FFMPEG_DECL_SWR_H
#define CVS_FFMPEG_DECL_SWR_H
#include "../libavutil/avutil_decl.h"
typedef struct SwrContext SwrContext;
int swr_alloc_set_opts2(SwrContext **ps, const AVChannelLayout *out_ch_layout, enum AVSampleFormat out_fmt, int out_rate,
                        const AVChannelLayout *in_ch_layout, enum AVSampleFormat in_fmt, int in_rate, int log_offset, void *log_ctx);
int swr_init(SwrContext *s);
void swr_free(SwrContext **s);
int64_t swr_get_delay(SwrContext *s, int64_t base);
int swr_convert(SwrContext *s, uint8_t **out, int out_count, const uint8_t **in, int in_count);
#endif
