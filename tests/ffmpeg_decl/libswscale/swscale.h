/* see tests/ffmpeg_decl/README.md -- declarations for a syntax check only */
#ifndef CVS_FFMPEG_DECL_SWS_H
#define CVS_FFMPEG_DECL_SWS_H
#include "../libavutil/avutil_decl.h"
#define SWS_BILINEAR 2
typedef struct SwsContext SwsContext;
typedef struct SwsFilter SwsFilter;
SwsContext *sws_getContext(int srcW, int srcH, enum AVPixelFormat srcFormat, int dstW, int dstH, enum AVPixelFormat dstFormat, int flags,
                           SwsFilter *srcFilter, SwsFilter *dstFilter, const double *param);
int sws_scale(SwsContext *c, const uint8_t *const srcSlice[], const int srcStride[], int srcSliceY, int srcSliceH,
              uint8_t *const dst[], const int dstStride[]);
void sws_freeContext(SwsContext *swsContext);
#endif
