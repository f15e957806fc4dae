/*
 * yuv422_oracle.h -- CPU restatement of the reference's 4:2:2 sibling path:
 * composite_video_process() (ffmpeg_to_composite.cpp:629-952) and render_field() (:1001-1129).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker.  The product (libcvs_ntsc.so) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_yuv422_oracle_vs_ref.py compares this restatement
 * byte-for-byte with the reference's own code (oracle/_ref/libref422.so, extracted at build
 * time from /root/reference/ffmpeg_to_composite.cpp); tests/golden/yuv422_*.npz hold committed
 * fixtures generated from that run (tests/golden/make_golden_yuv422.py).
 *
 * The restatement is line-major (all stages of one scanline, then the next), the reference is
 * plane-major (one stage over the whole field, then the next): agreement therefore also checks
 * the draw-offset arithmetic the GPU path relies on.
 */
#ifndef YUV422_ORACLE_H
#define YUV422_ORACLE_H

#include <stdint.h>
#include "../include/cvs_yuv422.h"
#include "ntsc_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

/* rand() draws one composite_video_process() call makes (depends on geometry, parity, parameters) */
unsigned long long oracle422_draws_per_field(const cvs422_params *p, int w, int h, unsigned field);

/* composite_video_process(dst, field, fieldno), :629.  In place; consumes draws from *g in the
 * reference's order.  Returns 0, or -1 for geometry this restatement does not define (odd width,
 * subcarrier_amplitude_back == 0: the reference writes out of bounds / divides by zero there).
 * The two luma bytes the reference reads past each row (:496) come from the plane when
 * y*linesize_y + w + 1 < linesize_y*h, else they are 0. */
int oracle422_composite_video_process(const cvs422_params *p, oracle_rng *g,
                                      uint8_t *y, int linesize_y, uint8_t *u, int linesize_u, uint8_t *v, int linesize_v,
                                      int w, int h, unsigned field, unsigned long long fieldno);

/* render_field(), :1001; second_field = ((field_number - src_pts) >= ticks_per_frame / 2), :1046-1050 */
void oracle422_render_field(uint8_t *const dst[3], const int dst_linesize[3], int dst_h,
                            const uint8_t *const src[3], const int src_linesize[3], int src_h,
                            const int row_bytes[3], int src_is_420,
                            int src_interlaced, int src_top_field_first, int second_field, unsigned field);

#ifdef __cplusplus
}
#endif
#endif
