// yuv422_plan.h -- host-side planning for the 4:2:2 scanline kernel (see field_plan.h for the idea).
//
// The draw layout of one composite_video_process() call (ffmpeg_to_composite.cpp:653-941):
//   luma noise   nl * w draws, row-major                       (:653-665)
//   head switch  4 draws                                        (:675-676)
//   chroma noise nl * 2*(w/2) draws, U then V per sample        (:738-754)
//   phase noise  nl draws                                       (:763)
//   dropout      nl draws                                       (:936)
#ifndef CVS_YUV422_PLAN_H
#define CVS_YUV422_PLAN_H

#include <vector>

#include "../../include/cvs_yuv422.h"
#include "field_plan.h"
#include "yuv422_pipeline.cuh"

namespace cvs422 {

// GeomPlan / FieldSide are the BGRA path's structures (field_plan.h); seek[row*62 + 0..30] takes the
// generator to the row's luma segment minus its warm-up, seek[row*62 + 31..61] to its chroma segment.
void build_geom_plan(const cvs422_params &p, int w, int h, unsigned field, cvs::GeomPlan &g);

// Per-field side data at cursor `at` (not moved): generator window, per-row phase-noise state, dropout
// and head-switch records.  A head switch that is a short delay is flagged RG_HEADSW (done in the ring);
// any other rotation is flagged RG_HEADSW_PRE with hs_first / hs_count / hs_shift for the pre-pass.
void build_field_side_at(const cvs422_params &p, const cvs::GeomPlan &g, const cvs::RandCursor &at, cvs::FieldSide &fs);

// launch constants; `lut` receives the {cos, sin} table K.phase_lut must point at.  Returns a CVS_* status
// (CVS_ERR_INVALID_ARG for odd width / zero amplitudes, CVS_ERR_UNSUPPORTED for > kMaxRecombine rounds).
int make_k422(const cvs422_params &p, int w, int h, K422 &K, DivPair &dv, std::vector<double> &lut);

inline int warm_samples(long long preceding) { return (int)(preceding < kWarm ? preceding : kWarm); }

}  // namespace cvs422
#endif
