// field_plan.h -- host-side planning for the scanline kernels.
//
// Two kinds of data make a sharded field byte-identical to the reference's serial run:
//  * GeomPlan  (per geometry + parameter block + field parity, built once): where in the libc
//    rand() stream each stage of one composite_layer() call draws, and for every
//    field row the jump polynomials that take the generator from the field's first draw to the
//    row's luma / chroma noise segments (minus the warm-up).
//  * FieldSide (per field, cheap): the 61-word generator window at the field's first draw plus the
//    per-row values that depend on the per-LINE draws, which must replay glibc exactly: the chroma
//    phase-noise state (ffmpeg_ntsc.cpp:1744-1746), the dropout decision (:1896) and the head-switch
//    shift schedule (:1654-1712).
#ifndef CVS_FIELD_PLAN_H
#define CVS_FIELD_PLAN_H

#include <cstdint>
#include <vector>

#include "../../include/cvs_ntsc.h"
#include "glibc_rand.h"
#include "lane_pipeline.cuh"

namespace cvs {

struct GeomPlan {
    int w = 0, h = 0, nl = 0;
    unsigned field = 0;
    uint64_t offL = 0, offH = 0, offC = 0, offP = 0, offD = 0, ndraws = 0;
    bool has_luma = false, has_hs = false, has_chroma = false, has_phase = false, has_loss = false;
    // seek[row*62 + 0..30] = x^(offL + row*w - warmL(row)), seek[row*62 + 31..61] = x^(offC + 2*row*w - warmC(row))
    std::vector<uint32_t> seek;
    RandPoly jumpH, jumpP, jumpD, jumpN;   // field base -> head-switch draws / phase segment / dropout segment / next field
};

struct FieldSide {
    uint32_t window[kRandWindow];      // q[base-31 .. base+29]
    std::vector<uint32_t> rowinfo;     // rowinfo_pack(phase state, row flags), nl entries
    int hs_first = 0, hs_count = 0;    // rows [hs_first, hs_first+hs_count) are rotated
    std::vector<int32_t> hs_shift;     // hs_count shifts (all non-zero)
};

// warm_px: noise warm-up length in pixels (kWarmPx; shorter only to exercise the kernels' second-chance path in tests)
inline int warm_draws_luma(int row, int w, int warm_px = kWarmPx) {
    const long long full = (long long)row * w;
    return (int)(full < warm_px ? full : warm_px);
}

void build_geom_plan(const cvs_params &p, int w, int h, unsigned field, GeomPlan &g, int warm_px = kWarmPx);

// `cur` must sit at the field's first draw; on return it sits at the next field's first draw.
void build_field_side(const cvs_params &p, const GeomPlan &g, RandCursor &cur, FieldSide &fs);
// The same without moving the cursor (for planning several fields of a batch on several host threads:
// the caller advances a cursor of its own by g.jumpN per field and hands out copies).
void build_field_side_at(const cvs_params &p, const GeomPlan &g, const RandCursor &at, FieldSide &fs);

// exact n % m for n < 2^31: q = umulhi(n, magic) >> shift
void mod_magic(uint32_t m, uint32_t &magic, uint32_t &shift);

// launch constants; `lut` receives the {sin, cos} table K.phase_lut must point at (host copy)
template <typename R>
void make_kconst(const cvs_params &p, int w, int h, bool outfull, KConst<R> &K, std::vector<R> &lut);

// which kernel instantiation a parameter block needs
struct Variant {
    bool vhs;
    int cd;         // 9 / 12 / 14
    bool outfull;   // output lowpass = composite_lowpass (delays 2/4) instead of _tv (delay 1)
};
Variant pick_variant(const cvs_params &p);

}  // namespace cvs
#endif
