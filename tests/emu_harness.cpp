// emu_harness.cpp -- TEST INFRASTRUCTURE: runs the product's lane pipeline on the CPU.
//
// composite_video_simulator_b200/csrc/lane_pipeline.cuh is host/device portable; this harness
// drives it exactly the way the CUDA kernels do (one lane per field row, 32 lanes per warp in
// lock-step, lane 0 = halo row, the vertical-blend exchange done between the two halves of a
// step) so the kernel LOGIC -- block lags, line-edge quirks, RNG seeking, noise warm-up, the
// head-switch pre-pass, the host-side planning in field_plan.cpp -- can be checked against the
// oracle without a GPU: precision 1 (double) must match the oracle bit for bit, precision 0
// (float) is the production arithmetic and must stay within +-1 LSB.
// It is never linked into libcvs_ntsc.so and no product entry point reaches it.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../composite_video_simulator_b200/csrc/field_plan.h"
#include "../composite_video_simulator_b200/csrc/lane_pipeline.cuh"

using namespace cvs;

namespace {

int g_warm_px = kWarmPx;      // emu_set_warm_px(): a short warm-up exercises the second-chance path (rewarm_bracket)

// the kernels' warm-up with its second chance (scanline_kernels.cuh)
template <typename R, typename G>
bool warm_up(const KConst<R> &K, bool chroma, G &g, uint32_t *ring, int stride, const uint32_t hist[31], int row, int w,
             int &s0, int &s1) {
    const long long full = (long long)row * w;
    const int nd = (int)(full < g_warm_px ? full : g_warm_px);
    const bool from_start = full <= g_warm_px;
    const int v = chroma ? K.cnoise : K.vnoise;
    const uint32_t m = (uint32_t)(2 * v + 1), magic = chroma ? K.cmagic : K.vmagic, shift = chroma ? K.cshift : K.vshift;
    const uint32_t n0 = kRngBase - (chroma ? 2u : 1u) * (uint32_t)nd;
    g.init(ring, 0, stride, hist, n0);
    bool ok = chroma ? warm_chroma(m, magic, shift, v, g, nd, from_start, s0, s1) : warm_luma(m, magic, shift, v, g, nd, from_start, s0);
    if (ok) return true;
    const long long avail = full - nd;
    const int extra = (int)(avail < kRewarmPx ? avail : kRewarmPx);
    int br[4];
    rewarm_bracket(hist, chroma ? 2 : 1, extra, extra == avail, m, magic, shift, v, br);
    g.init(ring, 0, stride, hist, n0);
    return chroma ? warm_chroma(m, magic, shift, v, g, nd, from_start, s0, s1, br) : warm_luma(m, magic, shift, v, g, nd, from_start, s0, br);
}

// NF: fast noise mode (cvs_set_noise_mode): the kernels' NF instantiations
template <typename R, bool VHS, int CD, bool OUTFULL, bool NF>
int run_field(const cvs_params &p, RandCursor &cur, uint8_t *dst, int dst_stride, const uint8_t *src,
              int src_stride, int w, int h, int interlaced, int tff, unsigned field,
              unsigned long long fieldno, int force_general) {
    typedef Lane<R, VHS, CD, OUTFULL> L;
    typedef Pipeline<R, VHS, CD, OUTFULL> P;
    GeomPlan g;
    build_geom_plan(p, w, h, field, g, g_warm_px);
    FieldSide fs;
    build_field_side(p, g, cur, fs);
    KConst<R> K;
    std::vector<R> lut;
    make_kconst<R>(p, w, h, OUTFULL, K, lut);
    if (force_general == 1) K.flags |= F_GENERAL;
    if (NF) K.flags |= F_NOISE_FAST;
    const int nl = g.nl;
    const int opposite = interlaced ? (tff ? 1 : 0) : 0;
    auto src_row = [&](int row) {
        int sy = (int)field + 2 * row + opposite;
        if (sy > h - 1) sy = h - 1;
        return (const uint32_t *)(src + (size_t)src_stride * (size_t)sy);
    };

    // head-switch pre-pass
    std::vector<int32_t> scratch((size_t)(fs.hs_count > 0 ? fs.hs_count : 1) * (size_t)w);
    for (int i = 0; i < fs.hs_count; i++) {
        const int row = fs.hs_first + i;
        Lane<R, false, 9, false> ln;
        ln.reset(K);
        RowConst<R> rc;
        row_setup<R>(K, field, fieldno, row, fs.rowinfo[(size_t)row], rc);
        uint32_t ring[kRngSlots];
        ln.nY = 0;
        ln.rngL.l1 = lcg_seed(fieldno, field, row, 0u);
        if (K.vnoise != 0 && !NF) {
            uint32_t hist[31];
            rng_rebase(fs.window, &g.seek[(size_t)row * 62], hist);
            ln.nY = 0;
            int unused = 0;
            if (!warm_up<R>(K, false, ln.rngL, ring, 1, hist, row, w, ln.nY, unused)) return CVS_ERR_NOISE_SYNC;
        }
        headswitch_row<R, NF>(K, rc, ln, src_row(row), &scratch[(size_t)i * w], fs.hs_shift[(size_t)i]);
    }

    const int nwarps = (nl + 30) / 31;
    const int nsteps = line_steps<VHS, CD>(w);
    int s_lo, s_hi;
    interior_steps<VHS, CD>(w, s_lo, s_hi);
    std::vector<uint32_t> rings((size_t)32 * 2 * kRngSlots);
    std::vector<R> tails((size_t)32 * 2 * kTailSlots);
    std::vector<R> hsring((size_t)32 * kHsSlots);
    for (int wp = 0; wp < nwarps; wp++) {
        L lane[32];
        RowConst<R> rc[32];
        const uint32_t *srow[32];
        uint32_t *drow[32];
        const int32_t *hsrow[32];
        bool valid[32];
        for (int l = 0; l < 32; l++) {
            int row = 31 * wp + l - 1;
            valid[l] = (l >= 1) && row < nl;
            if (row < 0) row = 0;
            if (row > nl - 1) row = nl - 1;
            L &ln = lane[l];
            ln.reset(K);
            row_setup<R>(K, field, fieldno, row, fs.rowinfo[(size_t)row], rc[l]);
            srow[l] = src_row(row);
            drow[l] = (uint32_t *)(dst + (size_t)dst_stride * (size_t)((int)field + 2 * row));
            hsrow[l] = (rc[l].rflags & RF_HEADSW) ? &scratch[(size_t)(row - fs.hs_first) * w] : (const int32_t *)0;
            ln.tailU = &tails[(size_t)l * 2 * kTailSlots];
            ln.tailV = ln.tailU + kTailSlots;
            ln.tail_stride = 1;
            ln.nY = ln.nU = ln.nV = 0;
            uint32_t hist[31];
            if (NF) {                                  // fast noise: per-row counter generators, no replay
                ln.rngL.l1 = lcg_seed(fieldno, field, row, 0u);
                ln.rngC.l1 = lcg_seed(fieldno, field, row, 1u);
            }
            if (K.vnoise != 0 && !NF) {
                rng_rebase(fs.window, &g.seek[(size_t)row * 62], hist);
                int unused = 0;
                if (!warm_up<R>(K, false, ln.rngL, &rings[(size_t)l * 2 * kRngSlots], 1, hist, row, w, ln.nY, unused))
                    return CVS_ERR_NOISE_SYNC;
            }
            if (K.cnoise != 0 && !NF) {
                rng_rebase(fs.window, &g.seek[(size_t)row * 62 + 31], hist);
                if (!warm_up<R>(K, true, ln.rngC, &rings[(size_t)l * 2 * kRngSlots + kRngSlots], 1, hist, row, w, ln.nU, ln.nV))
                    return CVS_ERR_NOISE_SYNC;
            }
        }
        bool warp_inl = false;
        for (int l = 0; l < 32; l++) warp_inl |= VHS && rc[l].hs_delay > 0;
        bool odd_any = false;
        for (int l = 0; l < 32; l++) odd_any |= (rc[l].xi & 1) != 0;
        for (int l = 0; l < 32; l++) rc[l].odd_any = odd_any;
        // the kernel's lean interior loop (scanline_kernels.cuh): no pre-pass head-switch row, no dropout row
        bool lean = true;
        for (int l = 0; l < 32; l++) lean &= hsrow[l] == nullptr && !(rc[l].rflags & RF_DROPOUT);
        for (int s = 0; s < nsteps; s++) {
            // the kernel's choice of code variant for this step (force_general: 1 = general everywhere,
            // 2 = edge variant everywhere, to exercise those variants on interior blocks as well)
            int mode = (K.flags & F_GENERAL) ? MODE_GENERAL : ((s >= s_lo && s < s_hi) ? MODE_FAST : MODE_EDGE);
            if (force_general == 2 && mode == MODE_FAST) mode = MODE_EDGE;
            // MODE_FAST warps whose rows all have an even line phase run the MODE_FAST_EVEN loop (scanline_kernels.cuh)
            if (mode == MODE_FAST && !odd_any && lean) mode = MODE_FAST_EVEN;
            BlendXchg<R> xo[32];
            R Yb[32][kT];
            V2<R> IQb[32][kT];
            for (int l = 0; l < 32; l++) {
                uint32_t px[kT], pxprev[kT];
                load_block_scalar(srow[l], s, w, px);
                if (s >= 1) load_block_scalar(srow[l], s - 1, w, pxprev);
                else for (int j = 0; j < kT; j++) pxprev[j] = 0;
                R C[kT];
                R *ring = &hsring[(size_t)l * kHsSlots];
#define CVS_HEAD(M)                                                                                          \
    {                                                                                                        \
        constexpr bool FAST = (M) <= 0;                                                                      \
        P::template stage_a<M, NF>(K, rc[l], lane[l], s, px, pxprev, hsrow[l], C);                           \
        if (FAST) headswitch_substitute<R>(rc[l], hsrow[l], s - 1, C);                                       \
        if (warp_inl && (FAST || s >= 1)) headswitch_delay_block<R>(ring, 1, s - 1, w, rc[l].hs_delay, C);   \
        P::template stage_b<M, NF>(K, rc[l], lane[l], s, C, Yb[l], IQb[l], xo[l]);                           \
    }
                if (mode == MODE_FAST_EVEN) CVS_HEAD(MODE_FAST_EVEN)
                else if (mode == MODE_FAST) CVS_HEAD(MODE_FAST)
                else if (mode == MODE_EDGE) CVS_HEAD(MODE_EDGE)
                else CVS_HEAD(MODE_GENERAL)
#undef CVS_HEAD
            }
            for (int l = 0; l < 32; l++) {
                R Yf[kT];
                V2<R> IQf[kT];
                int kf;
                uint32_t out[kT];
                bool have;
                const BlendXchg<R> &above = xo[l > 0 ? l - 1 : 0];
#define CVS_TAIL(M)                                                                                          \
    if (VHS) {                                                                                               \
        P::template stage_c<M>(K, rc[l], lane[l], s, Yb[l], xo[l], above, Yf, IQf, kf);                      \
        have = P::template stage_f<M, (M) == MODE_FAST_EVEN>(K, rc[l], lane[l], kf, Yf, IQf, out);           \
    } else {                                                                                                 \
        kf = s - 1 - kLB;                                                                                    \
        have = P::template stage_f<M, (M) == MODE_FAST_EVEN>(K, rc[l], lane[l], kf, Yb[l], IQb[l], out);     \
    }
                if (mode == MODE_FAST_EVEN) { CVS_TAIL(MODE_FAST_EVEN) }
                else if (mode == MODE_FAST) { CVS_TAIL(MODE_FAST) }
                else if (mode == MODE_EDGE) { CVS_TAIL(MODE_EDGE) }
                else { CVS_TAIL(MODE_GENERAL) }
#undef CVS_TAIL
                if (have && valid[l]) {
                    const int x0 = (kf - 1) * kT;
                    for (int j = 0; j < kT; j++)
                        if (x0 + j < w) drow[l][x0 + j] = out[j];
                }
            }
        }
    }
    return 0;
}

template <typename R, bool NF>
int dispatch(const cvs_params &p, RandCursor &cur, uint8_t *dst, int dst_stride, const uint8_t *src,
             int src_stride, int w, int h, int interlaced, int tff, unsigned field,
             unsigned long long fieldno, int force_general) {
    const Variant v = pick_variant(p);
#define CVS_RUN(VHS, CD, OF) \
    return run_field<R, VHS, CD, OF, NF>(p, cur, dst, dst_stride, src, src_stride, w, h, interlaced, tff, field, fieldno, force_general)
    if (!v.vhs) { if (v.outfull) CVS_RUN(false, 9, true); else CVS_RUN(false, 9, false); }
    if (v.cd == 9) { if (v.outfull) CVS_RUN(true, 9, true); else CVS_RUN(true, 9, false); }
    if (v.cd == 12) { if (v.outfull) CVS_RUN(true, 12, true); else CVS_RUN(true, 12, false); }
    if (v.outfull) CVS_RUN(true, 14, true); else CVS_RUN(true, 14, false);
#undef CVS_RUN
}

}  // namespace

extern "C" {

// tests: a warm-up shorter than kWarmPx makes the first attempt fail on most rows (second-chance path)
void emu_set_warm_px(int px) { g_warm_px = (px >= 0 && px <= kWarmPx) ? px : kWarmPx; }

// precision: 0 = float (production arithmetic), 1 = double (reference arithmetic).
// rng_pos in/out: absolute rand() position (draws since the default seed).
// noise_fast: CVS_NOISE_FAST as the library applies it (float only, small amplitudes only).
int emu_composite_layer_ex(const cvs_params *p, int precision, unsigned long long *rng_pos,
                           uint8_t *dst, int dst_stride, const uint8_t *src, int src_stride,
                           int w, int h, int interlaced, int tff, unsigned field, unsigned long long fieldno,
                           int force_general, int noise_fast) {
    if (!p || !dst || !src || w <= 0 || h <= 0 || dst_stride < 4 * w || src_stride < 4 * w) return CVS_ERR_INVALID_ARG;
    static RandCursor cur;
    cur.seek(*rng_pos);
    int rc;
    const bool nf = noise_fast && !precision && p->video_noise + 2 * p->video_chroma_noise <= 96;
    if (precision) rc = dispatch<double, false>(*p, cur, dst, dst_stride, src, src_stride, w, h, interlaced, tff, field, fieldno, force_general);
    else if (nf) rc = dispatch<float, true>(*p, cur, dst, dst_stride, src, src_stride, w, h, interlaced, tff, field, fieldno, force_general);
    else rc = dispatch<float, false>(*p, cur, dst, dst_stride, src, src_stride, w, h, interlaced, tff, field, fieldno, force_general);
    *rng_pos = cur.pos();
    return rc;
}
int emu_composite_layer(const cvs_params *p, int precision, unsigned long long *rng_pos,
                        uint8_t *dst, int dst_stride, const uint8_t *src, int src_stride,
                        int w, int h, int interlaced, int tff, unsigned field, unsigned long long fieldno,
                        int force_general) {
    return emu_composite_layer_ex(p, precision, rng_pos, dst, dst_stride, src, src_stride, w, h, interlaced, tff, field,
                                  fieldno, force_general, 0);
}

// ---- small probes of the product's host-side code (glibc_rand.cpp, field_plan.cpp) -------------------

// (unsigned)rand() value number `pos` (0-based) of the default-seeded stream, by jump-ahead
unsigned emu_rand_at(unsigned long long pos) {
    RandCursor c;
    c.seek(pos);
    return c.next();
}

// sequential draws from position `pos` into out[n]
void emu_rand_run(unsigned long long pos, int n, unsigned *out) {
    RandCursor c;
    c.seek(pos);
    for (int i = 0; i < n; i++) out[i] = c.next();
}

// returns the number of n in the probe set for which the multiply-shift reduction disagrees with n % m
int emu_mod_magic_mismatches(unsigned m) {
    uint32_t magic, shift;
    mod_magic(m, magic, shift);
    int bad = 0;
    auto check = [&](uint32_t raw) {
        const uint32_t n = raw >> 1;
        if ((uint32_t)draw_mod(raw, m, magic, shift) != n % m) bad++;
    };
    for (uint32_t i = 0; i < 200000; i++) {
        check(i);
        check(0xFFFFFFFFu - i);
        check(i * 2654435761u);
        check((uint32_t)(((uint64_t)i * m) << 1));          // multiples of m
        check((uint32_t)((((uint64_t)i * m) << 1) - 2));    // just below a multiple
    }
    return bad;
}

// head-switch schedule of one field at absolute rand() position pos: returns hs_count, fills first/shifts
int emu_head_switch(const cvs_params *p, int w, int h, unsigned field, unsigned long long pos, int *first,
                    int *shifts, int cap) {
    GeomPlan g;
    build_geom_plan(*p, w, h, field, g);
    RandCursor c;
    c.seek(pos);
    FieldSide fs;
    build_field_side(*p, g, c, fs);
    *first = fs.hs_first;
    for (int i = 0; i < fs.hs_count && i < cap; i++) shifts[i] = fs.hs_shift[(size_t)i];
    return fs.hs_count;
}

}  // extern "C"
