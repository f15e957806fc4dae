// field_plan.cpp -- see field_plan.h.
#include "field_plan.h"

#include <cmath>
#include <cstring>

namespace cvs {

void mod_magic(uint32_t m, uint32_t &magic, uint32_t &shift) {
    // k = 32 + floor(log2 m), magic = ceil(2^k / m).  For n < 2^31:
    // n*(magic*m - 2^k) < 2^31 * m <= 2^k  (2^floor(log2 m) > m/2)  => floor(n*magic / 2^k) == floor(n/m).
    uint32_t fl = 0;
    while ((2u << fl) <= m) fl++;
    const unsigned k = 32 + fl;
    const unsigned __int128 one = 1;
    const unsigned __int128 v = ((one << k) + m - 1) / m;
    magic = (uint32_t)v;
    shift = fl;
}

Variant pick_variant(const cvs_params &p) {
    Variant v;
    v.vhs = p.emulating_vhs != 0;
    v.cd = 9;
    if (v.vhs) {                                   // ffmpeg_ntsc.cpp:1773-1791
        switch (p.output_vhs_tape_speed) {
        case CVS_VHS_LP: v.cd = 12; break;
        case CVS_VHS_EP: v.cd = 14; break;
        default: v.cd = 9; break;
        }
    }
    v.outfull = p.composite_out_chroma_lowpass && !p.composite_out_chroma_lowpass_lite;   // :1903-1908
    return v;
}

static double pole_alpha(double hz) {              // LowpassFilter::setFilter, ffmpeg_ntsc.cpp:78-86
    const double rate = (315000000.00 * 4) / 88;
    const double timeInterval = 1.0 / rate;
    const double tau = 1 / (hz * 2 * M_PI);
    return timeInterval / (tau + timeInterval);
}

template <typename R>
void make_kconst(const cvs_params &p, int w, int h, bool outfull, KConst<R> &K, std::vector<R> &lut) {
    std::memset(&K, 0, sizeof(K));
    auto set = [](R &a, R &b, R &c, double alpha) { a = (R)alpha; b = (R)(1.0 - alpha); c = (R)(alpha * alpha * alpha); };
    R unused_c;
    set(K.a_in.x, K.b_in.x, K.c_in.x, pole_alpha(1300000));   // I, :1442
    set(K.a_in.y, K.b_in.y, K.c_in.y, pole_alpha(600000));     // Q
    const bool pre = p.composite_preemphasis != 0 && p.composite_preemphasis_cut > 0;   // :1614
    if (pre) set(K.a_pre, K.b_pre, unused_c, pole_alpha(p.composite_preemphasis_cut));
    K.preemph = (R)p.composite_preemphasis;
    double luma_cut = 2400000, chroma_cut = 320000;         // :1773-1791
    if (p.output_vhs_tape_speed == CVS_VHS_LP) { luma_cut = 1900000; chroma_cut = 300000; }
    if (p.output_vhs_tape_speed == CVS_VHS_EP) { luma_cut = 1400000; chroma_cut = 280000; }
    set(K.a_luma, K.b_luma, K.c_luma, pole_alpha(luma_cut));
    set(K.a_ch.x, K.b_ch.x, K.c_ch.x, pole_alpha(chroma_cut));
    K.a_ch.y = K.a_ch.x; K.b_ch.y = K.b_ch.x; K.c_ch.y = K.c_ch.x;
    set(K.a_sharp, K.b_sharp, K.c_sharp, pole_alpha(luma_cut * 4));    // :1874
    K.sharpen = (R)p.vhs_out_sharpen;
    {   // merged fp32 forms (lane_pipeline.cuh, Num<float>::vhs_luma / rgb2yiq2); products formed in double
        const double al = pole_alpha(luma_cut), as = pole_alpha(luma_cut * 4), g = p.vhs_out_sharpen;
        K.k_boost_r = (R)(2.6 * al * al * al);
        K.k_boost_p = (R)(-1.6 * al * al * al * al);
        K.k_sharp_y = (R)(1.0 + 2.0 * g);
        K.k_sharp_t = (R)(-2.0 * g * as * as * as);
        // I = -.27 (b - dY) + .74 (r - dY), Q = .41 (b - dY) + .48 (r - dY), dY = .30 r + .59 g + .11 b   (:1377-1382)
        const double ky[3] = {0.30, 0.59, 0.11};
        const double ci[3] = {0.74 - 0.47 * ky[0], -0.47 * ky[1], -0.27 - 0.47 * ky[2]};
        const double cq[3] = {0.48 - 0.89 * ky[0], -0.89 * ky[1], 0.41 - 0.89 * ky[2]};
        K.k_iq_r.x = (R)(256 * ci[0]); K.k_iq_g.x = (R)(256 * ci[1]); K.k_iq_b.x = (R)(256 * ci[2]);
        K.k_iq_r.y = (R)(256 * cq[0]); K.k_iq_g.y = (R)(256 * cq[1]); K.k_iq_b.y = (R)(256 * cq[2]);
    }
    if (outfull) {                                          // composite_lowpass, :1442
        set(K.a_out.x, K.b_out.x, K.c_out.x, pole_alpha(1300000));
        set(K.a_out.y, K.b_out.y, K.c_out.y, pole_alpha(600000));
    } else {                                                // composite_lowpass_tv, :1411
        set(K.a_out.x, K.b_out.x, K.c_out.x, pole_alpha(2600000));
        set(K.a_out.y, K.b_out.y, K.c_out.y, pole_alpha(2600000));
    }
    uint32_t f = 0;
    if (p.composite_in_chroma_lowpass) f |= F_IN_LP;
    if (p.composite_out_chroma_lowpass) f |= F_OUT_LP;
    if (pre) f |= F_PREEMPH;
    if (p.nocolor_subcarrier) f |= F_NOCOLOR;
    if (p.vhs_chroma_vert_blend && p.output_ntsc) f |= F_VBLEND;    // :1843
    if (p.vhs_svideo_out) f |= F_SVIDEO;
    if (p.video_chroma_phase_noise != 0) f |= F_PHASE;
    if (!(f & F_IN_LP) || !(f & F_OUT_LP) || (f & (F_PREEMPH | F_NOCOLOR | F_SVIDEO)) ||
        p.subcarrier_amplitude != 50 || p.subcarrier_amplitude_back != 50)
        f |= F_GENERAL;
    // the fast path also assumes each noise source is in its preset state (luma noise on; chroma
    // noise and phase noise on exactly when emulating VHS); anything else takes the general path
    const bool vhs = p.emulating_vhs != 0;
    if (p.video_noise == 0 || (p.video_chroma_noise != 0) != vhs || (p.video_chroma_phase_noise != 0) != vhs)
        f |= F_GENERAL;
    K.flags = f;
    K.pnoise = p.video_chroma_phase_noise < 0 ? -p.video_chroma_phase_noise : p.video_chroma_phase_noise;
    K.phase_shift = p.video_scanline_phase_shift;
    K.phase_offset = p.video_scanline_phase_shift_offset;
    K.amp = p.subcarrier_amplitude;
    K.amp_back = p.subcarrier_amplitude_back;
    K.vnoise = p.video_noise;
    K.cnoise = p.video_chroma_noise;
    if (K.vnoise != 0) mod_magic((uint32_t)(2 * K.vnoise + 1), K.vmagic, K.vshift);
    if (K.cnoise != 0) mod_magic((uint32_t)(2 * K.cnoise + 1), K.cmagic, K.cshift);
    K.w = w;
    K.h = h;
    // {sin, cos}(state * pi / 100) for every reachable phase-noise state (:1746-1749); libm on the
    // host, in double, exactly as the reference evaluates them
    lut.clear();
    for (int st = -K.pnoise; st <= K.pnoise; st++) {
        const double pi = ((double)st * M_PI) / 100;
        lut.push_back((R)std::sin(pi));
        lut.push_back((R)std::cos(pi));
    }
    K.phase_lut = lut.data();
}
template void make_kconst<float>(const cvs_params &, int, int, bool, KConst<float> &, std::vector<float> &);
template void make_kconst<double>(const cvs_params &, int, int, bool, KConst<double> &, std::vector<double> &);

void build_geom_plan(const cvs_params &p, int w, int h, unsigned field, GeomPlan &g, int warm_px) {
    g.w = w;
    g.h = h;
    g.field = field;
    g.nl = (h > (int)field) ? (h - (int)field + 1) / 2 : 0;
    const uint64_t nlw = (uint64_t)g.nl * (uint64_t)w;
    g.has_luma = p.video_noise != 0;                                                   // :1632
    g.has_hs = p.vhs_head_switching && p.vhs_head_switching_phase_noise != 0;         // :1654
    g.has_chroma = p.video_chroma_noise != 0;                                          // :1719
    g.has_phase = p.video_chroma_phase_noise != 0;                                     // :1736
    g.has_loss = p.video_chroma_loss != 0;                                             // :1891
    g.offL = 0;
    g.offH = g.offL + (g.has_luma ? nlw : 0);
    g.offC = g.offH + (g.has_hs ? 4 : 0);
    g.offP = g.offC + (g.has_chroma ? 2 * nlw : 0);
    g.offD = g.offP + (g.has_phase ? (uint64_t)g.nl : 0);
    g.ndraws = g.offD + (g.has_loss ? (uint64_t)g.nl : 0);
    g.jumpH = rand_poly_xpow(g.offH);
    g.jumpP = rand_poly_xpow(g.offP);
    g.jumpD = rand_poly_xpow(g.offD);
    g.jumpN = rand_poly_xpow(g.ndraws);

    g.seek.assign((size_t)g.nl * 62, 0);
    const RandPoly stepL = rand_poly_xpow((uint64_t)w), stepC = rand_poly_xpow(2ull * (uint64_t)w);
    RandPoly curL = rand_poly_one(), curC = rand_poly_one();
    bool steadyL = false, steadyC = false;
    for (int r = 0; r < g.nl; r++) {
        const uint64_t warm = (uint64_t)warm_draws_luma(r, w, warm_px);
        const uint64_t posL = g.offL + (uint64_t)r * w - warm;
        const uint64_t posC = g.offC + 2ull * (uint64_t)r * w - 2 * warm;
        // once the warm-up is at full length consecutive rows are a constant jump apart
        if (warm == (uint64_t)warm_px && steadyL) curL = rand_poly_mul(curL, stepL);
        else { curL = rand_poly_xpow(posL); steadyL = (warm == (uint64_t)warm_px); }
        if (warm == (uint64_t)warm_px && steadyC) curC = rand_poly_mul(curC, stepC);
        else { curC = rand_poly_xpow(posC); steadyC = (warm == (uint64_t)warm_px); }
        std::memcpy(&g.seek[(size_t)r * 62], curL.c, sizeof(curL.c));
        std::memcpy(&g.seek[(size_t)r * 62 + 31], curC.c, sizeof(curC.c));
    }
}

void build_field_side(const cvs_params &p, const GeomPlan &g, RandCursor &cur, FieldSide &fs) {
    build_field_side_at(p, g, cur, fs);
    cur.jump(g.jumpN, g.ndraws);
}

void build_field_side_at(const cvs_params &p, const GeomPlan &g, const RandCursor &cur, FieldSide &fs) {
    cur.window(fs.window);
    fs.rowinfo.assign((size_t)g.nl, 0);
    fs.hs_first = 0;
    fs.hs_count = 0;
    fs.hs_shift.clear();
    const int w = g.w, h = g.h;
    const unsigned field = g.field;

    // VHS head switching, ffmpeg_ntsc.cpp:1646-1713
    if (p.vhs_head_switching) {
        const unsigned twidth = (unsigned)w + (unsigned)w / 10;
        double noise = 0;
        if (g.has_hs) {
            RandCursor c = cur;
            c.jump(g.jumpH, g.offH);
            unsigned v = c.next() * c.next() * c.next() * c.next();     // wraps mod 2^32, :1655
            v %= 2000000000U;
            noise = ((double)v / 1000000000U) - 1.0;
            noise *= p.vhs_head_switching_phase_noise;
        }
        const double t = p.output_ntsc ? twidth * 262.5 : twidth * 312.5;
        unsigned pp = (unsigned)(std::fmod(p.vhs_head_switching_point + noise, 1.0) * t);
        int y = (int)(((pp / twidth) * 2) + field);
        pp = (unsigned)(std::fmod(p.vhs_head_switching_phase + noise, 1.0) * t);
        const unsigned hx = pp % twidth;
        y -= p.output_ntsc ? (262 - 240) * 2 : (312 - 288) * 2;
        const int ishif = (hx >= twidth / 2) ? (int)hx - (int)twidth : (int)hx;
        int shif = 0, shy = 0;
        std::vector<int> rows, shifts;
        while (y < h) {
            // the first affected row has shif == 0, so its start column never matters (:1683-1711)
            if (y >= 0 && shif != 0) {
                rows.push_back((y - (int)field) / 2);
                shifts.push_back(shif);
            } else if (!rows.empty() && shif == 0) {
                break;              // decayed: no later row is shifted
            }
            shif = (shy == 0) ? ishif : (shif * 7) / 8;
            y += 2;
            shy++;
        }
        // A negative shift -d moves the row right by d; the pixels that wrap in come from
        // tmp[twidth - d + x], which is zero padding for every x when d <= w/10.  Then the rotation is
        // a plain delay and k_fields does it in shared memory; anything else goes through the pre-pass.
        // (only the VHS kernels carry the delay ring: -vhs-head-switching without -vhs takes the pre-pass)
        bool inline_ok = !rows.empty() && p.emulating_vhs != 0;
        for (int sh : shifts)
            if (!(sh < 0 && -sh <= kHsMaxDelay && -sh <= w / 10)) inline_ok = false;
        for (size_t i = 0; i < rows.size(); i++) {
            if (inline_ok) {
                fs.rowinfo[(size_t)rows[i]] |= ((uint32_t)RF_HEADSW_INLINE << 16) | ((uint32_t)(-shifts[i]) << 24);
            } else {
                if (fs.hs_count == 0) fs.hs_first = rows[i];       // rows are consecutive
                fs.hs_shift.push_back(shifts[i]);
                fs.hs_count++;
                fs.rowinfo[(size_t)rows[i]] |= (uint32_t)RF_HEADSW << 16;
            }
        }
    }

    // chroma phase noise: one draw per row, the state carries down the field (:1736-1746)
    if (g.has_phase) {
        RandCursor c = cur;
        c.jump(g.jumpP, g.offP);
        const int pn = p.video_chroma_phase_noise;
        const unsigned mod = (unsigned)(pn * 2 + 1);
        int noise = 0;
        for (int r = 0; r < g.nl; r++) {
            noise += (int)(c.next() % mod) - pn;
            noise /= 2;
            fs.rowinfo[(size_t)r] |= (uint32_t)noise & 0xFFFFu;
        }
    }
    // chroma dropout: one draw per row (:1891-1901)
    if (g.has_loss) {
        RandCursor c = cur;
        c.jump(g.jumpD, g.offD);
        for (int r = 0; r < g.nl; r++)
            if ((c.next() % 100000U) < (unsigned)p.video_chroma_loss) fs.rowinfo[(size_t)r] |= (uint32_t)RF_DROPOUT << 16;
    }
}

}  // namespace cvs
