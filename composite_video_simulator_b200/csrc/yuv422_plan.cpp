// yuv422_plan.cpp -- see yuv422_plan.h.
#include "yuv422_plan.h"

#include <cmath>
#include <cstring>

namespace cvs422 {

using cvs::RandCursor;
using cvs::RandPoly;
using cvs::rand_poly_mul;
using cvs::rand_poly_one;
using cvs::rand_poly_xpow;

static double pole_alpha(double rate, double hz) {       // LowpassFilter::setFilter, ffmpeg_to_composite.cpp:104-109
    const double dt = 1.0 / rate;
    const double tau = 1 / (hz * 2 * M_PI);
    return dt / (tau + dt);
}

int make_k422(const cvs422_params &p, int w, int h, K422 &K, DivPair &dv, std::vector<double> &lut) {
    std::memset(&K, 0, sizeof(K));
    std::memset(&dv, 0, sizeof(dv));
    if (w <= 0 || h <= 0 || (w & 1)) return CVS_ERR_INVALID_ARG;     // the reference writes past the row for odd widths
    if (p.video_noise < 0 || p.video_chroma_noise < 0 || p.video_chroma_phase_noise < 0) return CVS_ERR_INVALID_ARG;
    if (p.video_noise > 32767 || p.video_chroma_noise > 32767 || p.video_chroma_phase_noise > 32767) return CVS_ERR_INVALID_ARG;
    const double rate_y = (315000000.00 * 4) / 88, rate_c = (315000000.00 * 4) / (88 * 2);
    const bool ntsc = p.output_ntsc != 0;
    uint32_t f = 0;
    // input lowpass (:366-383): NTSC U 1.3 MHz delay 2, V 0.6 MHz delay 4; PAL both 1.3 MHz delay 2
    const double cut_u = 1300000, cut_v = ntsc ? 600000 : 1300000;
    K.a_in[0] = pole_alpha(rate_c, cut_u);  K.a_inhp[0] = pole_alpha(rate_c, cut_u / 2);
    K.a_in[1] = pole_alpha(rate_c, cut_v);  K.a_inhp[1] = pole_alpha(rate_c, cut_v / 2);
    K.d_in[0] = 2;
    K.d_in[1] = ntsc ? 4 : 2;
    if (p.composite_in_chroma_lowpass) f |= G_IN_LP;
    if (p.composite_out_chroma_lowpass) {                              // (:948-951): full wins over lite
        f |= G_OUT_FULL;
        K.a_out[0] = K.a_in[0]; K.a_outhp[0] = K.a_inhp[0]; K.d_out[0] = K.d_in[0];
        K.a_out[1] = K.a_in[1]; K.a_outhp[1] = K.a_inhp[1]; K.d_out[1] = K.d_in[1];
    } else if (p.composite_out_chroma_lowpass_lite) {                  // (:407-416)
        f |= G_OUT_LITE;
        K.a_out[0] = K.a_out[1] = pole_alpha(rate_c, (315000000.00 * 4) / (88 * 2 * 4));
        K.d_out[0] = K.d_out[1] = 1;
    }
    if (p.composite_preemphasis != 0 && p.composite_preemphasis_cut > 0) {   // (:636)
        f |= G_PREEMPH;
        K.a_pre = pole_alpha(rate_y, p.composite_preemphasis_cut);
    }
    K.preemph = p.composite_preemphasis;
    double luma_cut = 2400000, chroma_cut = 320000;                    // (:789-807)
    K.cd = 4;
    if (p.output_vhs_tape_speed == CVS_VHS_LP) { luma_cut = 1900000; chroma_cut = 300000; K.cd = 5; }
    if (p.output_vhs_tape_speed == CVS_VHS_EP) { luma_cut = 1400000; chroma_cut = 280000; K.cd = 6; }
    K.a_luma = pole_alpha(rate_y, luma_cut);
    K.a_lsharp = pole_alpha(rate_y, luma_cut * 2);
    K.a_ch = pole_alpha(rate_c, chroma_cut);
    K.a_csharp = pole_alpha(rate_c, chroma_cut * 2);
    K.sharpen = p.vhs_out_sharpen;
    K.sharpen_c = p.vhs_out_sharpen_chroma;
    K.lagV = (K.cd > 4) ? 2 : 1;
    if (p.nocolor_subcarrier) f |= G_NOCOLOR;
    if (p.nocolor_subcarrier_after_yc_sep) f |= G_NOCOLOR_YC;
    if (p.emulating_vhs) f |= G_VHS;
    if (p.vhs_chroma_vert_blend && ntsc) f |= G_VBLEND;               // (:858)
    if (p.vhs_svideo_out) f |= G_SVIDEO;
    if (p.video_chroma_phase_noise != 0) f |= G_PHASE;
    K.amp = p.subcarrier_amplitude;
    K.amp_back = p.subcarrier_amplitude_back;
    K.recombine = p.video_yc_recombine > 0 ? p.video_yc_recombine : 0;
    // the interior (fast) variant of the kernel only knows the common switches
    if ((f & (G_PREEMPH | G_NOCOLOR | G_NOCOLOR_YC)) || K.amp != 50 || K.amp_back != 50 || K.recombine != 0) f |= G_GENERAL;
    K.flags = f;
    if (K.recombine > kMaxRecombine) return CVS_ERR_UNSUPPORTED;
    // divisors of the demodulators (:533): amp_back for the first, amp for the VHS / -yc-recomb ones
    const bool first_demod = !p.nocolor_subcarrier, later_demod = (p.emulating_vhs && !p.vhs_svideo_out) || K.recombine > 0;
    if ((first_demod && K.amp_back <= 0) || (later_demod && K.amp <= 0)) return CVS_ERR_INVALID_ARG;
    if (K.amp < -100000 || K.amp > 100000) return CVS_ERR_INVALID_ARG;
    if (K.amp_back > 0) cvs::mod_magic((uint32_t)K.amp_back, dv.back_magic, dv.back_shift);
    if (K.amp > 0) cvs::mod_magic((uint32_t)K.amp, dv.amp_magic, dv.amp_shift);
    K.vnoise = p.video_noise;
    K.cnoise = p.video_chroma_noise;
    K.pnoise = p.video_chroma_phase_noise;
    if (K.vnoise != 0) cvs::mod_magic((uint32_t)(2 * K.vnoise + 1), K.vmagic, K.vshift);
    if (K.cnoise != 0) cvs::mod_magic((uint32_t)(2 * K.cnoise + 1), K.cmagic, K.cshift);
    K.ntsc = ntsc;
    K.phase_shift = p.video_scanline_phase_shift;
    K.phase_offset = p.video_scanline_phase_shift_offset;
    K.w = w;
    K.h = h;
    K.cw = w / 2;
    lut.clear();
    for (int st = -K.pnoise; st <= K.pnoise; st++) {                   // (:765): libm on the host, as the reference
        const double pi = ((double)st * M_PI) / 100;
        lut.push_back(std::cos(pi));
        lut.push_back(std::sin(pi));
    }
    if (K.pnoise != 0 && K.pnoise <= kPhaseMapMax) {
        // The rotation's input is an 8-bit sample: tabulate its output byte for every state, U then V, with the
        // reference's operations in the reference's order (:772-779; this file is compiled with -ffp-contract=off)
        const size_t nst = (size_t)(2 * K.pnoise + 1);
        std::vector<uint8_t> maps(nst * 512);
        for (size_t i = 0; i < nst; i++) {
            const double c = lut[2 * i], sn = lut[2 * i + 1];
            for (int b = 0; b < 256; b++) {
                const double x = (double)(b - 128);
                const double xc = x * c, xs = x * sn;
                const double u_ = xc - xs, v_ = xc + xs;
                const double uo = u_ + 128.0, vo = v_ + 128.0;
                const int ui = (int)uo, vi = (int)vo;
                maps[i * 512 + (size_t)b] = (uint8_t)(ui < 0 ? 0 : (ui > 255 ? 255 : ui));
                maps[i * 512 + 256 + (size_t)b] = (uint8_t)(vi < 0 ? 0 : (vi > 255 ? 255 : vi));
            }
        }
        const size_t nd = lut.size();
        lut.resize(nd + maps.size() / sizeof(double));
        std::memcpy(lut.data() + nd, maps.data(), maps.size());
        K.flags |= G_PHASE_MAP;
    }
    K.phase_lut = lut.data();
    return CVS_OK;
}

void build_geom_plan(const cvs422_params &p, int w, int h, unsigned field, cvs::GeomPlan &g) {
    g.w = w;
    g.h = h;
    g.field = field;
    g.nl = (h > (int)field) ? (h - (int)field + 1) / 2 : 0;
    const uint64_t cw = (uint64_t)(w / 2), nl = (uint64_t)g.nl;
    g.has_luma = p.video_noise != 0;
    g.has_hs = p.vhs_head_switching && p.vhs_head_switching_phase_noise != 0;
    g.has_chroma = p.video_chroma_noise != 0;
    g.has_phase = p.video_chroma_phase_noise != 0;
    g.has_loss = p.video_chroma_loss != 0;
    g.offL = 0;
    g.offH = g.offL + (g.has_luma ? nl * (uint64_t)w : 0);
    g.offC = g.offH + (g.has_hs ? 4 : 0);
    g.offP = g.offC + (g.has_chroma ? nl * 2 * cw : 0);
    g.offD = g.offP + (g.has_phase ? nl : 0);
    g.ndraws = g.offD + (g.has_loss ? nl : 0);
    g.jumpH = rand_poly_xpow(g.offH);
    g.jumpP = rand_poly_xpow(g.offP);
    g.jumpD = rand_poly_xpow(g.offD);
    g.jumpN = rand_poly_xpow(g.ndraws);

    g.seek.assign((size_t)g.nl * 62, 0);
    const RandPoly stepL = rand_poly_xpow((uint64_t)w), stepC = rand_poly_xpow(2 * cw);
    RandPoly curL = rand_poly_one(), curC = rand_poly_one();
    bool steadyL = false, steadyC = false;
    for (int r = 0; r < g.nl; r++) {
        const uint64_t warmL = (uint64_t)warm_samples((long long)r * w);            // luma draws before the row
        const uint64_t warmC = (uint64_t)warm_samples((long long)r * (long long)cw); // chroma SAMPLES before the row (2 draws each)
        const uint64_t posL = g.offL + (uint64_t)r * (uint64_t)w - warmL;
        const uint64_t posC = g.offC + (uint64_t)r * 2 * cw - 2 * warmC;
        if (warmL == (uint64_t)kWarm && steadyL) curL = rand_poly_mul(curL, stepL);
        else { curL = rand_poly_xpow(posL); steadyL = (warmL == (uint64_t)kWarm); }
        if (warmC == (uint64_t)kWarm && steadyC) curC = rand_poly_mul(curC, stepC);
        else { curC = rand_poly_xpow(posC); steadyC = (warmC == (uint64_t)kWarm); }
        std::memcpy(&g.seek[(size_t)r * 62], curL.c, sizeof(curL.c));
        std::memcpy(&g.seek[(size_t)r * 62 + 31], curC.c, sizeof(curC.c));
    }
}

void build_field_side_at(const cvs422_params &p, const cvs::GeomPlan &g, const RandCursor &cur, cvs::FieldSide &fs) {
    cur.window(fs.window);
    fs.rowinfo.assign((size_t)g.nl, 0);
    fs.hs_first = 0;
    fs.hs_count = 0;
    fs.hs_shift.clear();
    const int w = g.w, h = g.h;
    const unsigned field = g.field;
    if (p.vhs_head_switching) {                                        // (:668-733)
        const unsigned twidth = (unsigned)w + (unsigned)w / 10;
        double noise = 0;
        if (g.has_hs) {
            RandCursor c = cur;
            c.jump(g.jumpH, g.offH);
            unsigned v = c.next() * c.next() * c.next() * c.next();
            v %= 2000000000U;
            noise = ((double)v / 1000000000U) - 1.0;
            noise *= p.vhs_head_switching_phase_noise;
        }
        const double t = p.output_ntsc ? twidth * 262.5 : twidth * 312.5;
        const unsigned pp = (unsigned)(std::fmod(p.vhs_head_switching_phase + noise, 1.0) * t);
        const unsigned hx = pp % twidth;
        int y = (int)(((pp / twidth) * 2) + field);
        y -= p.output_ntsc ? (262 - 240) * 2 : (312 - 288) * 2;
        const int ishif = (hx >= twidth / 2) ? (int)hx - (int)twidth : (int)hx;
        int shif = 0, shy = 0;
        std::vector<int> rows, shifts;
        while (y < h) {
            // the row the switch falls on has shif == 0 and is left alone; its start column is never used
            if (y >= 0 && shif != 0) {
                rows.push_back((y - (int)field) / 2);
                shifts.push_back(shif);
            } else if (shy > 1 && shif == 0) {
                break;
            }
            shif = (shy == 0) ? ishif : (shif * 7) / 8;
            y += 2;
            shy++;
        }
        // a shift of -d with d <= w/10 is a delay by d with fill 16 (the wrapped-in samples are padding):
        // done in the ring; anything else goes through the pre-pass
        bool inline_ok = true;
        for (int sh : shifts)
            if (!(sh < 0 && -sh <= kHsMaxDelay && -sh <= w / 10 && -sh <= 255)) inline_ok = false;
        for (size_t i = 0; i < rows.size(); i++) {
            if (inline_ok) {
                fs.rowinfo[(size_t)rows[i]] |= ((uint32_t)RG_HEADSW << 16) | ((uint32_t)(-shifts[i]) << 24);
            } else {
                if (fs.hs_count == 0) fs.hs_first = rows[i];
                fs.hs_shift.push_back(shifts[i]);
                fs.hs_count++;
                fs.rowinfo[(size_t)rows[i]] |= (uint32_t)RG_HEADSW_PRE << 16;
            }
        }
    }
    if (g.has_phase) {                                                 // (:755-765)
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
    if (g.has_loss) {                                                  // (:932-941)
        RandCursor c = cur;
        c.jump(g.jumpD, g.offD);
        for (int r = 0; r < g.nl; r++)
            if ((c.next() % 100000U) < (unsigned)p.video_chroma_loss) fs.rowinfo[(size_t)r] |= (uint32_t)RG_DROPOUT << 16;
    }
}

}  // namespace cvs422
