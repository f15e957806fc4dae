/*
 * yuv422_oracle.c -- see yuv422_oracle.h.  TEST INFRASTRUCTURE ONLY.
 *
 * Every function cites the lines of /root/reference/ffmpeg_to_composite.cpp it restates.
 * Build with -ffp-contract=off (oracle/Makefile): the one-pole filters must not be contracted.
 */
#include "yuv422_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* one-pole filter, LowpassFilter :99-131 */
typedef struct { double alpha, prev; } pole;

static void pole_set(pole *f, double rate, double hz, double reset) {
    const double dt = 1.0 / rate;
    const double tau = 1 / (hz * 2 * M_PI);
    f->alpha = dt / (tau + dt);                                 /* :104-109 */
    f->prev = reset;
}
static double pole_lp(pole *f, double s) {                      /* :114-118 */
    const double a = s * f->alpha;
    const double b = f->prev - (f->prev * f->alpha);
    f->prev = a + b;
    return f->prev;
}
static double pole_hp(pole *f, double s) { return s - pole_lp(f, s); }   /* :119-123 */

static int clamp_u8(int x) { return x > 255 ? 255 : (x < 0 ? 0 : x); } /* clampu8 :335-342 */
static int q8(double s) { return clamp_u8((int)s); }            /* the implicit double -> int of every clampu8(s) call */

#define RATE_LUMA   ((315000000.00 * 4) / 88)
#define RATE_CHROMA ((315000000.00 * 4) / (88 * 2))

/* subcarrier phase index of a scanline, :449-459 (and again :511-525) */
static unsigned line_xi(const cvs422_params *p, unsigned long long fieldno, unsigned y) {
    if (!p->output_ntsc) return (unsigned)((fieldno + y) & 3);
    if (p->video_scanline_phase_shift == 90)
        return (unsigned)((fieldno + (unsigned long long)(long long)p->video_scanline_phase_shift_offset + (y >> 1)) & 3);
    if (p->video_scanline_phase_shift == 180)
        return (unsigned)((((fieldno + y) & 2) + (unsigned long long)(long long)p->video_scanline_phase_shift_offset) & 3);
    if (p->video_scanline_phase_shift == 270)
        return (unsigned)((fieldno + (unsigned long long)(long long)p->video_scanline_phase_shift_offset - (y >> 1)) & 3);
    return 0;
}

/* composite_video_chroma_lowpass, one row of one plane (:353-393).  plane 1 = U, 2 = V. */
static void row_chroma_lowpass(const cvs422_params *p, uint8_t *P, int cw, int plane) {
    pole lp[3], hp;
    double cutoff;
    int delay, x, f;
    if (p->output_ntsc) { cutoff = (plane == 1) ? 1300000 : 600000; delay = (plane == 1) ? 2 : 4; }   /* :366-370 */
    else { cutoff = 1300000; delay = 2; }                                                              /* :371-375 */
    pole_set(&hp, RATE_CHROMA, cutoff / 2, 128);                                                       /* :377-378 */
    for (f = 0; f < 3; f++) pole_set(&lp[f], RATE_CHROMA, cutoff, 128);
    for (x = 0; x < cw; x++) {
        double s = P[x];
        s += pole_hp(&hp, s);
        for (f = 0; f < 3; f++) s = pole_lp(&lp[f], s);
        if (x >= delay) P[x - delay] = (uint8_t)q8(s);
    }
}

/* composite_video_chroma_lowpass_lite, one row of one plane (:395-431) */
static void row_chroma_lowpass_lite(uint8_t *P, int cw) {
    pole lp[3];
    int x, f;
    for (f = 0; f < 3; f++) pole_set(&lp[f], RATE_CHROMA, (315000000.00 * 4) / (88 * 2 * 4), 128);
    for (x = 0; x < cw; x++) {
        double s = P[x];
        for (f = 0; f < 3; f++) s = pole_lp(&lp[f], s);
        if (x >= 1) P[x - 1] = (uint8_t)q8(s);
    }
}

/* composite_video_yuv_to_ntsc, one row (:434-477) */
static void row_modulate(const cvs422_params *p, uint8_t *Y, uint8_t *U, uint8_t *V, int w, unsigned xi, int amp) {
    static const int um[4] = { 1, 0, -1, 0 }, vm[4] = { 0, 1, 0, -1 };
    int x, sx;
    for (x = 0; x < w; x += 2) {
        const int c = x >> 1;
        for (sx = 0; sx < 2; sx++) {
            const unsigned ph = (xi + (unsigned)x + (unsigned)sx) & 3;
            int chroma = ((int)U[c] - 128) * amp * um[ph];
            chroma += ((int)V[c] - 128) * amp * vm[ph];
            Y[x + sx] = (uint8_t)clamp_u8(Y[x + sx] + (chroma / 50));
        }
        if (p->nocolor_subcarrier) U[c] = V[c] = 128;                       /* :473-474 */
    }
}

/* composite_ntsc_to_yuv, one row (:480-553).  Y has w+2 readable entries. */
static void row_demodulate(const cvs422_params *p, uint8_t *Y, uint8_t *U, uint8_t *V, int w, unsigned xi,
                           int amp_back, uint8_t *chroma /* w + 4 */) {
    unsigned d0 = 16, d1 = 16, d2, d3, sum = 16 * 2;
    int x;
    d2 = Y[0]; sum += d2;                                                     /* precharge, :491-492 */
    d3 = Y[1]; sum += d3;
    for (x = 0; x < w; x++) {
        const unsigned c = Y[x + 2];                                          /* unguarded read, :496 */
        sum -= d0;
        d0 = d1; d1 = d2; d2 = d3; d3 = c;
        sum += c;
        Y[x] = (uint8_t)(sum / 4);
        chroma[x] = (uint8_t)clamp_u8((int)c + 128 - (int)Y[x]);
        if (p->nocolor_subcarrier_after_yc_sep) {                             /* :504-508 */
            Y[x] = chroma[x];
            U[x / 2] = V[x / 2] = 128;
        }
    }
    if (p->nocolor_subcarrier_after_yc_sep) return;
    for (x = (int)((4 - xi) & 3); x < w; x += 4) {                            /* :527-530; writes past w are dead */
        if (x + 2 < w) chroma[x + 2] = (uint8_t)(255 - chroma[x + 2]);
        if (x + 3 < w) chroma[x + 3] = (uint8_t)(255 - chroma[x + 3]);
    }
    for (x = 0; x < w; x++)
        chroma[x] = (uint8_t)clamp_u8(((((int)chroma[x] - 128) * 50) / amp_back) + 128);   /* :532-534 */
    for (x = 0; x < w / 2; x++) {                                             /* :537-549 */
        if (xi & 1) { U[x] = (uint8_t)(255 - chroma[2 * x + 1]); V[x] = (uint8_t)(255 - chroma[2 * x]); }
        else        { U[x] = (uint8_t)(255 - chroma[2 * x]);     V[x] = (uint8_t)(255 - chroma[2 * x + 1]); }
    }
}

static void vhs_speed(const cvs422_params *p, double *luma_cut, double *chroma_cut, int *chroma_delay) {   /* :789-807 */
    switch (p->output_vhs_tape_speed) {
    case CVS_VHS_LP: *luma_cut = 1900000; *chroma_cut = 300000; *chroma_delay = 5; break;
    case CVS_VHS_EP: *luma_cut = 1400000; *chroma_cut = 280000; *chroma_delay = 6; break;
    default:         *luma_cut = 2400000; *chroma_cut = 320000; *chroma_delay = 4; break;
    }
}

static int field_rows(int h, unsigned field) { return (h > (int)field) ? (h - (int)field + 1) / 2 : 0; }

unsigned long long oracle422_draws_per_field(const cvs422_params *p, int w, int h, unsigned field) {
    const unsigned long long nl = (unsigned long long)field_rows(h, field);
    unsigned long long n = 0;
    if (p->video_noise != 0) n += nl * (unsigned long long)w;                                   /* :653-665 */
    if (p->vhs_head_switching && p->vhs_head_switching_phase_noise != 0) n += 4;                /* :675-676 */
    if (p->video_chroma_noise != 0) n += nl * 2ull * (unsigned long long)(w / 2);               /* :738-754 */
    if (p->video_chroma_phase_noise != 0) n += nl;                                              /* :755-764 */
    if (p->video_chroma_loss != 0) n += nl;                                                     /* :931-941 */
    return n;
}

int oracle422_composite_video_process(const cvs422_params *p, oracle_rng *g,
                                      uint8_t *yp, int ly, uint8_t *up, int lu, uint8_t *vp, int lv,
                                      int w, int h, unsigned field, unsigned long long fieldno) {
    const int nl = field_rows(h, field), cw = w / 2;
    const unsigned long long ndraw = oracle422_draws_per_field(p, w, h, field);
    uint32_t *draw;
    unsigned long long k, offL, offH, offC, offP, offD;
    uint8_t *Y, *YA, *U, *V, *ch, *prevU, *prevV;
    int r, x, f, i;
    int nY = 0, nU = 0, nV = 0, nP = 0;
    const int pre_on = p->composite_preemphasis != 0 && p->composite_preemphasis_cut > 0;       /* :636 */
    double luma_cut = 0, chroma_cut = 0;
    int chroma_delay = 0;
    /* head switch schedule */
    int hs_y0 = 0, hs_ishif = 0, hs_active = 0;
    const int twidth = w + w / 10;

    if (w <= 0 || h <= 0 || (w & 1) || field > 1) return -1;
    if (p->subcarrier_amplitude_back == 0 || (p->emulating_vhs && !p->vhs_svideo_out && p->subcarrier_amplitude == 0)) return -1;
    if (p->video_yc_recombine > 0 && p->subcarrier_amplitude == 0) return -1;

    /* the reference draws stage by stage over the whole field; take the field's draws up front and
       index them by stage offset (this is also the arithmetic the GPU planner uses) */
    draw = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(ndraw + 1));
    for (k = 0; k < ndraw; k++) draw[k] = oracle_rng_next(g);
    offL = 0;
    offH = offL + ((p->video_noise != 0) ? (unsigned long long)nl * w : 0);
    offC = offH + ((p->vhs_head_switching && p->vhs_head_switching_phase_noise != 0) ? 4 : 0);
    offP = offC + ((p->video_chroma_noise != 0) ? (unsigned long long)nl * 2 * cw : 0);
    offD = offP + ((p->video_chroma_phase_noise != 0) ? (unsigned long long)nl : 0);

    if (p->emulating_vhs) vhs_speed(p, &luma_cut, &chroma_cut, &chroma_delay);

    if (p->vhs_head_switching) {                                                                /* :668-733 */
        double noise = 0, t;
        unsigned pp, hx;
        if (p->vhs_head_switching_phase_noise != 0) {
            unsigned v = draw[offH] * draw[offH + 1] * draw[offH + 2] * draw[offH + 3];
            v %= 2000000000U;
            noise = ((double)v / 1000000000U) - 1.0;
            noise *= p->vhs_head_switching_phase_noise;
        }
        t = p->output_ntsc ? twidth * 262.5 : twidth * 312.5;
        pp = (unsigned)(fmod(p->vhs_head_switching_phase + noise, 1.0) * t);
        hx = pp % (unsigned)twidth;
        hs_y0 = (int)(((pp / (unsigned)twidth) * 2) + field);
        hs_y0 -= p->output_ntsc ? (262 - 240) * 2 : (312 - 288) * 2;
        hs_ishif = (hx >= (unsigned)twidth / 2) ? (int)hx - twidth : (int)hx;
        hs_active = 1;
    }

    Y = (uint8_t *)malloc((size_t)w + 8);
    YA = (uint8_t *)malloc((size_t)twidth + 8);
    U = (uint8_t *)malloc((size_t)cw + 8);
    V = (uint8_t *)malloc((size_t)cw + 8);
    ch = (uint8_t *)malloc((size_t)w + 8);
    prevU = (uint8_t *)malloc((size_t)cw + 8);
    prevV = (uint8_t *)malloc((size_t)cw + 8);
    memset(prevU, 128, (size_t)cw + 8);                                                         /* :869-870 */
    memset(prevV, 128, (size_t)cw + 8);

    for (r = 0; r < nl; r++) {
        const unsigned y = field + 2u * (unsigned)r;
        const unsigned xi = line_xi(p, fieldno, y);
        uint8_t *gy = yp + (size_t)y * (size_t)ly, *gu = up + (size_t)y * (size_t)lu, *gv = vp + (size_t)y * (size_t)lv;
        memcpy(Y, gy, (size_t)w);
        for (i = 0; i < 2; i++) {
            const long long off = (long long)y * ly + w + i;
            Y[w + i] = (off < (long long)ly * h) ? gy[w + i] : 0;
        }
        memcpy(U, gu, (size_t)cw);
        memcpy(V, gv, (size_t)cw);

        if (p->composite_in_chroma_lowpass) {                                                   /* :632 */
            row_chroma_lowpass(p, U, cw, 1);
            row_chroma_lowpass(p, V, cw, 2);
        }
        row_modulate(p, Y, U, V, w, xi, p->subcarrier_amplitude);                               /* :633 */

        if (pre_on) {                                                                           /* :636-650 */
            pole pre;
            pole_set(&pre, RATE_LUMA, p->composite_preemphasis_cut, 16);
            for (x = 0; x < w; x++) {
                double s = Y[x];
                s += pole_hp(&pre, s) * p->composite_preemphasis;
                Y[x] = (uint8_t)q8(s);
            }
        }
        if (p->video_noise != 0) {                                                              /* :653-665; nY carries down the field */
            const int vn = p->video_noise;
            for (x = 0; x < w; x++) {
                Y[x] = (uint8_t)clamp_u8(Y[x] + nY);
                nY += ((int)(draw[offL + (unsigned long long)r * w + x] % (unsigned)((vn * 2) + 1))) - vn;
                nY /= 2;
            }
        }
        if (hs_active && (int)y > hs_y0 && (int)y >= 0) {                                       /* :697-731 */
            /* rows below the switch row: shift schedule ishif, then *7/8 per field row */
            int n = ((int)y - hs_y0) / 2, shif = hs_ishif, j;
            for (j = 1; j < n; j++) shif = (shif * 7) / 8;
            if (shif != 0) {
                unsigned x2 = (unsigned)(twidth + shif) % (unsigned)twidth;                     /* tx == 0 on every shifted row */
                memset(YA, 16, (size_t)twidth);
                memcpy(YA, Y, (size_t)w);
                for (x = 0; x < w; x++) {
                    Y[x] = YA[x2];
                    if (++x2 == (unsigned)twidth) x2 = 0;
                }
            }
        }
        if (!p->nocolor_subcarrier)                                                             /* :735-736 */
            row_demodulate(p, Y, U, V, w, xi, p->subcarrier_amplitude_back, ch);

        if (p->video_chroma_noise != 0) {                                                       /* :738-754 */
            const int cn = p->video_chroma_noise;
            const unsigned m = (unsigned)((cn * 2) + 1);
            for (x = 0; x < cw; x++) {
                const unsigned long long at = offC + ((unsigned long long)r * cw + (unsigned long long)x) * 2;
                U[x] = (uint8_t)clamp_u8(U[x] + nU);
                V[x] = (uint8_t)clamp_u8(V[x] + nV);
                nU += ((int)(draw[at] % m)) - cn;
                nU /= 2;
                nV += ((int)(draw[at + 1] % m)) - cn;
                nV /= 2;
            }
        }
        if (p->video_chroma_phase_noise != 0) {                                                 /* :755-783 */
            const int pn = p->video_chroma_phase_noise;
            double pi;
            nP += ((int)(draw[offP + (unsigned long long)r] % (unsigned)((pn * 2) + 1))) - pn;
            nP /= 2;
            pi = ((double)nP * M_PI) / 100;
            for (x = 0; x < cw; x++) {
                const double u = (int)U[x] - 128, v = (int)V[x] - 128;
                const double u_ = (u * cos(pi)) - (u * sin(pi));                                /* sic, :772-773 */
                const double v_ = (v * cos(pi)) + (v * sin(pi));
                U[x] = (uint8_t)q8(u_ + 128);
                V[x] = (uint8_t)q8(v_ + 128);
            }
        }

        if (p->emulating_vhs) {                                                                 /* :788-929 */
            pole lp[3], pre, lu3[3], lv3[3];
            for (f = 0; f < 3; f++) pole_set(&lp[f], RATE_LUMA, luma_cut, 16);                  /* :810-827 */
            pole_set(&pre, RATE_LUMA, luma_cut, 16);
            for (x = 0; x < w; x++) {
                double s = Y[x];
                for (f = 0; f < 3; f++) s = pole_lp(&lp[f], s);
                s += pole_hp(&pre, s) * 1.6;
                Y[x] = (uint8_t)q8(s);
            }
            for (f = 0; f < 3; f++) {                                                           /* :830-851 */
                pole_set(&lu3[f], RATE_CHROMA, chroma_cut, 128);
                pole_set(&lv3[f], RATE_CHROMA, chroma_cut, 128);
            }
            for (x = 0; x < cw; x++) {
                double s = U[x];
                for (f = 0; f < 3; f++) s = pole_lp(&lu3[f], s);
                if (x >= chroma_delay) U[x - chroma_delay] = (uint8_t)q8(s);
                s = V[x];
                for (f = 0; f < 3; f++) s = pole_lp(&lv3[f], s);
                if (x >= chroma_delay) V[x - chroma_delay] = (uint8_t)q8(s);
            }
            if (p->vhs_chroma_vert_blend && p->output_ntsc && r >= 1) {                         /* :858-883: starts at field+2 */
                for (x = 0; x < cw; x++) {
                    const uint8_t cu = U[x], cv = V[x];
                    U[x] = (uint8_t)((prevU[x] + cu + 1) >> 1);
                    V[x] = (uint8_t)((prevV[x] + cv + 1) >> 1);
                    prevU[x] = cu;
                    prevV[x] = cv;
                }
            }
            for (f = 0; f < 3; f++) pole_set(&lp[f], RATE_LUMA, luma_cut * 2, 16);              /* :888-901 */
            for (x = 0; x < w; x++) {
                double s, ts;
                s = ts = Y[x];
                for (f = 0; f < 3; f++) ts = pole_lp(&lp[f], ts);
                Y[x] = (uint8_t)q8(s + ((s - ts) * p->vhs_out_sharpen));
            }
            for (f = 0; f < 3; f++) {                                                           /* :904-925 */
                pole_set(&lu3[f], RATE_CHROMA, chroma_cut * 2, 128);
                pole_set(&lv3[f], RATE_CHROMA, chroma_cut * 2, 128);
            }
            for (x = 0; x < cw; x++) {
                double s, ts;
                s = ts = U[x];
                for (f = 0; f < 3; f++) ts = pole_lp(&lu3[f], ts);
                U[x] = (uint8_t)q8(s + ((s - ts) * p->vhs_out_sharpen_chroma));
                s = ts = V[x];
                for (f = 0; f < 3; f++) ts = pole_lp(&lv3[f], ts);
                V[x] = (uint8_t)q8(s + ((s - ts) * p->vhs_out_sharpen_chroma));
            }
            if (!p->vhs_svideo_out) {                                                           /* :927-930 */
                row_modulate(p, Y, U, V, w, xi, p->subcarrier_amplitude);
                row_demodulate(p, Y, U, V, w, xi, p->subcarrier_amplitude, ch);
            }
        }

        if (p->video_chroma_loss != 0) {                                                        /* :932-941 */
            if ((draw[offD + (unsigned long long)r] % 100000) < (unsigned)p->video_chroma_loss) {
                memset(U, 128, (size_t)cw);
                memset(V, 128, (size_t)cw);
            }
        }
        for (i = 0; i < p->video_yc_recombine; i++) {                                           /* :943-946 */
            row_modulate(p, Y, U, V, w, xi, p->subcarrier_amplitude);
            row_demodulate(p, Y, U, V, w, xi, p->subcarrier_amplitude, ch);
        }
        if (p->composite_out_chroma_lowpass) {                                                  /* :948-951 */
            row_chroma_lowpass(p, U, cw, 1);
            row_chroma_lowpass(p, V, cw, 2);
        } else if (p->composite_out_chroma_lowpass_lite) {
            row_chroma_lowpass_lite(U, cw);
            row_chroma_lowpass_lite(V, cw);
        }

        memcpy(gy, Y, (size_t)w);
        memcpy(gu, U, (size_t)cw);
        memcpy(gv, V, (size_t)cw);
    }

    free(draw); free(Y); free(YA); free(U); free(V); free(ch); free(prevU); free(prevV);
    return 0;
}

/* render_field(), :1001-1129 */
void oracle422_render_field(uint8_t *const dst[3], const int dst_linesize[3], int dst_h,
                            const uint8_t *const src[3], const int src_linesize[3], int src_h,
                            const int row_bytes[3], int src_is_420,
                            int src_interlaced, int src_top_field_first, int second_field, unsigned field) {
    const unsigned chroma_h = src_is_420 ? (unsigned)src_h >> 1 : (unsigned)src_h;             /* :1005-1008 */
    unsigned y;
    for (y = field; y < (unsigned)dst_h; y += 2) {
        unsigned sy = (y * 0x100u * (unsigned)src_h) / (unsigned)dst_h;                        /* 8.8 source row, :1022-1024 */
        unsigned syf = sy & 0xFF, sy2, csy, csyf, csy2;
        int p;
        sy >>= 8;
        csy = sy; csyf = syf;
        if (src_is_420) { if (!(csy & 1)) csyf = 0; csy >>= 1; }                               /* :1028-1031 */
        if (src_interlaced) {                                                                   /* :1033-1081 */
            unsigned which = src_top_field_first ? 0u : 1u;
            if (second_field) which ^= 1u;
            if (which == 0) {
                if (sy & 1u) { sy++; syf = 0; }             /* odd -> the even line above it... one line down, no interpolation */
                if (csy & 1u) { csy++; csyf = 0; }
            } else {
                if (!(sy & 1u)) { sy++; syf = 0; }
                if (!(csy & 1u)) { csy++; csyf = 0; }
            }
            if (sy >= (unsigned)(src_h - 2)) { sy = (unsigned)(src_h - 2); syf = 0; }
            sy2 = sy + 2;
            if (csy >= chroma_h - 2) { csy = chroma_h - 2; csyf = 0; }
            csy2 = csy + 1;                                                                     /* sic: +1, :1080 */
        } else {                                                                                /* :1082-1094 */
            if (sy >= (unsigned)(src_h - 1)) { sy = (unsigned)(src_h - 1); syf = 0; }
            sy2 = sy + 1;
            if (csy >= chroma_h - 1) { csy = chroma_h - 1; csyf = 0; }
            csy2 = csy + 1;
        }
        for (p = 0; p < 3; p++) {
            /* 4:2:0 sources use the chroma row/fraction for planes 1,2 (:1096-1116); 4:2:2 sources use
               the luma row and fraction for all three planes (:1117-1137) */
            const unsigned a = (src_is_420 && p > 0) ? csy : sy, b = (src_is_420 && p > 0) ? csy2 : sy2;
            const unsigned fr = (src_is_420 && p > 0) ? csyf : syf;
            const uint8_t *s1 = src[p] + (size_t)src_linesize[p] * a;
            const uint8_t *s2 = src[p] + (size_t)src_linesize[p] * b;
            uint8_t *d = dst[p] + (size_t)dst_linesize[p] * y;
            int x;
            if (fr == 0) memcpy(d, s1, (size_t)row_bytes[p]);
            else
                for (x = 0; x < row_bytes[p]; x++)
                    d[x] = (uint8_t)(s1[x] + ((uint8_t)((((int)s2[x] - (int)s1[x]) * (int)fr) >> 8)));
        }
    }
}
