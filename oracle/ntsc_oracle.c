/*
 * ntsc_oracle.c -- CPU restatement of composite_layer() (ffmpeg_ntsc.cpp:1570-1921).
 *
 * TEST INFRASTRUCTURE ONLY (see ntsc_oracle.h).  Parity status: PINNED against the
 * reference's own code (oracle/_ref/libref.so) by tests/test_oracle_vs_ref.py.
 *
 * This is NOT a transcription of the reference: the reference is stage-major (each stage
 * sweeps the whole field over three heap planes and pulls libc rand() as it goes); this
 * restatement is LINE-major (each scanline goes through every stage before the next line
 * starts, with only the true cross-line state carried: three noise accumulators, the
 * vertical-blend delay line and the head-switch shift schedule) and draws from an explicit
 * model of glibc's generator, pre-drawn per call and indexed by the reference's draw
 * order.  That is the same decomposition the CUDA kernels use, so a disagreement between
 * the two structures and the reference shows up here, on the CPU, first.
 * Arithmetic (double op order, truncations, integer divisions) follows the cited lines
 * exactly; build with -ffp-contract=off.
 */
#include "ntsc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * glibc random_r() TYPE_3: r[i] = r[i-31] + r[i-3] (mod 2^32), output r[i] >> 1, seeded by
 * the Lehmer sequence 16807*x mod (2^31-1) and 310 discarded outputs.  (glibc 2.39
 * stdlib/random_r.c; verified against this container's rand() in tests/test_rng.py.)
 * ------------------------------------------------------------------------------------------ */
void oracle_rng_seed(oracle_rng *g, unsigned seed) {
    int32_t word;
    int i;
    if (seed == 0) seed = 1;
    g->r[0] = seed;
    word = (int32_t)seed;
    for (i = 1; i < 31; i++) {
        long hi = word / 127773, lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        g->r[i] = (uint32_t)word;
    }
    g->f = 3;
    g->b = 0;
    g->pos = 0;
    for (i = 0; i < 310; i++) (void)oracle_rng_next(g);
    g->pos = 0;
}

uint32_t oracle_rng_next(oracle_rng *g) {
    uint32_t v;
    g->r[g->f] += g->r[g->b];
    v = g->r[g->f] >> 1;
    if (++g->f == 31) g->f = 0;
    if (++g->b == 31) g->b = 0;
    g->pos++;
    return v;
}

/* ------------------------------------------------------------------------------------------ */

/* FNV-1a-64 over a byte buffer: the known-answer hash of SURVEY.md App. D. */
unsigned long long oracle_fnv1a64(const void *buf, unsigned long long n) {
    const unsigned char *b = (const unsigned char *)buf;
    unsigned long long h = 1469598103934665603ULL, i;
    for (i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ULL; }
    return h;
}

/* "field deinterlace" of the reference's main loop (ffmpeg_ntsc.cpp:2232-2257): line-double the field
 * that composite_layer() just wrote, in place. */
void oracle_bob(uint8_t *pic, int stride, int w, int h, unsigned field) {
    int y;
    if (field) {
        for (y = (int)field; y < h; y += 2)                           /* :2237-2245 */
            memcpy(pic + (size_t)stride * (size_t)(y - 1), pic + (size_t)stride * (size_t)y, (size_t)w * 4);
    } else {
        for (y = 1; y + 1 < h; y += 2)                                /* :2247-2255 */
            memcpy(pic + (size_t)stride * (size_t)y, pic + (size_t)stride * (size_t)(y + 1), (size_t)w * 4);
    }
}

static oracle_tap_fn g_tap = 0;
static void *g_tap_user = 0;
void oracle_set_tap(oracle_tap_fn fn, void *user) { g_tap = fn; g_tap_user = user; }

/* One-pole lowpass, ffmpeg_ntsc.cpp:74-106.  Three roundings per step, no FMA. */
typedef struct { double alpha, prev; } pole;

static double pole_alpha(double hz) {                      /* setFilter, :78-86 */
    const double rate = (315000000.00 * 4) / 88;            /* 4 x fsc sample clock */
    double timeInterval = 1.0 / rate;
    double tau = 1 / (hz * 2 * M_PI);
    return timeInterval / (tau + timeInterval);
}
static double pole_lp(pole *f, double s) {                  /* lowpass, :90-94 */
    const double stage1 = s * f->alpha;
    const double stage2 = f->prev - (f->prev * f->alpha);
    return (f->prev = (stage1 + stage2));
}
static double pole_hp(pole *f, double s) { return s - pole_lp(f, s); }   /* highpass, :95-99 */

static void cascade3_init(pole c[3], double alpha, double reset) {
    int k;
    for (k = 0; k < 3; k++) { c[k].alpha = alpha; c[k].prev = reset; }
}
static double cascade3(pole c[3], double s) {
    s = pole_lp(&c[0], s);
    s = pole_lp(&c[1], s);
    return pole_lp(&c[2], s);
}

/* Three cascaded poles over one row with the result written `delay` samples early; the last
   `delay` samples keep their input values.  composite_lowpass :1429-1458, composite_lowpass_tv
   :1399-1427, VHS chroma lowpass :1814-1836. */
static void row_lowpass_shift(int *P, int w, double alpha, int delay) {
    pole c[3];
    int x;
    cascade3_init(c, alpha, 0);
    for (x = 0; x < w; x++) {
        double s = cascade3(c, (double)P[x]);
        if (x >= delay) P[x - delay] = (int)s;
    }
}

/* Subcarrier phase index of a scanline, ffmpeg_ntsc.cpp:1473-1480 (== :1530-1537). */
static unsigned line_phase(const cvs_params *p, unsigned long long fieldno, unsigned y) {
    const int off = p->video_scanline_phase_shift_offset;
    switch (p->video_scanline_phase_shift) {
    case 90:  return (unsigned)((fieldno + off + (y >> 1)) & 3);
    case 180: return (unsigned)((((fieldno + y) & 2) + off) & 3);
    case 270: return (unsigned)((fieldno + off - (y >> 1)) & 3);
    default:  return (unsigned)(off & 3);
    }
}

/* QAM-modulate I/Q onto the 4-sample subcarrier and add to luma; chroma_into_luma :1460-1495. */
static void row_modulate(int *Y, int *I, int *Q, int w, unsigned xi, int amp) {
    static const int Umult[4] = { 1, 0, -1, 0 };
    static const int Vmult[4] = { 0, 1, 0, -1 };
    int x;
    for (x = 0; x < w; x++) {
        unsigned ph = (xi + (unsigned)x) & 3;
        int chroma = I[x] * amp * Umult[ph] + Q[x] * amp * Vmult[ph];
        Y[x] += chroma / 50;
        I[x] = 0;
        Q[x] = 0;
    }
}

/* Box-filter Y/C separation + QAM demodulation; chroma_from_luma :1497-1567.
   `ch` is scratch of w ints. */
static void row_demodulate(int *Y, int *I, int *Q, int *ch, int w, unsigned xi, int amp) {
    int x;
    int ym1 = 0;                                   /* original Y[x-1] (box reads originals) */
    for (x = 0; x < w; x++) {
        int y0 = Y[x];
        int y1 = (x + 1 < w) ? Y[x + 1] : 0;
        int y2 = (x + 2 < w) ? Y[x + 2] : 0;
        int box = (ym1 + y0 + y1 + y2) / 4;        /* 4-tap box centred between x and x+1, :1506-1523 */
        ch[x] = y2 - box;                          /* note the +2 sample offset, :1514-1524 */
        ym1 = y0;
        Y[x] = box;
    }
    for (x = (int)((4 - xi) & 3); x + 3 < w; x += 4) {    /* :1539-1542 */
        ch[x + 2] = -ch[x + 2];
        ch[x + 3] = -ch[x + 3];
    }
    for (x = 0; x < w; x++) ch[x] = (ch[x] * 50) / amp;    /* :1544-1546 */
    for (x = 0; x + (int)xi + 1 < w; x += 2) {             /* :1549-1552 */
        I[x] = -ch[x + xi];
        Q[x] = -ch[x + xi + 1];
    }
    for (; x < w; x += 2) { I[x] = 0; Q[x] = 0; }          /* :1553-1556 */
    for (x = 0; x + 2 < w; x += 2) {                       /* :1557-1560 */
        I[x + 1] = (I[x] + I[x + 2]) >> 1;
        Q[x + 1] = (Q[x] + Q[x + 2]) >> 1;
    }
    for (; x < w; x++) { I[x] = 0; Q[x] = 0; }             /* :1561-1564 */
}

unsigned long long oracle_draws_per_field(const cvs_params *p, int w, int h, unsigned field) {
    unsigned long long nl = 0, n = 0;
    if (h > (int)field) nl = (unsigned long long)((h - (int)field + 1) / 2);
    if (p->video_noise != 0) n += nl * (unsigned long long)w;                                   /* :1632-1644 */
    if (p->vhs_head_switching && p->vhs_head_switching_phase_noise != 0) n += 4;                /* :1654-1655 */
    if (p->video_chroma_noise != 0) n += 2 * nl * (unsigned long long)w;                        /* :1719-1735 */
    if (p->video_chroma_phase_noise != 0) n += nl;                                              /* :1736-1764 */
    if (p->video_chroma_loss != 0) n += nl;                                                     /* :1891-1901 */
    return n;
}

int oracle_composite_layer(const cvs_params *p, oracle_rng *g,
                           uint8_t *dst, int dst_stride,
                           const uint8_t *src, int src_stride,
                           int w, int h, int src_interlaced, int src_tff,
                           unsigned field, unsigned long long fieldno) {
    unsigned long long ndraw, k;
    uint32_t *draw;
    size_t offL, offH, offC, offP, offD;
    int nl, r, x;
    int *Y, *I, *Q, *ch, *tmp, *prevU, *prevV, *hs_shift;
    int opposite;
    int noiseY = 0, noiseU = 0, noiseV = 0, noiseP = 0;
    double a_inI, a_inQ, a_tv, a_pre = 0, a_luma = 0, a_chroma = 0, a_sharp = 0;
    int chroma_delay = 0;
    const unsigned twidth = (unsigned)w + (unsigned)w / 10;

    /* guards, :1578-1583 */
    if (!dst || !src) return -1;
    if (w <= 0 || h <= 0) return -1;
    if (dst_stride < w * 4 || src_stride < w * 4) return -1;

    opposite = src_interlaced ? (src_tff ? 1 : 0) : 0;               /* :1585-1588 */
    nl = (h > (int)field) ? (h - (int)field + 1) / 2 : 0;

    /* Pre-draw this call's rand() values in the reference's order:
       [luma noise nl*w][head-switch 4][chroma noise 2*nl*w][phase noise nl][dropout nl]. */
    ndraw = oracle_draws_per_field(p, w, h, field);
    draw = (uint32_t *)malloc((size_t)(ndraw + 1) * sizeof(uint32_t));
    if (!draw) return -1;
    for (k = 0; k < ndraw; k++) draw[k] = oracle_rng_next(g);
    offL = 0;
    offH = offL + ((p->video_noise != 0) ? (size_t)nl * w : 0);
    offC = offH + ((p->vhs_head_switching && p->vhs_head_switching_phase_noise != 0) ? 4 : 0);
    offP = offC + ((p->video_chroma_noise != 0) ? (size_t)2 * nl * w : 0);
    offD = offP + ((p->video_chroma_phase_noise != 0) ? (size_t)nl : 0);

    Y = (int *)malloc(sizeof(int) * (size_t)w);
    I = (int *)malloc(sizeof(int) * (size_t)w);
    Q = (int *)malloc(sizeof(int) * (size_t)w);
    ch = (int *)malloc(sizeof(int) * (size_t)w);
    tmp = (int *)malloc(sizeof(int) * (size_t)(twidth + 1));
    prevU = (int *)calloc((size_t)w, sizeof(int));                   /* zero-initialised delay line, :1847-1848 */
    prevV = (int *)calloc((size_t)w, sizeof(int));
    hs_shift = (int *)calloc((size_t)(nl + 1), sizeof(int));

    /* filter constants, SURVEY App. A.4 */
    a_inI = pole_alpha(1300000);                                     /* :1442 */
    a_inQ = pole_alpha(600000);
    a_tv = pole_alpha(2600000);                                      /* :1411 */
    if (p->composite_preemphasis != 0 && p->composite_preemphasis_cut > 0)
        a_pre = pole_alpha(p->composite_preemphasis_cut);            /* :1621 */
    if (p->emulating_vhs) {                                          /* :1773-1791 */
        double luma_cut, chroma_cut;
        switch (p->output_vhs_tape_speed) {
        case CVS_VHS_SP: luma_cut = 2400000; chroma_cut = 320000; chroma_delay = 9; break;
        case CVS_VHS_LP: luma_cut = 1900000; chroma_cut = 300000; chroma_delay = 12; break;
        case CVS_VHS_EP: luma_cut = 1400000; chroma_cut = 280000; chroma_delay = 14; break;
        default: abort();
        }
        a_luma = pole_alpha(luma_cut);
        a_chroma = pole_alpha(chroma_cut);
        a_sharp = pole_alpha(luma_cut * 4);                          /* :1874 */
    }

    /* VHS head-switch schedule, :1646-1713: which rows get rotated, and by how much. */
    if (p->vhs_head_switching) {
        double noise = 0, t;
        unsigned pp, hx;
        int y, shif = 0, ishif, shy = 0;
        if (p->vhs_head_switching_phase_noise != 0) {
            unsigned v = draw[offH] * draw[offH + 1] * draw[offH + 2] * draw[offH + 3];   /* wraps mod 2^32 */
            v %= 2000000000U;
            noise = ((double)v / 1000000000U) - 1.0;
            noise *= p->vhs_head_switching_phase_noise;
        }
        t = p->output_ntsc ? twidth * 262.5 : twidth * 312.5;
        pp = (unsigned)(fmod(p->vhs_head_switching_point + noise, 1.0) * t);
        y = (int)(((pp / twidth) * 2) + field);
        pp = (unsigned)(fmod(p->vhs_head_switching_phase + noise, 1.0) * t);
        hx = pp % twidth;
        y -= p->output_ntsc ? (262 - 240) * 2 : (312 - 288) * 2;
        ishif = (hx >= twidth / 2) ? (int)hx - (int)twidth : (int)hx;
        while (y < h) {
            /* the first affected row has shif == 0 (a no-op), so the start column `tx`
               only ever matters as 0, :1683-1711 */
            if (y >= 0 && shif != 0) hs_shift[(y - (int)field) / 2] = shif;
            shif = (shy == 0) ? ishif : (shif * 7) / 8;
            y += 2;
            shy++;
        }
    }

    for (r = 0; r < nl; r++) {
        const int y = (int)field + 2 * r;
        const unsigned xi = line_phase(p, fieldno, (unsigned)y);
        int sy = y + opposite;
        const uint8_t *srow;
        uint32_t *drow;
        if (sy > h - 1) sy = h - 1;                                   /* :1599 */
        srow = src + (size_t)src_stride * (size_t)sy;

        /* RGB -> YIQ, :1375-1383, :1598-1606 */
        for (x = 0; x < w; x++) {
            const int b = srow[4 * x + 0], gch = srow[4 * x + 1], rch = srow[4 * x + 2];
            double dY = (0.30 * rch) + (0.59 * gch) + (0.11 * b);
            Y[x] = (int)(256 * dY);
            I[x] = (int)(256 * ((-0.27 * (b - dY)) + (0.74 * (rch - dY))));
            Q[x] = (int)(256 * ((0.41 * (b - dY)) + (0.48 * (rch - dY))));
        }

        if (p->composite_in_chroma_lowpass) {                        /* :1608-1609 */
            row_lowpass_shift(I, w, a_inI, 2);
            row_lowpass_shift(Q, w, a_inQ, 4);
        }

        row_modulate(Y, I, Q, w, xi, p->subcarrier_amplitude);      /* :1611 */

        if (p->composite_preemphasis != 0 && p->composite_preemphasis_cut > 0) {   /* :1613-1629 */
            pole pre;
            pre.alpha = a_pre;
            pre.prev = 16;
            for (x = 0; x < w; x++) {
                double s = Y[x];
                s += pole_hp(&pre, s) * p->composite_preemphasis;
                Y[x] = (int)s;
            }
        }

        if (p->video_noise != 0) {                                   /* :1631-1644; state carries across rows */
            const unsigned mod = (unsigned)(p->video_noise * 2 + 1);
            const uint32_t *d = draw + offL + (size_t)r * w;
            for (x = 0; x < w; x++) {
                Y[x] += noiseY;
                noiseY += (int)(d[x] % mod) - p->video_noise;
                noiseY /= 2;
            }
        }
        if (g_tap) g_tap(g_tap_user, 1, r, w, Y, I, Q);

        if (hs_shift[r] != 0) {                                      /* :1687-1700 */
            unsigned x2 = (twidth + (unsigned)hs_shift[r]) % twidth;
            memset(tmp, 0, sizeof(int) * twidth);
            memcpy(tmp, Y, sizeof(int) * (size_t)w);
            for (x = 0; x < w; x++) {
                Y[x] = tmp[x2];
                if (++x2 == twidth) x2 = 0;
            }
        }

        if (!p->nocolor_subcarrier)                                  /* :1715-1716 */
            row_demodulate(Y, I, Q, ch, w, xi, p->subcarrier_amplitude_back);

        if (p->video_chroma_noise != 0) {                            /* :1718-1735 */
            const unsigned mod = (unsigned)(p->video_chroma_noise * 2 + 1);
            const uint32_t *d = draw + offC + (size_t)2 * r * w;
            for (x = 0; x < w; x++) {
                I[x] += noiseU;
                Q[x] += noiseV;
                noiseU += (int)(d[2 * x] % mod) - p->video_chroma_noise;
                noiseU /= 2;
                noiseV += (int)(d[2 * x + 1] % mod) - p->video_chroma_noise;
                noiseV /= 2;
            }
        }
        if (p->video_chroma_phase_noise != 0) {                      /* :1736-1764 */
            const unsigned mod = (unsigned)(p->video_chroma_phase_noise * 2 + 1);
            double pi, sinpi, cospi;
            noiseP += (int)(draw[offP + r] % mod) - p->video_chroma_phase_noise;
            noiseP /= 2;
            pi = ((double)noiseP * M_PI) / 100;
            sinpi = sin(pi);
            cospi = cos(pi);
            for (x = 0; x < w; x++) {
                double u = I[x], v = Q[x];
                double u_ = (u * cospi) - (v * sinpi);
                double v_ = (u * sinpi) + (v * cospi);
                I[x] = (int)u_;
                Q[x] = (int)v_;
            }
        }
        if (g_tap) g_tap(g_tap_user, 2, r, w, Y, I, Q);

        if (p->emulating_vhs) {                                      /* :1769-1889 */
            pole lp[3], pre;
            cascade3_init(lp, a_luma, 16);                           /* luma lowpass + 1.6x HF, :1793-1812 */
            pre.alpha = a_luma;
            pre.prev = 16;
            for (x = 0; x < w; x++) {
                double s = cascade3(lp, (double)Y[x]);
                s += pole_hp(&pre, s) * 1.6;
                Y[x] = (int)s;
            }
            row_lowpass_shift(I, w, a_chroma, chroma_delay);         /* :1814-1836 (U and V are independent) */
            row_lowpass_shift(Q, w, a_chroma, chroma_delay);

            if (p->vhs_chroma_vert_blend && p->output_ntsc && r >= 1) {   /* :1843-1863; loop starts at field+2 */
                for (x = 0; x < w; x++) {
                    int cU = I[x], cV = Q[x];
                    I[x] = (prevU[x] + cU + 1) >> 1;
                    Q[x] = (prevV[x] + cV + 1) >> 1;
                    prevU[x] = cU;
                    prevV[x] = cV;
                }
            }

            cascade3_init(lp, a_sharp, 0);                           /* sharpen, :1865-1883 */
            for (x = 0; x < w; x++) {
                double s = Y[x], ts = cascade3(lp, s);
                Y[x] = (int)(s + ((s - ts) * p->vhs_out_sharpen * 2));
            }

            if (!p->vhs_svideo_out) {                                /* :1885-1888 (amplitude, not _back) */
                row_modulate(Y, I, Q, w, xi, p->subcarrier_amplitude);
                row_demodulate(Y, I, Q, ch, w, xi, p->subcarrier_amplitude);
            }
        }
        if (g_tap) g_tap(g_tap_user, 3, r, w, Y, I, Q);

        if (p->video_chroma_loss != 0) {                             /* :1891-1901 */
            if ((draw[offD + r] % 100000U) < (unsigned)p->video_chroma_loss) {
                memset(I, 0, sizeof(int) * (size_t)w);
                memset(Q, 0, sizeof(int) * (size_t)w);
            }
        }

        if (p->composite_out_chroma_lowpass) {                       /* :1903-1908 */
            if (p->composite_out_chroma_lowpass_lite) {
                row_lowpass_shift(I, w, a_tv, 1);
                row_lowpass_shift(Q, w, a_tv, 1);
            } else {
                row_lowpass_shift(I, w, a_inI, 2);
                row_lowpass_shift(Q, w, a_inQ, 4);
            }
        }
        if (g_tap) g_tap(g_tap_user, 4, r, w, Y, I, Q);

        /* YIQ -> RGB, :1385-1396, :1910-1916; alpha byte = 0 */
        drow = (uint32_t *)(dst + (size_t)dst_stride * (size_t)y);
        for (x = 0; x < w; x++) {
            int rr = (int)(((1.000 * Y[x]) + (0.956 * I[x]) + (0.621 * Q[x])) / 256);
            int gg = (int)(((1.000 * Y[x]) + (-0.272 * I[x]) + (-0.647 * Q[x])) / 256);
            int bb = (int)(((1.000 * Y[x]) + (-1.106 * I[x]) + (1.703 * Q[x])) / 256);
            if (rr < 0) rr = 0; else if (rr > 255) rr = 255;
            if (gg < 0) gg = 0; else if (gg > 255) gg = 255;
            if (bb < 0) bb = 0; else if (bb > 255) bb = 255;
            drow[x] = ((uint32_t)rr << 16) + ((uint32_t)gg << 8) + (uint32_t)bb;
        }
    }

    free(hs_shift); free(prevV); free(prevU); free(tmp); free(ch);
    free(Q); free(I); free(Y); free(draw);
    return 0;
}
