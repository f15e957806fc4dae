/* convert_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the two picture conversions that sit either side of the hot path in the reference's field
 * loop and that the product runs on the device:
 *   - InputFile::frame_copy_scale()  (ffmpeg_ntsc.cpp:544-613):  decoder picture -> BGRA at the output size,
 *     sws_scale() with SWS_BILINEAR (:574-585, :603-610);
 *   - the encoder-side sws_scale()   (ffmpeg_ntsc.cpp:2266-2274, context :2118-2131, SMPTE170M / MPEG range
 *     :2100-2101):  finished BGRA picture -> planar YUV 4:2:0 / 4:2:2.
 *
 * Both are calls into libswscale, a third-party dependency of the reference (it needs FFmpeg 3.x; no FFmpeg source or
 * headers exist in this environment).  PARITY PINNED for the library's default route: libswscale 9.1.100 (FFmpeg 8.0) is
 * present in this image as a binary (inside opencv-python-headless); tests/swscale_ref.py binds it with ctypes, makes the
 * reference's own calls, and tests/test_swscale_pin.py compares the functions below with it byte for byte (library C
 * code; tests/golden/swscale_*.npz carry its outputs elsewhere).  What each function restates is written above it, with
 * the names of the library routines that do it; the text was written from the library's documented structure and
 * validated against the binary -- none of it is the library's source.
 *
 * PARITY UNPINNED for one corner: a BGRA source of ODD width reduced to half its width or less.  The library's rule for it
 * is known and restated (oracle_sws_bgra_to_bgra) but was found after the last GPU run of the round, so the product still
 * routes it to the repository's own resampler, and oracle_scale_to_bgra follows the product (first part of this file:
 * triangle-kernel resampling with 14-bit weights and a 15-bit intermediate, channel by channel), restated here from its
 * specification in csrc/scale_convert.cuh and NOT from the kernel; a BGRA source of the output size is a copy in both.
 */
#include <stdint.h>
#include <stdlib.h>

static int clamp8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
static int clampi(long long v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : (int)v); }
static long long floordiv(long long a, long long b) { long long q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) q--; return q; }

/* One output sample of an axis: source position P / D, triangle kernel of half-width H = max(D, 2 n_src) (in 1/D
 * units), 14-bit weights that sum to 16384 (remainder to the largest tap, first of equals).  Calls back with
 * (source index, weight) for every tap, indices NOT yet clamped. */
typedef struct { long long j; int w; } tap;
static int axis_taps(int i, int n_dst, int n_src_luma, int sub, int off, tap *out, int max_out) {
    const long long D = 2LL * n_dst * sub;
    const long long P = (2LL * i + 1) * n_src_luma - n_dst - (long long)off * n_dst;
    const long long H = D > 2LL * n_src_luma ? D : 2LL * n_src_luma;
    /* integers j with |j D - P| < H */
    const long long jlo = floordiv(P - H, D) + 1, jhi = -floordiv(-(P + H), D) - 1;   /* ceil((P+H)/D) - 1 */
    long long sum = 0, best_t = -1;
    int n = 0, best = 0;
    for (long long j = jlo; j <= jhi && n < max_out; j++) {
        long long d = j * D - P;
        if (d < 0) d = -d;
        const long long t = H - d;
        if (t <= 0) continue;
        out[n].j = j;
        out[n].w = 0;
        sum += t;
        if (t > best_t) { best_t = t; best = n; }
        n++;
    }
    long long acc = 0;
    for (int k = 0; k < n; k++) {
        long long d = out[k].j * D - P;
        if (d < 0) d = -d;
        out[k].w = (int)(((H - d) * 16384) / sum);
        acc += out[k].w;
    }
    out[best].w += (int)(16384 - acc);
    return n;
}

/* one sample of one plane (pw x ph samples, `step` bytes apart in a row) at destination (x, y) */
static int plane_sample(const uint8_t *plane, int linesize, int pw, int ph, int step,
                        int x, int dw, int sw_luma, int subx, int y, int dh, int sh_luma, int suby, int offy) {
    tap tx[64], ty[64];
    const int nx = axis_taps(x, dw, sw_luma, subx, 0, tx, 64);
    const int ny = axis_taps(y, dh, sh_luma, suby, offy, ty, 64);
    long long acc = 1 << 20;
    for (int k = 0; k < ny; k++) {
        const uint8_t *row = plane + (size_t)clampi(ty[k].j, 0, ph - 1) * (size_t)linesize;
        long long h = 64;
        for (int j = 0; j < nx; j++) h += (long long)tx[j].w * row[(size_t)clampi(tx[j].j, 0, pw - 1) * (size_t)step];
        acc += (long long)ty[k].w * (h >> 7);
    }
    return clamp8((int)(acc >> 21));
}

typedef struct { int size; int *pos; int *coef; } swsfilter;      /* coef[i * size + j] applies to sample pos[i] + j */
static int sws_bilinear_filter(swsfilter *f, int srcn, int dstn, int one, int srcpos, int dstpos);
static void swsfilter_free(swsfilter *f);
static int sws_q(double c);

/* ---- planar YUV -> BGRA at another size: libswscale's C path, restated ------------------------------------------------
 *
 * PINNED against libswscale 9.1.100 (tests/test_swscale_pin.py: sws_getContext(sw, sh, YUV420P | YUV422P | NV12, dw, dh,
 * BGRA, SWS_BILINEAR, NULL, NULL, NULL) + sws_scale(), the reference's call at ffmpeg_ntsc.cpp:574-585, 603-610, library
 * C code).  What the library does for EVEN destination widths (its function names, for orientation; odd ones further down):
 *   horizontal  swscale.c hScale8To15_c: every source row of Y to dw samples, of U and V to ceil(dw/2) samples -- the
 *               library keeps ONE chroma sample per pair of output pixels -- with the bilinear banks of utils.c
 *               initFilter (14-bit weights), s15 = min(sum >> 7, 32767);
 *   vertical    banks with 12-bit weights, luma sh -> dh, chroma rows -> dh; vscale.c packed_vscale picks the output
 *               routine by the banks' tap counts:
 *                 1 luma tap, 1 chroma tap   (s + 64) >> 7 for Y, U, V                              (output.c yuv2rgb_1_c_template)
 *                 1 luma tap, 2 chroma taps  Y as above; C = (c0 (4096 - a) + c1 a + (128 << 11)) >> 19     (same, a = 2nd weight)
 *                 2 and 2                    (s0 (4096 - a) + s1 a) >> 19, no rounding term               (yuv2rgb_2_c_template)
 *                 anything else              ((1 << 18) + sum s_j w_j) >> 19                              (yuv2rgb_X_c_template)
 *   colour      yuv2rgb.c ff_yuv2rgb_c_init_tables (ITU-R 601, MPEG range in): with cy = 65536 * 255 / 219 and
 *               c' = (c * 65536 + 32768) / cy for c = 104597 (V->R), 132201 (U->B), -25675 (U->G), -53279 (V->G),
 *               channel = clip8((k cy - (384 << 16) - (16 << 16) + 326 cy + 32768) >> 16) at
 *               k = Y + (c' C >> 16) - (c' >> 9) (both chroma terms for G), C clipped to 0..255; alpha = 255.
 *   exception   YUV420P at the SAME size with an even height takes the library's direct converter (yuv2rgb.c
 *               yuv2rgb_c_32): no filtering at all, the chroma sample of a 2x2 block serves its four pixels.
 * BGRA sources: sws_bgra_to_bgra further down.
 */
static int sws_rgb_k(long long k) {
    const long long cy = (65536LL * 255) / 219;
    long long v = (k * cy - (384LL << 16) - (16LL << 16) + 326 * cy + 32768) >> 16;
    return clamp8((int)v);
}
static uint32_t sws_pixel(int Y, int U, int V) {
    const long long cy = (65536LL * 255) / 219;
    const long long crv = (104597LL * 65536 + 32768) / cy, cbu = (132201LL * 65536 + 32768) / cy;
    const long long cgu = -((25675LL * 65536 - 32768) / cy), cgv = -((53279LL * 65536 - 32768) / cy);   /* C division of a negative numerator */
    U = clamp8(U); V = clamp8(V);
    const int r = sws_rgb_k(Y + ((crv * V) >> 16) - (crv >> 9));
    const int b = sws_rgb_k(Y + ((cbu * U) >> 16) - (cbu >> 9));
    const int g = sws_rgb_k(Y + ((cgu * U) >> 16) - (cgu >> 9) + ((cgv * V) >> 16) - (cgv >> 9));
    return 0xFF000000u | ((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b;
}
/* all rows of one plane scaled horizontally to 15 bits; `step` bytes between the samples of a source row */
static int *sws_hscale(const uint8_t *plane, int linesize, int step, int rows, const swsfilter *f, int dstn) {
    int *out = (int *)malloc(sizeof(int) * (size_t)rows * (size_t)dstn);
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < dstn; x++) {
            long long v = 0;
            for (int j = 0; j < f->size; j++)
                v += (long long)plane[(size_t)y * (size_t)linesize + (size_t)(f->pos[x] + j) * (size_t)step] * f->coef[(size_t)x * f->size + j];
            v >>= 7;
            out[(size_t)y * dstn + x] = (int)(v < 32767 ? v : 32767);
        }
    return out;
}
/* Odd destination widths: the library turns on full horizontal chroma interpolation (utils.c, "Forcing full internal H
 * chroma due to odd output size"): chroma is scaled to dw samples and the writers are output.c yuv2rgb_full_{1,2,X}_c_template
 * + yuv2rgb_write_full -- no tables but 32-bit integer arithmetic with the coefficients of yuv2rgb.c rounded to 16 bits
 * (y 9539, offset 8192, V->R 13075, V->G -6660, U->G -3209, U->B 16525), on samples at 2^9 per code:
 *     Y' = (Y - 8192) 9539 + 2^21;  R = Y' + V 13075;  G = Y' + V (-6660) + U (-3209);  B = Y' + U 16525   (mod 2^32)
 * clipped to 30 bits when any of them leaves that range, then >> 22.  The sums wrap at 32 bits before they are clipped, so
 * a far-out-of-gamut sample (Y = U = 255) comes out 0 where 255 would be expected: the library's behaviour, kept. */
static int clip_uintp2_30(int32_t a) { return (a & ~((1 << 30) - 1)) ? ((~a) >> 31) & ((1 << 30) - 1) : a; }
static uint32_t sws_pixel_full(int Y, int U, int V) {
    const uint32_t y = (uint32_t)(Y - 8192) * 9539u + (1u << 21);
    int32_t R = (int32_t)(y + (uint32_t)V * 13075u);
    int32_t G = (int32_t)(y + (uint32_t)V * (uint32_t)-6660 + (uint32_t)U * (uint32_t)-3209);
    int32_t B = (int32_t)(y + (uint32_t)U * 16525u);
    if ((R | G | B) & 0xC0000000) { R = clip_uintp2_30(R); G = clip_uintp2_30(G); B = clip_uintp2_30(B); }
    return 0xFF000000u | ((uint32_t)(R >> 22) << 16) | ((uint32_t)(G >> 22) << 8) | (uint32_t)(B >> 22);
}

static int sws_yuv_to_bgra(uint8_t *dst, int dst_stride, int dw, int dh, const uint8_t *p0, const uint8_t *p1, const uint8_t *p2,
                           int l0, int l1, int l2, int sw, int sh, int format) {
    const int full = dw & 1;                                   /* full horizontal chroma: one chroma sample per pixel */
    const int cw = (sw + 1) / 2, ch = (format == 2) ? sh : (sh + 1) / 2, cdw = full ? dw : (dw + 1) / 2;
    const uint8_t *pu = p1, *pv = (format == 3) ? p1 + 1 : p2;
    const int lu = l1, lv = (format == 3) ? l1 : l2, cstep = (format == 3) ? 2 : 1;
    if (format == 1 && sw == dw && sh == dh && (dh & 1) == 0 && !full) {   /* the direct converter */
        for (int y = 0; y < dh; y++)
            for (int x = 0; x < dw; x++)
                ((uint32_t *)(dst + (size_t)y * (size_t)dst_stride))[x] =
                    sws_pixel(p0[(size_t)y * l0 + x], pu[(size_t)(y / 2) * lu + x / 2], pv[(size_t)(y / 2) * lv + x / 2]);
        return 0;
    }
    swsfilter hl, hc, vl, vc;
    sws_bilinear_filter(&hl, sw, dw, 1 << 14, 128, 128);
    sws_bilinear_filter(&hc, cw, cdw, 1 << 14, 128, 128);
    sws_bilinear_filter(&vl, sh, dh, 1 << 12, 128, 128);
    sws_bilinear_filter(&vc, ch, dh, 1 << 12, 128, 128);
    int *L = sws_hscale(p0, l0, 1, sh, &hl, dw), *CU = sws_hscale(pu, lu, cstep, ch, &hc, cdw), *CV = sws_hscale(pv, lv, cstep, ch, &hc, cdw);
    for (int y = 0; y < dh; y++) {
        const int *lw = vl.coef + (size_t)y * vl.size, *cwt = vc.coef + (size_t)y * vc.size;
        const int *l_ = L + (size_t)vl.pos[y] * dw, *u_ = CU + (size_t)vc.pos[y] * cdw, *v_ = CV + (size_t)vc.pos[y] * cdw;
        const int two_c = vc.size == 2 && cwt[0] + cwt[1] == 4096 && cwt[1] >= 0 && cwt[1] <= 4096;
        const int two_l = vl.size == 2 && lw[0] + lw[1] == 4096 && lw[1] >= 0 && lw[1] <= 4096;
        int mode;
        if (vl.size == 1 && vc.size == 1) mode = 0;
        else if (vl.size == 1 && two_c) mode = 1;
        else if (two_l && two_c) mode = 2;
        else mode = 3;
        uint32_t *row = (uint32_t *)(dst + (size_t)y * (size_t)dst_stride);
        for (int x = 0; x < dw && full; x++) {                 /* yuv2rgb_full_{1,2,X}: samples at 2^9 per code, chroma minus 128 */
            int Y, U, V;
            if (mode == 0 || mode == 1) {
                const int a = mode == 1 ? cwt[1] : 0;
                Y = l_[x] * 4;
                U = mode == 1 ? (u_[x] * (4096 - a) + u_[cdw + x] * a - (128 << 19)) >> 10 : (u_[x] - (128 << 7)) * 4;
                V = mode == 1 ? (v_[x] * (4096 - a) + v_[cdw + x] * a - (128 << 19)) >> 10 : (v_[x] - (128 << 7)) * 4;
            } else if (mode == 2) {
                const int a = lw[1], c = cwt[1];
                Y = (l_[x] * (4096 - a) + l_[dw + x] * a) >> 10;
                U = (u_[x] * (4096 - c) + u_[cdw + x] * c - (128 << 19)) >> 10;
                V = (v_[x] * (4096 - c) + v_[cdw + x] * c - (128 << 19)) >> 10;
            } else {
                Y = 1 << 9;
                U = V = (1 << 9) - (128 << 19);
                for (int j = 0; j < vl.size; j++) Y += l_[(size_t)j * dw + x] * lw[j];
                for (int j = 0; j < vc.size; j++) { U += u_[(size_t)j * cdw + x] * cwt[j]; V += v_[(size_t)j * cdw + x] * cwt[j]; }
                Y >>= 10; U >>= 10; V >>= 10;
            }
            row[x] = sws_pixel_full(Y, U, V);
        }
        for (int x = 0; x < dw && !full; x++) {
            const int i = x >> 1;
            int Y, U, V;
            if (mode == 0) {
                Y = (l_[x] + 64) >> 7; U = (u_[i] + 64) >> 7; V = (v_[i] + 64) >> 7;
            } else if (mode == 1) {
                const int a = cwt[1];
                Y = (l_[x] + 64) >> 7;
                U = (u_[i] * (4096 - a) + u_[cdw + i] * a + (128 << 11)) >> 19;
                V = (v_[i] * (4096 - a) + v_[cdw + i] * a + (128 << 11)) >> 19;
            } else if (mode == 2) {
                const int a = lw[1], c = cwt[1];
                Y = (l_[x] * (4096 - a) + l_[dw + x] * a) >> 19;
                U = (u_[i] * (4096 - c) + u_[cdw + i] * c) >> 19;
                V = (v_[i] * (4096 - c) + v_[cdw + i] * c) >> 19;
            } else {
                Y = U = V = 1 << 18;
                for (int j = 0; j < vl.size; j++) Y += l_[(size_t)j * dw + x] * lw[j];
                for (int j = 0; j < vc.size; j++) { U += u_[(size_t)j * cdw + i] * cwt[j]; V += v_[(size_t)j * cdw + i] * cwt[j]; }
                Y >>= 19; U >>= 19; V >>= 19;
            }
            row[x] = sws_pixel(Y, U, V);
        }
    }
    free(L); free(CU); free(CV);
    swsfilter_free(&hl); swsfilter_free(&hc); swsfilter_free(&vl); swsfilter_free(&vc);
    return 0;
}

/* ---- BGRA -> BGRA at another size: libswscale's C path, restated ---------------------------------------------------------
 *
 * PINNED like the rest (tests/test_swscale_pin.py).  Packed RGB on both sides makes the library convert to YUV(A) at the
 * precision of its 15-bit intermediates and back with the full-chroma writers, alpha as a fourth plane:
 *   input       input.c rgb16_32ToY_c_template / rgb16_32ToUV_c_template (14-bit samples; chroma per PIXEL) and rgbaToA_c
 *               (a14 = a << 6 | a >> 2); when the width shrinks to half or less (dw <= sw / 2) chroma comes from the SUM of
 *               pixel pairs instead (rgb16_32ToUV_half_c_template);
 *   horizontal  swscale.c hScale16To15_c for Y, U, V, A alike: min(sum >> 13, 32767), banks sw -> dw (chroma: its own width);
 *   vertical    one bank sh -> dh for all four planes; output.c yuv2rgb_full_{1,2,X}_c_template (hasAlpha) by its tap count:
 *               1 tap   Y = s 4, C = (c - (128 << 7)) 4, A = (a + 64) >> 7
 *               2 taps  (s0 (4096 - w) + s1 w) >> 10 without a rounding term (chroma minus 128 << 19), A = (.. + 2^18) >> 19
 *               else    (2^9 + sum) >> 10, A = (2^18 + sum) >> 19
 *   colour      yuv2rgb_write_full as above (sws_pixel_full), alpha clipped to 0..255.
 * A source of the same size is copied (the library does not convert at all).  With an ODD source width the library keeps
 * chroma per pixel at every ratio; the product does not take that geometry through this route yet when dw <= sw / 2 (see
 * below), the repository's resampler serves it. */
static int *sws_hscale16(const long long *plane, int srcn, int rows, const swsfilter *f, int dstn) {
    int *out = (int *)malloc(sizeof(int) * (size_t)rows * (size_t)dstn);
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < dstn; x++) {
            long long v = 0;
            for (int j = 0; j < f->size; j++) v += plane[(size_t)y * srcn + f->pos[x] + j] * f->coef[(size_t)x * f->size + j];
            v >>= 13;
            out[(size_t)y * dstn + x] = (int)(v < 32767 ? v : 32767);
        }
    return out;
}
static int sws_bgra_to_bgra(uint8_t *dst, int dst_stride, int dw, int dh, const uint8_t *src, int ls, int sw, int sh) {
    /* (odd sw: the library keeps chroma per pixel whatever the ratio -- established after the GPU budget of the round was
     * spent, so the product still routes "odd sw, dw <= sw / 2" to its own resampler and oracle_scale_to_bgra follows it;
     * tests/test_swscale_pin.py::test_bgra_odd_width_corner shows the library's rule through oracle_sws_bgra_to_bgra) */
    const int ry = sws_q(0.299 * 219 / 255), gy = sws_q(0.587 * 219 / 255), by = sws_q(0.114 * 219 / 255);
    const int ru = -sws_q(0.169 * 224 / 255), gu = -sws_q(0.331 * 224 / 255), bu = sws_q(0.500 * 224 / 255);
    const int rv = sws_q(0.500 * 224 / 255), gv = -sws_q(0.419 * 224 / 255), bv = -sws_q(0.081 * 224 / 255);
    const int half = !(sw & 1) && dw <= (sw >> 1), cw = half ? sw / 2 : sw;
    long long *Y = (long long *)malloc(sizeof(long long) * (size_t)sw * sh), *A = (long long *)malloc(sizeof(long long) * (size_t)sw * sh);
    long long *U = (long long *)malloc(sizeof(long long) * (size_t)cw * sh), *V = (long long *)malloc(sizeof(long long) * (size_t)cw * sh);
    for (int y = 0; y < sh; y++) {
        const uint8_t *row = src + (size_t)y * (size_t)ls;
        for (int x = 0; x < sw; x++) {
            const uint8_t *p = row + 4 * (size_t)x;
            Y[(size_t)y * sw + x] = ((long long)ry * p[2] + (long long)gy * p[1] + (long long)by * p[0] + (16LL << 15) + 256) >> 9;
            A[(size_t)y * sw + x] = (p[3] << 6) | (p[3] >> 2);
            if (!half) {
                U[(size_t)y * cw + x] = ((long long)ru * p[2] + (long long)gu * p[1] + (long long)bu * p[0] + (256LL << 14) + 256) >> 9;
                V[(size_t)y * cw + x] = ((long long)rv * p[2] + (long long)gv * p[1] + (long long)bv * p[0] + (256LL << 14) + 256) >> 9;
            } else if ((x & 1) == 0) {
                const int r = p[2] + p[6], g = p[1] + p[5], b = p[0] + p[4];
                U[(size_t)y * cw + x / 2] = ((long long)ru * r + (long long)gu * g + (long long)bu * b + (256LL << 15) + 512) >> 10;
                V[(size_t)y * cw + x / 2] = ((long long)rv * r + (long long)gv * g + (long long)bv * b + (256LL << 15) + 512) >> 10;
            }
        }
    }
    swsfilter hl, hc, vf;
    sws_bilinear_filter(&hl, sw, dw, 1 << 14, 128, 128);
    sws_bilinear_filter(&hc, cw, dw, 1 << 14, 128, 128);
    sws_bilinear_filter(&vf, sh, dh, 1 << 12, 128, 128);
    int *L = sws_hscale16(Y, sw, sh, &hl, dw), *AL = sws_hscale16(A, sw, sh, &hl, dw);
    int *CU = sws_hscale16(U, cw, sh, &hc, dw), *CV = sws_hscale16(V, cw, sh, &hc, dw);
    for (int y = 0; y < dh; y++) {
        const int *w = vf.coef + (size_t)y * vf.size;
        const size_t o = (size_t)vf.pos[y] * dw;
        const int two = vf.size == 2 && w[0] + w[1] == 4096 && w[1] >= 0 && w[1] <= 4096;
        uint32_t *row = (uint32_t *)(dst + (size_t)y * (size_t)dst_stride);
        for (int x = 0; x < dw; x++) {
            int yv, uv, vv, av;
            if (vf.size == 1) {
                yv = L[o + x] * 4; uv = (CU[o + x] - (128 << 7)) * 4; vv = (CV[o + x] - (128 << 7)) * 4;
                av = (AL[o + x] + 64) >> 7;
            } else if (two) {
                const int a = w[1], a1 = 4096 - w[1];
                yv = (L[o + x] * a1 + L[o + dw + x] * a) >> 10;
                uv = (CU[o + x] * a1 + CU[o + dw + x] * a - (128 << 19)) >> 10;
                vv = (CV[o + x] * a1 + CV[o + dw + x] * a - (128 << 19)) >> 10;
                av = (AL[o + x] * a1 + AL[o + dw + x] * a + (1 << 18)) >> 19;
            } else {
                yv = 1 << 9; uv = vv = (1 << 9) - (128 << 19); av = 1 << 18;
                for (int j = 0; j < vf.size; j++) {
                    yv += L[o + (size_t)j * dw + x] * w[j]; uv += CU[o + (size_t)j * dw + x] * w[j];
                    vv += CV[o + (size_t)j * dw + x] * w[j]; av += AL[o + (size_t)j * dw + x] * w[j];
                }
                yv >>= 10; uv >>= 10; vv >>= 10; av >>= 19;
            }
            row[x] = (sws_pixel_full(yv, uv, vv) & 0x00FFFFFFu) | ((uint32_t)clamp8(av) << 24);
        }
    }
    free(Y); free(A); free(U); free(V); free(L); free(AL); free(CU); free(CV);
    swsfilter_free(&hl); swsfilter_free(&hc); swsfilter_free(&vf);
    return 0;
}

/* the library's route for a BGRA source of ANY geometry at another size (see the note in sws_bgra_to_bgra) */
int oracle_sws_bgra_to_bgra(uint8_t *dst, int dst_stride, int dw, int dh, const uint8_t *src, int ls, int sw, int sh) {
    if (!dst || !src || dw <= 0 || dh <= 0 || sw <= 0 || sh <= 0 || (sw == dw && sh == dh)) return -1;
    return sws_bgra_to_bgra(dst, dst_stride, dw, dh, src, ls, sw, sh);
}

/* format: 0 BGRA, 1 YUV420P, 2 YUV422P, 3 NV12 (the product's enum); dst: BGRA */
int oracle_scale_to_bgra(uint8_t *dst, int dst_stride, int dw, int dh,
                         const uint8_t *p0, const uint8_t *p1, const uint8_t *p2, int l0, int l1, int l2,
                         int sw, int sh, int format) {
    if (!dst || !p0 || dw <= 0 || dh <= 0 || sw <= 0 || sh <= 0 || format < 0 || format > 3) return -1;
    if (sw > 16 * dw || sh > 16 * dh) return -5;          /* more taps than the tables hold */
    if (format != 0) return sws_yuv_to_bgra(dst, dst_stride, dw, dh, p0, p1, p2, l0, l1, l2, sw, sh, format);
    if ((sw != dw || sh != dh) && !((sw & 1) && dw <= (sw >> 1))) return sws_bgra_to_bgra(dst, dst_stride, dw, dh, p0, l0, sw, sh);
    const int cw = (sw + 1) / 2, ch = (format == 2) ? sh : (sh + 1) / 2;
    const int suby = (format == 2) ? 1 : 2, offy = (format == 2) ? 0 : 1;
    for (int y = 0; y < dh; y++) {
        uint32_t *row = (uint32_t *)(dst + (size_t)y * (size_t)dst_stride);
        for (int x = 0; x < dw; x++) {
            if (format == 0) {
                uint32_t px = 0;
                for (int c = 0; c < 4; c++)
                    px |= (uint32_t)plane_sample(p0 + c, l0, sw, sh, 4, x, dw, sw, 1, y, dh, sh, 1, 0) << (8 * c);
                row[x] = px;
                continue;
            }
            const int Y = plane_sample(p0, l0, sw, sh, 1, x, dw, sw, 1, y, dh, sh, 1, 0);
            int U, V;
            if (format == 3) {
                U = plane_sample(p1, l1, cw, ch, 2, x, dw, sw, 2, y, dh, sh, suby, offy);
                V = plane_sample(p1 + 1, l1, cw, ch, 2, x, dw, sw, 2, y, dh, sh, suby, offy);
            } else {
                U = plane_sample(p1, l1, cw, ch, 1, x, dw, sw, 2, y, dh, sh, suby, offy);
                V = plane_sample(p2, l2, cw, ch, 1, x, dw, sw, 2, y, dh, sh, suby, offy);
            }
            const int c = 298 * (Y - 16), d = U - 128, e = V - 128;
            const int r = clamp8((c + 409 * e + 128) >> 8);
            const int g = clamp8((c - 100 * d - 208 * e + 128) >> 8);
            const int b = clamp8((c + 516 * d + 128) >> 8);
            row[x] = 0xFF000000u | ((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b;
        }
    }
    return 0;
}

/* ---- BGRA -> planar YUV 4:2:0 / 4:2:2: libswscale's C path, restated -------------------------------------------------
 *
 * PINNED against the real library: libswscale 9.1.100 (FFmpeg 8.0) is present in this image (bundled with
 * opencv-python-headless); tests/test_swscale_pin.py runs sws_getContext(w, h, BGRA, w, h, YUV420P | YUV422P,
 * SWS_BILINEAR, NULL, NULL, NULL) + sws_scale() -- the reference's own call, ffmpeg_ntsc.cpp:2118-2131, 2266-2274 -- with
 * the library's CPU extensions switched off (av_force_cpu_flags(0): its portable C code, which is also what
 * SWS_ACCURATE_RND | SWS_BITEXACT select) and compares this function with it byte for byte; tests/golden/
 * swscale_bgra_yuv.npz carries outputs of the library to machines that do not have it.  (The library's x86 vertical
 * scaler deviates from its own C code by +-1 on ~7 % of the 4:2:0 chroma samples; measured in the same test.)
 *
 * What the library does for this call (function names of libswscale, for orientation; nothing here is its code):
 *   luma    input.c rgb16_32ToY_c_template (14-bit sample from the 15-bit BT.601 matrix, bias 16.5 * 2^15 + 2^8),
 *           swscale.c hScale16To15_c with the unit filter (x 2, limited to 32767), output.c yuv2plane1_8_c ((s + 64) >> 7);
 *   chroma  even width: rgb16_32ToUV_half_c_template on the SUM of two neighbouring pixels; odd width (the library
 *           then keeps chroma at full width): rgb16_32ToUV_c_template per pixel and a horizontal bilinear filter with
 *           14-bit weights (utils.c initFilter);  then hScale16To15_c;  vertically 4:2:2 is yuv2plane1_8_c, 4:2:0 a
 *           bilinear filter over 2:1 (weights 1/8 3/8 3/8 1/8 in 12 bits, folded at the picture's edges) in
 *           output.c yuv2planeX_8_c ((64 << 12 + sum) >> 19).  Chroma positions are the library's defaults (centred).
 *   matrix  utils.c fill_rgb2yuv_table, the SWS_CS_DEFAULT (ITU-R 601) special case: (int)(c * 219 / 255 * 2^15 + .5)
 *           for luma, c * 224 / 255 for chroma, negative ones negated after rounding.
 */
static long long rounded_div(long long a, long long b) { return a >= 0 ? (a + (b >> 1)) / b : -((-a + (b >> 1)) / b); }
static int ilog2(unsigned v) { int n = 0; while (v >>= 1) n++; return n; }

/* The bilinear filter bank of one axis (the general branch of utils.c initFilter with SWS_BILINEAR, no source /
 * destination filter vectors): weights in 2^54 units, near-zero taps dropped (cut-off 0.002), out-of-picture taps
 * folded onto the border sample, then normalised to `one` with the rounding error carried from tap to tap. */
static int sws_bilinear_filter(swsfilter *f, int srcn, int dstn, int one, int srcpos, int dstpos) {
    const long long xinc = (((long long)srcn << 16) + (dstn >> 1)) / dstn;
    const int lg = ilog2((unsigned)(srcn / dstn > 0 ? srcn / dstn : 1));
    const long long fone = 1LL << (54 - (lg < 8 ? lg : 8));
    f->pos = (int *)malloc(sizeof(int) * (size_t)dstn);
    if (llabs(xinc - 0x10000) < 10 && srcpos == dstpos) {                 /* not scaled */
        f->size = 1;
        f->coef = (int *)malloc(sizeof(int) * (size_t)dstn);
        for (int i = 0; i < dstn; i++) { f->pos[i] = i; f->coef[i] = one; }
        return 0;
    }
    int fs = xinc <= (1 << 16) ? 3 : 1 + (int)((2LL * srcn + dstn - 1) / dstn);
    if (fs > srcn - 2) fs = srcn - 2;
    if (fs < 1) fs = 1;
    long long *w = (long long *)calloc((size_t)dstn * (size_t)fs, sizeof(long long));
    long long x = ((dstpos * xinc) >> 7) - (((long long)srcpos * 0x10000LL) >> 7);
    for (int i = 0; i < dstn; i++, x += 2 * xinc) {
        int xx = (int)((x - (long long)(fs - 2) * (1LL << 16)) / (1 << 17));    /* C division: towards zero */
        f->pos[i] = xx;
        for (int j = 0; j < fs; j++, xx++) {
            long long d = llabs((long long)xx * (1 << 17) - x) << 13;
            if (xinc > (1 << 16)) d = d * dstn / srcn;
            long long c = (1LL << 30) - d;
            w[(size_t)i * fs + j] = c < 0 ? 0 : c * (fone >> 30);
        }
    }
    const double cutoff = 0.002 * (double)fone;
    int minsize = 0;
    for (int i = dstn - 1; i >= 0; i--) {
        long long *r = w + (size_t)i * fs;
        long long acc = 0;
        int need = fs;
        for (int j = 0; j < fs; j++) {                                       /* near-zero taps on the left: shift them out */
            acc += llabs(r[0]);
            if ((double)acc > cutoff) break;
            if (i < dstn - 1 && f->pos[i] >= f->pos[i + 1]) break;           /* positions stay monotonic */
            for (int k = 1; k < fs; k++) r[k - 1] = r[k];
            r[fs - 1] = 0;
            f->pos[i]++;
        }
        acc = 0;
        for (int j = fs - 1; j > 0; j--) {                                   /* near-zero taps on the right */
            acc += llabs(r[j]);
            if ((double)acc > cutoff) break;
            need--;
        }
        if (need > minsize) minsize = need;
    }
    const int n = minsize;
    f->size = n;
    f->coef = (int *)malloc(sizeof(int) * (size_t)dstn * (size_t)n);
    for (int i = 0; i < dstn; i++) {
        long long *r = w + (size_t)i * fs;                                   /* only r[0..n-1] is used from here on */
        if (f->pos[i] < 0) {                                                 /* taps before the first sample */
            for (int j = 1; j < n; j++) {
                const int left = j + f->pos[i] > 0 ? j + f->pos[i] : 0;
                r[left] += r[j];
                r[j] = 0;
            }
            f->pos[i] = 0;
        }
        if (f->pos[i] + n > srcn) {                                          /* taps behind the last sample */
            const int shift = f->pos[i] + (n - srcn < 0 ? n - srcn : 0);
            long long acc = 0;
            for (int j = n - 1; j >= 0; j--)
                if (f->pos[i] + j >= srcn) { acc += r[j]; r[j] = 0; }
            for (int j = n - 1; j >= 0; j--) r[j] = j < shift ? 0 : r[j - shift];
            f->pos[i] -= shift;
            r[srcn - 1 - f->pos[i]] += acc;
        }
        long long sum = 0, err = 0;
        for (int j = 0; j < n; j++) sum += r[j];
        sum = (sum + one / 2) / one;
        if (!sum) sum = 1;
        for (int j = 0; j < n; j++) {
            const long long v = r[j] + err;
            const long long iv = rounded_div(v, sum);
            f->coef[(size_t)i * n + j] = (int)iv;
            err = v - iv * sum;
        }
    }
    free(w);
    return 0;
}
static void swsfilter_free(swsfilter *f) { free(f->pos); free(f->coef); }

static int sws_q(double c) { return (int)(c * (double)(1 << 15) + 0.5); }
static int to15(long long s14) { const long long v = (s14 * 16384) >> 13; return (int)(v < 32767 ? v : 32767); }

/* v420: 1 = 4:2:0, 0 = 4:2:2 */
int oracle_bgra_to_yuv(uint8_t *y, int ly, uint8_t *u, int lu, uint8_t *v, int lv, const uint8_t *bgra, int stride,
                       int w, int h, int v420) {
    if (!y || !u || !v || !bgra || w <= 0 || h <= 0) return -1;
    const int ry = sws_q(0.299 * 219 / 255), gy = sws_q(0.587 * 219 / 255), by = sws_q(0.114 * 219 / 255);
    const int ru = -sws_q(0.169 * 224 / 255), gu = -sws_q(0.331 * 224 / 255), bu = sws_q(0.500 * 224 / 255);
    const int rv = sws_q(0.500 * 224 / 255), gv = -sws_q(0.419 * 224 / 255), bv = -sws_q(0.081 * 224 / 255);
    for (int yy = 0; yy < h; yy++)
        for (int x = 0; x < w; x++) {
            const uint8_t *p = bgra + (size_t)yy * (size_t)stride + 4 * (size_t)x;
            const int y14 = (int)(((long long)ry * p[2] + (long long)gy * p[1] + (long long)by * p[0] + (16LL << 15) + 256) >> 9);
            y[(size_t)yy * (size_t)ly + x] = (uint8_t)clamp8((to15(y14) + 64) >> 7);
        }
    const int cw = (w + 1) / 2, chh = v420 ? (h + 1) / 2 : h;
    const int halfw = (w & 1) == 0;                       /* even width: chroma from pixel pairs, no horizontal filter */
    swsfilter hf, vf;
    if (!halfw) sws_bilinear_filter(&hf, w, cw, 1 << 14, 128, 128);
    sws_bilinear_filter(&vf, h, chh, 1 << 12, 128, 128);  /* 4:2:2: srcn == dstn, the unit filter */
    /* horizontally scaled chroma, 15 bits, every source row */
    int *hu = (int *)malloc(sizeof(int) * (size_t)cw * (size_t)h), *hv = (int *)malloc(sizeof(int) * (size_t)cw * (size_t)h);
    for (int yy = 0; yy < h; yy++) {
        const uint8_t *row = bgra + (size_t)yy * (size_t)stride;
        for (int cx = 0; cx < cw; cx++) {
            long long su, sv;
            if (halfw) {
                const uint8_t *p = row + 8 * (size_t)cx;
                const int r = p[2] + p[6], g = p[1] + p[5], b = p[0] + p[4];
                su = ((long long)ru * r + (long long)gu * g + (long long)bu * b + (256LL << 15) + 512) >> 10;
                sv = ((long long)rv * r + (long long)gv * g + (long long)bv * b + (256LL << 15) + 512) >> 10;
                hu[(size_t)yy * cw + cx] = to15(su);
                hv[(size_t)yy * cw + cx] = to15(sv);
            } else {
                long long au = 0, av = 0;
                for (int j = 0; j < hf.size; j++) {
                    const uint8_t *p = row + 4 * (size_t)(hf.pos[cx] + j);
                    su = ((long long)ru * p[2] + (long long)gu * p[1] + (long long)bu * p[0] + (256LL << 14) + 256) >> 9;
                    sv = ((long long)rv * p[2] + (long long)gv * p[1] + (long long)bv * p[0] + (256LL << 14) + 256) >> 9;
                    au += su * hf.coef[(size_t)cx * hf.size + j];
                    av += sv * hf.coef[(size_t)cx * hf.size + j];
                }
                au >>= 13; av >>= 13;
                hu[(size_t)yy * cw + cx] = (int)(au < 32767 ? au : 32767);
                hv[(size_t)yy * cw + cx] = (int)(av < 32767 ? av : 32767);
            }
        }
    }
    for (int cy = 0; cy < chh; cy++)
        for (int cx = 0; cx < cw; cx++) {
            int ou, ov;
            if (vf.size == 1) {
                ou = (hu[(size_t)vf.pos[cy] * cw + cx] + 64) >> 7;
                ov = (hv[(size_t)vf.pos[cy] * cw + cx] + 64) >> 7;
            } else {
                int au = 64 << 12, av = 64 << 12;
                for (int j = 0; j < vf.size; j++) {
                    au += hu[(size_t)(vf.pos[cy] + j) * cw + cx] * vf.coef[(size_t)cy * vf.size + j];
                    av += hv[(size_t)(vf.pos[cy] + j) * cw + cx] * vf.coef[(size_t)cy * vf.size + j];
                }
                ou = au >> 19; ov = av >> 19;
            }
            u[(size_t)cy * (size_t)lu + cx] = (uint8_t)clamp8(ou);
            v[(size_t)cy * (size_t)lv + cx] = (uint8_t)clamp8(ov);
        }
    free(hu); free(hv);
    if (!halfw) swsfilter_free(&hf);
    swsfilter_free(&vf);
    return 0;
}

/* the filter bank on its own, for tests of the product's table builder: returns the tap count, fills pos[dstn] and
 * coef[dstn * taps] (taps <= max_taps, else -1) */
int oracle_sws_bilinear_filter(int srcn, int dstn, int one, int *pos, int *coef, int max_taps) {
    swsfilter f;
    sws_bilinear_filter(&f, srcn, dstn, one, 128, 128);
    const int n = f.size;
    if (n <= max_taps) {
        for (int i = 0; i < dstn; i++) pos[i] = f.pos[i];
        for (int i = 0; i < dstn * n; i++) coef[i] = f.coef[i];
    }
    swsfilter_free(&f);
    return n <= max_taps ? n : -1;
}
