/* convert_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the two picture conversions that sit either side of the hot path in the reference's field
 * loop and that the product runs on the device:
 *   - InputFile::frame_copy_scale()  (ffmpeg_ntsc.cpp:544-613):  decoder picture -> BGRA at the output size,
 *     sws_scale() with SWS_BILINEAR (:574-585, :603-610);
 *   - the encoder-side sws_scale()   (ffmpeg_ntsc.cpp:2266-2274, context :2118-2131, SMPTE170M / MPEG range
 *     :2100-2101):  finished BGRA picture -> planar YUV 4:2:0 / 4:2:2.
 *
 * PARITY UNPINNED: both are calls into libswscale, a third-party dependency that is absent from this environment
 * (no FFmpeg headers, libraries or binary; the reference needs FFmpeg 3.x).  What is restated here is the algorithm
 * as this repository specifies it (include/cvs_ntsc.h, "picture conversions"), written from that text and NOT from
 * the kernels, so that the kernels are checked against something they were not derived from.  It is the swscale
 * family of algorithms (triangle-kernel resampling with 14-bit weights and a 15-bit intermediate, BT.601 integer
 * matrices) but bit-equality with any libswscale build is not claimed.
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

/* format: 0 BGRA, 1 YUV420P, 2 YUV422P, 3 NV12 (the product's enum); dst: BGRA */
int oracle_scale_to_bgra(uint8_t *dst, int dst_stride, int dw, int dh,
                         const uint8_t *p0, const uint8_t *p1, const uint8_t *p2, int l0, int l1, int l2,
                         int sw, int sh, int format) {
    if (!dst || !p0 || dw <= 0 || dh <= 0 || sw <= 0 || sh <= 0 || format < 0 || format > 3) return -1;
    if (sw > 16 * dw || sh > 16 * dh) return -5;          /* more taps than the tables hold */
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
typedef struct { int size; int *pos; int *coef; } swsfilter;      /* coef[i * size + j] applies to sample pos[i] + j */

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
