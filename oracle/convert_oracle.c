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

/* BGRA -> planar YUV, BT.601 limited range: 15-bit coefficients ROUNDED TO NEAREST
 *   Y = 16 + 219/255 (.299 R + .587 G + .114 B),  U = 128 + 224/255 (-.169 R - .331 G + .500 B),
 *   V = 128 + 224/255 (.500 R - .419 G - .081 B);
 * luma per pixel; a chroma sample from the SUM of the pixels it covers (2x1 for 4:2:2, 2x2 for 4:2:0; at odd
 * right / bottom edges the last column / row is repeated), rounded to nearest.  v420: 1 = 4:2:0, 0 = 4:2:2. */
static int q15(double c) { const double s = c * 32768.0; return (int)(s < 0 ? -(long long)(-s + 0.5) : (long long)(s + 0.5)); }
int oracle_bgra_to_yuv(uint8_t *y, int ly, uint8_t *u, int lu, uint8_t *v, int lv, const uint8_t *bgra, int stride,
                       int w, int h, int v420) {
    if (!y || !u || !v || !bgra || w <= 0 || h <= 0) return -1;
    const double ys = 219.0 / 255.0, cs = 224.0 / 255.0;
    const int ry = q15(0.299 * ys), gy = q15(0.587 * ys), by = q15(0.114 * ys);
    const int ru = q15(-0.169 * cs), gu = q15(-0.331 * cs), bu = q15(0.500 * cs);
    const int rv = q15(0.500 * cs), gv = q15(-0.419 * cs), bv = q15(-0.081 * cs);
    for (int yy = 0; yy < h; yy++)
        for (int x = 0; x < w; x++) {
            const uint8_t *p = bgra + (size_t)yy * (size_t)stride + 4 * (size_t)x;
            y[(size_t)yy * (size_t)ly + x] = (uint8_t)((ry * p[2] + gy * p[1] + by * p[0] + (16 << 15) + (1 << 14)) >> 15);
        }
    const int rows = v420 ? 2 : 1, cw = (w + 1) / 2, chh = v420 ? (h + 1) / 2 : h, cnt = 2 * rows;
    const int sh = 15 + (v420 ? 2 : 1);
    for (int cy = 0; cy < chh; cy++)
        for (int cx = 0; cx < cw; cx++) {
            int sr = 0, sg = 0, sb = 0;
            for (int r = 0; r < rows; r++)
                for (int k = 0; k < 2; k++) {
                    const int yy = cy * rows + r < h ? cy * rows + r : h - 1, x = 2 * cx + k < w ? 2 * cx + k : w - 1;
                    const uint8_t *p = bgra + (size_t)yy * (size_t)stride + 4 * (size_t)x;
                    sr += p[2]; sg += p[1]; sb += p[0];
                }
            u[(size_t)cy * (size_t)lu + cx] = (uint8_t)((ru * sr + gu * sg + bu * sb + ((128 * cnt) << 15) + (1 << (sh - 1))) >> sh);
            v[(size_t)cy * (size_t)lv + cx] = (uint8_t)((rv * sr + gv * sg + bv * sb + ((128 * cnt) << 15) + (1 << (sh - 1))) >> sh);
        }
    return 0;
}
