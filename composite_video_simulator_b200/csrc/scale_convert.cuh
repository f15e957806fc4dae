// scale_convert.cuh -- the input side of the reference's field loop on the device: decoder picture -> BGRA at the
// output size, InputFile::frame_copy_scale() (ffmpeg_ntsc.cpp:544-613: sws_getContext(src w, h, format -> output
// w, h, BGRA, SWS_BILINEAR) at :574-585, sws_scale() at :603-610).  SURVEY section 8f-1.
//
// The kernels:
//
//  k_sws_yuv_to_bgra   YUV420P / YUV422P / NV12 sources: the bytes of libswscale's C code (9.1.100;
//      = SWS_ACCURATE_RND | SWS_BITEXACT).  PINNED: oracle/convert_oracle.c restates the library and is compared with it
//      byte for byte (tests/test_swscale_pin.py); the kernel is compared with the oracle, the library and its committed
//      outputs (tests/test_gpu_scale.py).  All integer:
//        horizontal   every source row through the library's bilinear banks (sws_filter.cpp, 14-bit weights): Y to dw
//                     samples, U and V to dw/2 samples (ONE chroma sample per pair of output pixels),
//                     s15 = min(sum >> 7, 32767);
//        vertical     banks with 12-bit weights (luma rows -> dh, chroma rows -> dh); the combination depends on the banks'
//                     tap counts exactly as in the library:  1/1 taps: (s + 64) >> 7;  1 luma / 2 chroma: chroma
//                     (c0 (4096 - a) + c1 a + (128 << 11)) >> 19;  2/2: (s0 (4096 - a) + s1 a) >> 19 without a rounding
//                     term;  otherwise ((1 << 18) + sum s w) >> 19;
//        colour       the library's ITU-R 601 tables as arithmetic: cy = 65536 * 255 / 219, c' = (c 65536 + 32768) / cy for
//                     c = 104597 (V->R), 132201 (U->B), -25675 (U->G), -53279 (V->G); k = Y + (c' C >> 16) - (c' >> 9);
//                     channel = clip8((k cy - (400 << 16) + 326 cy + 32768) >> 16); alpha 255;
//        exception    YUV420P at the same size and an even height: the library's direct converter, no filtering, the chroma
//                     sample of a 2x2 block serves its four pixels.
//      One thread = one pair of output pixels; it filters the few source samples it needs itself (2 x 2 taps when
//      enlarging), so no intermediate picture exists.  Odd output widths take the library's other writers
//      (k_sws_yuv_to_bgra_full, further down).
//
//  k_sws_bgra_to_bgra  BGRA sources at another size: the library's RGB -> YUV(A) -> RGB route (further down), pinned too.
//
//  k_scale_to_bgra     BGRA sources of the output size (a copy, as in the library) and the ONE geometry the pinned kernels do
//      not take yet -- a BGRA source of odd width reduced to half its width or less (the library keeps chroma per pixel
//      there; known too late in the round to route and verify): the repository's OWN resampler, NOT pinned, specified here:
//  * per axis, destination sample i of n_dst takes its value at source position P / D (centre aligned),
//        D = 2 n_dst sub,   P = (2 i + 1) n_src - n_dst - off n_dst,
//    n_src the LUMA size of the source along the axis, sub = 1 for luma / BGRA planes and 2 for a subsampled chroma
//    axis, off = 1 for vertically centred 4:2:0 chroma (MPEG-1/2 siting), else 0 (chroma co-sited with even luma
//    columns horizontally);
//  * kernel half-width H = max(D, 2 n_src) in units of 1 / D: two taps when enlarging, a wider triangle when
//    shrinking; source sample j weighs t_j = H - |j D - P| when that is positive; indices are clamped to the plane;
//  * weights in 14 bits: w_j = floor(16384 t_j / sum t), the remainder goes to the largest t_j (the first of equals);
//  * horizontal pass h = (sum w_j s_j + 64) >> 7 (15 bits), vertical pass v = clamp8((sum w_k h_k + 2^20) >> 21);
//  * planar YUV input is scaled plane by plane to full resolution and then converted, BT.601 limited range ->
//    full-range RGB:  c = 298 (Y - 16),  R = clamp8((c + 409 (V-128) + 128) >> 8),
//    G = clamp8((c - 100 (U-128) - 208 (V-128) + 128) >> 8),  B = clamp8((c + 516 (U-128) + 128) >> 8),  A = 255;
//    BGRA input is scaled channel by channel (alpha included); a BGRA source of the output size is copied, as the
//    library does.
// Tap tables are built on the host once per geometry and shared by all pictures.
#ifndef CVS_SCALE_CONVERT_CUH
#define CVS_SCALE_CONVERT_CUH

#include <stdint.h>

#include <vector>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif

namespace cvs {

constexpr int kScaleOne = 16384;

struct ScaleAxis {              // host copy of one axis' tap table
    int n_dst = 0, taps = 0;    // taps per destination sample (padded with zero weights)
    std::vector<int32_t> first; // first source index (unclamped) of sample i
    std::vector<int16_t> w;     // [n_dst][taps]
};

// the tap table of one axis (see the header comment); n_plane = samples of the source plane along the axis
inline void scale_build_axis(int n_dst, int n_src_luma, int sub, int off, ScaleAxis &ax) {
    const long long D = 2LL * n_dst * sub;
    const long long H = D > 2LL * n_src_luma ? D : 2LL * n_src_luma;
    ax.n_dst = n_dst;
    ax.taps = (int)((2 * H + D - 1) / D) + 1;
    ax.first.assign((size_t)n_dst, 0);
    ax.w.assign((size_t)n_dst * (size_t)ax.taps, 0);
    std::vector<long long> t((size_t)ax.taps);
    for (int i = 0; i < n_dst; i++) {
        const long long P = (2LL * i + 1) * n_src_luma - n_dst - (long long)off * n_dst;
        // smallest j with j D - P > -H:  j > (P - H) / D
        long long num = P - H, j0 = num >= 0 ? num / D + 1 : -((-num) / D) + (((-num) % D) == 0 ? 1 : 0);
        long long sum = 0;
        int best = 0;
        for (int k = 0; k < ax.taps; k++) {
            const long long d = (j0 + k) * D - P;
            const long long tv = H - (d < 0 ? -d : d);
            t[(size_t)k] = tv > 0 ? tv : 0;
            sum += t[(size_t)k];
            if (t[(size_t)k] > t[(size_t)best]) best = k;
        }
        long long acc = 0;
        for (int k = 0; k < ax.taps; k++) {
            const long long wq = (t[(size_t)k] * kScaleOne) / sum;
            ax.w[(size_t)i * ax.taps + k] = (int16_t)wq;
            acc += wq;
        }
        ax.w[(size_t)i * ax.taps + best] = (int16_t)(ax.w[(size_t)i * ax.taps + best] + (kScaleOne - acc));
        ax.first[(size_t)i] = (int32_t)j0;
    }
}

enum { SCALE_BGRA = 0, SCALE_YUV420P = 1, SCALE_YUV422P = 2, SCALE_NV12 = 3 };

struct ScaleArgs {
    uint8_t *dst;                       // BGRA
    long long dst_pic_stride;
    int dst_stride, dw, dh;
    const uint8_t *src[3];              // BGRA: [0]; planar: Y, U, V; NV12: Y, UV, -
    long long src_pic_stride[3];
    int src_linesize[3];
    int sw, sh, cw, ch;                 // luma / chroma plane sizes
    int format, n;
    // device tap tables: luma x / y, chroma x / y
    const int32_t *fx, *fy, *cfx, *cfy;
    const int16_t *wx, *wy, *cwx, *cwy;
    int tx, ty, ctx_, cty;              // taps per sample
};

#if defined(__CUDACC__)
__device__ __forceinline__ int scale_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// one sample of one plane: vertical taps of horizontally filtered rows; `step` = bytes between samples of the plane
// in a row (1 planar, 2 for the U / V of NV12, 4 for a BGRA channel)
__device__ __forceinline__ int scale_sample(const uint8_t *plane, int linesize, int pw, int ph, int step,
                                            const int32_t *fx, const int16_t *wx, int tx, int x,
                                            const int32_t *fy, const int16_t *wy, int ty, int y) {
    const int x0 = fx[x], y0 = fy[y];
    const int16_t *wxr = wx + (size_t)x * tx, *wyr = wy + (size_t)y * ty;
    int acc = 1 << 20;
    for (int k = 0; k < ty; k++) {
        const int wv = wyr[k];
        if (wv == 0) continue;
        const uint8_t *row = plane + (size_t)scale_clampi(y0 + k, 0, ph - 1) * (size_t)linesize;
        int h = 64;
        for (int j = 0; j < tx; j++) {
            const int wh = wxr[j];
            if (wh != 0) h += wh * (int)row[(size_t)scale_clampi(x0 + j, 0, pw - 1) * (size_t)step];
        }
        acc += wv * (h >> 7);
    }
    return scale_clampi(acc >> 21, 0, 255);
}

// one thread = one destination pixel (the tap tables and the few source rows a block touches stay in L1 / L2); BGRA sources
// only: the same size (the weights are then the identity: a copy) and the one geometry the pinned kernels do not take
__global__ void __launch_bounds__(256) k_scale_to_bgra(const __grid_constant__ ScaleArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, k = blockIdx.z;
    if (x >= a.dw) return;
    uint32_t out;
    {
        const uint8_t *p = a.src[0] + (long long)k * a.src_pic_stride[0];
        out = 0;
        for (int c = 0; c < 4; c++)
            out |= (uint32_t)scale_sample(p + c, a.src_linesize[0], a.sw, a.sh, 4, a.fx, a.wx, a.tx, x, a.fy, a.wy, a.ty, y) << (8 * c);
    }
    uint32_t *drow = reinterpret_cast<uint32_t *>(a.dst + (long long)k * a.dst_pic_stride + (long long)y * a.dst_stride);
    drow[x] = out;
}

// ---- libswscale's bytes for planar YUV sources (see the header comment) --------------------------------------------
struct SwsScaleArgs {
    uint8_t *dst;
    long long dst_pic_stride;
    int dst_stride, dw, dh;
    const uint8_t *y, *u, *v;           // NV12: u = the interleaved plane, v = u + 1
    long long sp_y, sp_c;               // picture strides
    int ly, lc, cstep;                  // line sizes; bytes between chroma samples (1 planar, 2 NV12)
    int direct;                         // same-size YUV420P, even height: no filtering
    // banks: luma columns, chroma columns, luma rows, chroma rows
    const int32_t *hl_pos, *hl_w, *hc_pos, *hc_w, *vl_pos, *vl_w, *vc_pos, *vc_w;
    int hl_t, hc_t, vl_t, vc_t;
    int n;
};

__device__ __forceinline__ int sws_h15(const uint8_t *row, int step, const int32_t *w, int pos, int taps) {
    int acc = 0;
    for (int j = 0; j < taps; j++) acc += (int)row[(size_t)(pos + j) * (size_t)step] * w[j];
    return min(acc >> 7, 32767);
}
__device__ __forceinline__ int sws_chan(int k) {
    constexpr int cy = (65536 * 255) / 219;
    return scale_clampi((k * cy - (400 << 16) + 326 * cy + 32768) >> 16, 0, 255);
}
__device__ __forceinline__ uint32_t sws_pixel(int Y, int U, int V) {
    constexpr long long cy = (65536LL * 255) / 219;
    constexpr int crv = (int)((104597LL * 65536 + 32768) / cy), cbu = (int)((132201LL * 65536 + 32768) / cy);
    constexpr int cgu = -(int)((25675LL * 65536 - 32768) / cy), cgv = -(int)((53279LL * 65536 - 32768) / cy);
    U = scale_clampi(U, 0, 255);
    V = scale_clampi(V, 0, 255);
    const int r = sws_chan(Y + ((crv * V) >> 16) - (crv >> 9));
    const int b = sws_chan(Y + ((cbu * U) >> 16) - (cbu >> 9));
    const int g = sws_chan(Y + ((cgu * U) >> 16) - (cgu >> 9) + ((cgv * V) >> 16) - (cgv >> 9));
    return 0xFF000000u | ((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b;
}

__global__ void __launch_bounds__(256) k_sws_yuv_to_bgra(const __grid_constant__ SwsScaleArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, k = blockIdx.z;     // i: pair of output pixels
    if (2 * i >= a.dw) return;
    const uint8_t *py = a.y + (long long)k * a.sp_y, *pu = a.u + (long long)k * a.sp_c, *pv = a.v + (long long)k * a.sp_c;
    uint32_t *out = reinterpret_cast<uint32_t *>(a.dst + (long long)k * a.dst_pic_stride + (long long)y * a.dst_stride) + 2 * i;
    if (a.direct) {
        const uint8_t *ry = py + (size_t)y * a.ly + 2 * i;
        const int U = pu[(size_t)(y >> 1) * a.lc + (size_t)i * a.cstep], V = pv[(size_t)(y >> 1) * a.lc + (size_t)i * a.cstep];
        out[0] = sws_pixel(ry[0], U, V);
        out[1] = sws_pixel(ry[1], U, V);
        return;
    }
    const int x0 = 2 * i, x1 = 2 * i + 1;
    const int32_t *wl0 = a.hl_w + (size_t)x0 * a.hl_t, *wl1 = a.hl_w + (size_t)x1 * a.hl_t, *wc = a.hc_w + (size_t)i * a.hc_t;
    const int pl0 = a.hl_pos[x0], pl1 = a.hl_pos[x1], pc = a.hc_pos[i];
    const int32_t *vl = a.vl_w + (size_t)y * a.vl_t, *vc = a.vc_w + (size_t)y * a.vc_t;
    const int rl = a.vl_pos[y], rc = a.vc_pos[y];
    const bool two_c = a.vc_t == 2 && vc[0] + vc[1] == 4096 && vc[1] >= 0 && vc[1] <= 4096;
    const bool two_l = a.vl_t == 2 && vl[0] + vl[1] == 4096 && vl[1] >= 0 && vl[1] <= 4096;
    auto L = [&](int r, int which) {
        const uint8_t *row = py + (size_t)(rl + r) * a.ly;
        return which ? sws_h15(row, 1, wl1, pl1, a.hl_t) : sws_h15(row, 1, wl0, pl0, a.hl_t);
    };
    auto CU = [&](int r) { return sws_h15(pu + (size_t)(rc + r) * a.lc, a.cstep, wc, pc, a.hc_t); };
    auto CV = [&](int r) { return sws_h15(pv + (size_t)(rc + r) * a.lc, a.cstep, wc, pc, a.hc_t); };
    int Y0, Y1, U, V;
    if (a.vl_t == 1 && (a.vc_t == 1 || two_c)) {
        Y0 = (L(0, 0) + 64) >> 7;
        Y1 = (L(0, 1) + 64) >> 7;
        if (a.vc_t == 1) {
            U = (CU(0) + 64) >> 7;
            V = (CV(0) + 64) >> 7;
        } else {
            const int c = vc[1];
            U = (CU(0) * (4096 - c) + CU(1) * c + (128 << 11)) >> 19;
            V = (CV(0) * (4096 - c) + CV(1) * c + (128 << 11)) >> 19;
        }
    } else if (two_l && two_c) {
        const int l = vl[1], c = vc[1];
        Y0 = (L(0, 0) * (4096 - l) + L(1, 0) * l) >> 19;
        Y1 = (L(0, 1) * (4096 - l) + L(1, 1) * l) >> 19;
        U = (CU(0) * (4096 - c) + CU(1) * c) >> 19;
        V = (CV(0) * (4096 - c) + CV(1) * c) >> 19;
    } else {
        Y0 = Y1 = U = V = 1 << 18;
        for (int j = 0; j < a.vl_t; j++) {
            const int w = vl[j];
            if (w == 0) continue;
            Y0 += L(j, 0) * w;
            Y1 += L(j, 1) * w;
        }
        for (int j = 0; j < a.vc_t; j++) {
            const int w = vc[j];
            if (w == 0) continue;
            U += CU(j) * w;
            V += CV(j) * w;
        }
        Y0 >>= 19; Y1 >>= 19; U >>= 19; V >>= 19;
    }
    out[0] = sws_pixel(Y0, U, V);
    out[1] = sws_pixel(Y1, U, V);
}

// Odd output widths: the library turns on full horizontal chroma interpolation -- chroma is scaled to dw samples, one per
// pixel -- and writes with 32-bit integer arithmetic instead of its tables (output.c yuv2rgb_full_{1,2,X}_c_template +
// yuv2rgb_write_full): samples at 2^9 per code, chroma minus 128;
//     Y' = (Y - 8192) 9539 + 2^21;  R = Y' + 13075 V;  G = Y' - 6660 V - 3209 U;  B = Y' + 16525 U     (mod 2^32),
// clipped to 30 bits when any of the three leaves that range, >> 22.  The sums wrap at 32 bits BEFORE the clip, so a
// far-out-of-gamut sample (Y = U = 255) comes out 0 where 255 would be expected: the library's behaviour, reproduced.
// One thread = one output pixel.  (Encoders do not take odd widths; this exists so that every size gives the library's bytes.)
__device__ __forceinline__ int sws_clip30(int a) { return (a & ~((1 << 30) - 1)) ? ((~a) >> 31) & ((1 << 30) - 1) : a; }
__device__ __forceinline__ uint32_t sws_pixel_full(int Y, int U, int V) {
    const uint32_t y = (uint32_t)(Y - 8192) * 9539u + (1u << 21);
    int R = (int)(y + (uint32_t)V * 13075u);
    int G = (int)(y + (uint32_t)V * (uint32_t)-6660 + (uint32_t)U * (uint32_t)-3209);
    int B = (int)(y + (uint32_t)U * 16525u);
    if ((R | G | B) & 0xC0000000) { R = sws_clip30(R); G = sws_clip30(G); B = sws_clip30(B); }
    return 0xFF000000u | ((uint32_t)(R >> 22) << 16) | ((uint32_t)(G >> 22) << 8) | (uint32_t)(B >> 22);
}
__global__ void __launch_bounds__(256) k_sws_yuv_to_bgra_full(const __grid_constant__ SwsScaleArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, k = blockIdx.z;
    if (x >= a.dw) return;
    const uint8_t *py = a.y + (long long)k * a.sp_y, *pu = a.u + (long long)k * a.sp_c, *pv = a.v + (long long)k * a.sp_c;
    const int32_t *wl = a.hl_w + (size_t)x * a.hl_t, *wc = a.hc_w + (size_t)x * a.hc_t;      // the chroma bank has dw entries here
    const int pl = a.hl_pos[x], pc = a.hc_pos[x];
    const int32_t *vl = a.vl_w + (size_t)y * a.vl_t, *vc = a.vc_w + (size_t)y * a.vc_t;
    const int rl = a.vl_pos[y], rc = a.vc_pos[y];
    const bool two_c = a.vc_t == 2 && vc[0] + vc[1] == 4096 && vc[1] >= 0 && vc[1] <= 4096;
    const bool two_l = a.vl_t == 2 && vl[0] + vl[1] == 4096 && vl[1] >= 0 && vl[1] <= 4096;
    auto L = [&](int r) { return sws_h15(py + (size_t)(rl + r) * a.ly, 1, wl, pl, a.hl_t); };
    auto CU = [&](int r) { return sws_h15(pu + (size_t)(rc + r) * a.lc, a.cstep, wc, pc, a.hc_t); };
    auto CV = [&](int r) { return sws_h15(pv + (size_t)(rc + r) * a.lc, a.cstep, wc, pc, a.hc_t); };
    int Y, U, V;
    if (a.vl_t == 1 && (a.vc_t == 1 || two_c)) {
        Y = L(0) * 4;
        if (a.vc_t == 1) {
            U = (CU(0) - (128 << 7)) * 4;
            V = (CV(0) - (128 << 7)) * 4;
        } else {
            const int c = vc[1];
            U = (CU(0) * (4096 - c) + CU(1) * c - (128 << 19)) >> 10;
            V = (CV(0) * (4096 - c) + CV(1) * c - (128 << 19)) >> 10;
        }
    } else if (two_l && two_c) {
        const int l = vl[1], c = vc[1];
        Y = (L(0) * (4096 - l) + L(1) * l) >> 10;
        U = (CU(0) * (4096 - c) + CU(1) * c - (128 << 19)) >> 10;
        V = (CV(0) * (4096 - c) + CV(1) * c - (128 << 19)) >> 10;
    } else {
        Y = 1 << 9;
        U = V = (1 << 9) - (128 << 19);
        for (int j = 0; j < a.vl_t; j++) if (vl[j] != 0) Y += L(j) * vl[j];
        for (int j = 0; j < a.vc_t; j++) if (vc[j] != 0) { U += CU(j) * vc[j]; V += CV(j) * vc[j]; }
        Y >>= 10; U >>= 10; V >>= 10;
    }
    reinterpret_cast<uint32_t *>(a.dst + (long long)k * a.dst_pic_stride + (long long)y * a.dst_stride)[x] = sws_pixel_full(Y, U, V);
}

// BGRA sources at another size.  Packed RGB on both sides makes the library go to YUV(A) at the precision of its 15-bit
// intermediates and back through the same full-chroma writers, alpha as a fourth plane:
//   input       y14 = (RY r + GY g + BY b + (16 << 15) + 256) >> 9, chroma per PIXEL c14 = (RU r + .. + (256 << 14) + 256) >> 9,
//               a14 = a << 6 | a >> 2; when the width shrinks to half or less (dw <= sw / 2) chroma comes from the SUM of a pixel
//               pair instead, (.. + (256 << 15) + 512) >> 10;
//   horizontal  min(sum >> 13, 32767) for Y, U, V, A alike (banks sw -> dw; chroma: its own width);
//   vertical    ONE bank sh -> dh for all four planes, writer by its tap count: 1 tap: Y = s 4, C = (c - (128 << 7)) 4,
//               A = (a + 64) >> 7;  2 taps: (s0 (4096 - w) + s1 w) >> 10 without a rounding term, A = (.. + 2^18) >> 19;
//               otherwise (2^9 + sum) >> 10, A = (2^18 + sum) >> 19;  colour by sws_pixel_full, alpha clipped to a byte.
// One thread = one output pixel, the RGB -> YUV of its taps on the fly.  (A source of the output size never comes here:
// it is a copy.  An odd source width reduced to half or less is the one geometry left on k_scale_to_bgra.)
struct SwsRgbArgs {
    uint8_t *dst;
    long long dst_pic_stride;
    int dst_stride, dw, dh;
    const uint8_t *src;
    long long sp;
    int ls, half;
    const int32_t *hl_pos, *hl_w, *hc_pos, *hc_w, *v_pos, *v_w;
    int hl_t, hc_t, v_t;
    int ry, gy, by, ru, gu, bu, rv, gv, bv;
};
__global__ void __launch_bounds__(256) k_sws_bgra_to_bgra(const __grid_constant__ SwsRgbArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, k = blockIdx.z;
    if (x >= a.dw) return;
    const uint8_t *src = a.src + (long long)k * a.sp;
    const int32_t *wl = a.hl_w + (size_t)x * a.hl_t, *wc = a.hc_w + (size_t)x * a.hc_t, *wv = a.v_w + (size_t)y * a.v_t;
    const int pl = a.hl_pos[x], pc = a.hc_pos[x], r0 = a.v_pos[y];
    // the four horizontally filtered 15-bit samples of source row r
    auto row15 = [&](int r, int &Y, int &U, int &V, int &A) {
        const uint8_t *rowb = src + (size_t)r * a.ls;
        auto px = [&](int i) {                     // (byte loads: a BGRA row need not be 4-byte aligned)
            const uint8_t *q = rowb + 4 * (size_t)i;
            return (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
        };
        int sy = 0, sa = 0, su = 0, sv = 0;
        for (int j = 0; j < a.hl_t; j++) {
            const uint32_t p = px(pl + j);
            const int b = p & 0xFF, g = (p >> 8) & 0xFF, rr = (p >> 16) & 0xFF, al = p >> 24;
            sy += ((a.ry * rr + a.gy * g + a.by * b + (16 << 15) + 256) >> 9) * wl[j];
            sa += ((al << 6) | (al >> 2)) * wl[j];
        }
        for (int j = 0; j < a.hc_t; j++) {
            int cu, cv;
            if (a.half) {
                const uint32_t p = px(2 * (pc + j)), q = px(2 * (pc + j) + 1);
                const int b = (p & 0xFF) + (q & 0xFF), g = ((p >> 8) & 0xFF) + ((q >> 8) & 0xFF), rr = ((p >> 16) & 0xFF) + ((q >> 16) & 0xFF);
                cu = (a.ru * rr + a.gu * g + a.bu * b + (256 << 15) + 512) >> 10;
                cv = (a.rv * rr + a.gv * g + a.bv * b + (256 << 15) + 512) >> 10;
            } else {
                const uint32_t p = px(pc + j);
                const int b = p & 0xFF, g = (p >> 8) & 0xFF, rr = (p >> 16) & 0xFF;
                cu = (a.ru * rr + a.gu * g + a.bu * b + (256 << 14) + 256) >> 9;
                cv = (a.rv * rr + a.gv * g + a.bv * b + (256 << 14) + 256) >> 9;
            }
            su += cu * wc[j];
            sv += cv * wc[j];
        }
        Y = min(sy >> 13, 32767); A = min(sa >> 13, 32767); U = min(su >> 13, 32767); V = min(sv >> 13, 32767);
    };
    int Y, U, V, A;
    if (a.v_t == 1) {
        int y0, u0, v0, a0;
        row15(r0, y0, u0, v0, a0);
        Y = y0 * 4; U = (u0 - (128 << 7)) * 4; V = (v0 - (128 << 7)) * 4;
        A = (a0 + 64) >> 7;
    } else if (a.v_t == 2 && wv[0] + wv[1] == 4096 && wv[1] >= 0 && wv[1] <= 4096) {
        int y0, u0, v0, a0, y1, u1, v1, a1;
        row15(r0, y0, u0, v0, a0);
        row15(r0 + 1, y1, u1, v1, a1);
        const int w1 = wv[1], w0 = 4096 - w1;
        Y = (y0 * w0 + y1 * w1) >> 10;
        U = (u0 * w0 + u1 * w1 - (128 << 19)) >> 10;
        V = (v0 * w0 + v1 * w1 - (128 << 19)) >> 10;
        A = (a0 * w0 + a1 * w1 + (1 << 18)) >> 19;
    } else {
        Y = 1 << 9; U = V = (1 << 9) - (128 << 19); A = 1 << 18;
        for (int j = 0; j < a.v_t; j++) {
            if (wv[j] == 0) continue;
            int yj, uj, vj, aj;
            row15(r0 + j, yj, uj, vj, aj);
            Y += yj * wv[j]; U += uj * wv[j]; V += vj * wv[j]; A += aj * wv[j];
        }
        Y >>= 10; U >>= 10; V >>= 10; A >>= 19;
    }
    reinterpret_cast<uint32_t *>(a.dst + (long long)k * a.dst_pic_stride + (long long)y * a.dst_stride)[x] =
        (sws_pixel_full(Y, U, V) & 0x00FFFFFFu) | ((uint32_t)scale_clampi(A, 0, 255) << 24);
}

// The same conversion for the enlarging / same-size geometries (at most two taps per axis and plane), where neighbouring
// output rows share their source rows: one thread = one pair of output pixels over kSwsRows consecutive output rows.
// Its horizontal taps stay in registers, and the horizontally filtered samples of the two most recent source rows are
// kept (the vertical taps of the next output row are the same rows or the next one), so every source row is filtered
// once per thread instead of once per output row.  Bit for bit the arithmetic of k_sws_yuv_to_bgra.
constexpr int kSwsRows = 8;
__global__ void __launch_bounds__(256) k_sws_yuv_to_bgra_rows(const __grid_constant__ SwsScaleArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.z;
    if (2 * i >= a.dw) return;
    const int ya = blockIdx.y * kSwsRows, yb = min(ya + kSwsRows, a.dh);
    const uint8_t *py = a.y + (long long)k * a.sp_y, *pu = a.u + (long long)k * a.sp_c, *pv = a.v + (long long)k * a.sp_c;
    // horizontal taps of the two pixels and of their chroma sample; a missing second tap re-reads the first with weight 0
    const int x0 = 2 * i, x1 = 2 * i + 1;
    const int pl0 = a.hl_pos[x0], pl1 = a.hl_pos[x1], pc = a.hc_pos[i];
    const int wl0a = a.hl_w[(size_t)x0 * a.hl_t], wl0b = a.hl_t > 1 ? a.hl_w[(size_t)x0 * a.hl_t + 1] : 0;
    const int wl1a = a.hl_w[(size_t)x1 * a.hl_t], wl1b = a.hl_t > 1 ? a.hl_w[(size_t)x1 * a.hl_t + 1] : 0;
    const int wca = a.hc_w[(size_t)i * a.hc_t], wcb = a.hc_t > 1 ? a.hc_w[(size_t)i * a.hc_t + 1] : 0;
    const int dl = a.hl_t > 1 ? 1 : 0, dc = a.hc_t > 1 ? a.cstep : 0;
    // the two most recent source rows, horizontally filtered (scalars: indexed arrays would live in local memory)
    int lr0 = -1, lr1 = -1, la0 = 0, la1 = 0, lb0 = 0, lb1 = 0;      // luma: source row, pixel 0, pixel 1
    int cr0 = -1, cr1 = -1, cu0 = 0, cu1 = 0, cv0 = 0, cv1 = 0;      // chroma
    auto luma = [&](int r, int &A, int &B) {
        if (lr0 == r) { A = la0; B = lb0; return; }
        if (lr1 == r) { A = la1; B = lb1; return; }
        const uint8_t *row = py + (size_t)r * a.ly;
        A = min(((int)row[pl0] * wl0a + (int)row[pl0 + dl] * wl0b) >> 7, 32767);
        B = min(((int)row[pl1] * wl1a + (int)row[pl1 + dl] * wl1b) >> 7, 32767);
        if (lr0 <= lr1) { lr0 = r; la0 = A; lb0 = B; }             // replace the older row
        else { lr1 = r; la1 = A; lb1 = B; }
    };
    auto chroma = [&](int r, int &U, int &V) {
        if (cr0 == r) { U = cu0; V = cv0; return; }
        if (cr1 == r) { U = cu1; V = cv1; return; }
        const uint8_t *ru = pu + (size_t)r * a.lc + (size_t)pc * a.cstep, *rv = pv + (size_t)r * a.lc + (size_t)pc * a.cstep;
        U = min(((int)ru[0] * wca + (int)ru[dc] * wcb) >> 7, 32767);
        V = min(((int)rv[0] * wca + (int)rv[dc] * wcb) >> 7, 32767);
        if (cr0 <= cr1) { cr0 = r; cu0 = U; cv0 = V; }
        else { cr1 = r; cu1 = U; cv1 = V; }
    };
    for (int y = ya; y < yb; y++) {
        const int32_t *vl = a.vl_w + (size_t)y * a.vl_t, *vc = a.vc_w + (size_t)y * a.vc_t;
        const int rl = a.vl_pos[y], rc = a.vc_pos[y];
        const bool two_c = a.vc_t == 2 && vc[0] + vc[1] == 4096 && vc[1] >= 0 && vc[1] <= 4096;
        const bool two_l = a.vl_t == 2 && vl[0] + vl[1] == 4096 && vl[1] >= 0 && vl[1] <= 4096;
        int Y0, Y1, U, V;
        if (a.vl_t == 1 && (a.vc_t == 1 || two_c)) {
            int A, B, u0, v0;
            luma(rl, A, B);
            Y0 = (A + 64) >> 7;
            Y1 = (B + 64) >> 7;
            chroma(rc, u0, v0);
            if (a.vc_t == 1) {
                U = (u0 + 64) >> 7;
                V = (v0 + 64) >> 7;
            } else {
                int u1, v1;
                chroma(rc + 1, u1, v1);
                const int c = vc[1];
                U = (u0 * (4096 - c) + u1 * c + (128 << 11)) >> 19;
                V = (v0 * (4096 - c) + v1 * c + (128 << 11)) >> 19;
            }
        } else if (two_l && two_c) {
            int A0, B0, A1, B1, u0, v0, u1, v1;
            luma(rl, A0, B0);
            luma(rl + 1, A1, B1);
            chroma(rc, u0, v0);
            chroma(rc + 1, u1, v1);
            const int l = vl[1], c = vc[1];
            Y0 = (A0 * (4096 - l) + A1 * l) >> 19;
            Y1 = (B0 * (4096 - l) + B1 * l) >> 19;
            U = (u0 * (4096 - c) + u1 * c) >> 19;
            V = (v0 * (4096 - c) + v1 * c) >> 19;
        } else {
            Y0 = Y1 = U = V = 1 << 18;
            for (int j = 0; j < a.vl_t; j++) {
                int A, B;
                luma(rl + j, A, B);
                Y0 += A * vl[j];
                Y1 += B * vl[j];
            }
            for (int j = 0; j < a.vc_t; j++) {
                int u0, v0;
                chroma(rc + j, u0, v0);
                U += u0 * vc[j];
                V += v0 * vc[j];
            }
            Y0 >>= 19; Y1 >>= 19; U >>= 19; V >>= 19;
        }
        uint32_t *out = reinterpret_cast<uint32_t *>(a.dst + (long long)k * a.dst_pic_stride + (long long)y * a.dst_stride) + 2 * i;
        out[0] = sws_pixel(Y0, U, V);
        out[1] = sws_pixel(Y1, U, V);
    }
}
#endif

}  // namespace cvs
#endif
