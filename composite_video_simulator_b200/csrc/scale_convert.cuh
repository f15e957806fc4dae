// scale_convert.cuh -- the input side of the reference's field loop on the device: decoder picture -> BGRA at the
// output size, InputFile::frame_copy_scale() (ffmpeg_ntsc.cpp:544-613: sws_getContext(src w, h, format -> output
// w, h, BGRA, SWS_BILINEAR) at :574-585, sws_scale() at :603-610).  SURVEY section 8f-1.
//
// libswscale is a third-party dependency that is absent from this environment (no FFmpeg headers, libraries or
// binary), so this is NOT pinned against the reference: it implements the resampler written down below, which
// oracle/convert_oracle.c restates independently and tests/test_gpu_scale.py compares bit for bit.
//
// The resampler ("bilinear" in swscale's sense: a triangle kernel that widens when shrinking), all in integers:
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
//    BGRA input is scaled channel by channel (alpha included).
// Tap tables are built on the host once per geometry (scale_build_axis) and shared by all pictures.
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

// one thread = one destination pixel (the tap tables and the few source rows a block touches stay in L1 / L2)
__global__ void __launch_bounds__(256) k_scale_to_bgra(const __grid_constant__ ScaleArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, k = blockIdx.z;
    if (x >= a.dw) return;
    uint32_t out;
    if (a.format == SCALE_BGRA) {
        const uint8_t *p = a.src[0] + (long long)k * a.src_pic_stride[0];
        out = 0;
        for (int c = 0; c < 4; c++)
            out |= (uint32_t)scale_sample(p + c, a.src_linesize[0], a.sw, a.sh, 4, a.fx, a.wx, a.tx, x, a.fy, a.wy, a.ty, y) << (8 * c);
    } else {
        const int Y = scale_sample(a.src[0] + (long long)k * a.src_pic_stride[0], a.src_linesize[0], a.sw, a.sh, 1,
                                   a.fx, a.wx, a.tx, x, a.fy, a.wy, a.ty, y);
        int U, V;
        if (a.format == SCALE_NV12) {
            const uint8_t *uv = a.src[1] + (long long)k * a.src_pic_stride[1];
            U = scale_sample(uv, a.src_linesize[1], a.cw, a.ch, 2, a.cfx, a.cwx, a.ctx_, x, a.cfy, a.cwy, a.cty, y);
            V = scale_sample(uv + 1, a.src_linesize[1], a.cw, a.ch, 2, a.cfx, a.cwx, a.ctx_, x, a.cfy, a.cwy, a.cty, y);
        } else {
            U = scale_sample(a.src[1] + (long long)k * a.src_pic_stride[1], a.src_linesize[1], a.cw, a.ch, 1,
                             a.cfx, a.cwx, a.ctx_, x, a.cfy, a.cwy, a.cty, y);
            V = scale_sample(a.src[2] + (long long)k * a.src_pic_stride[2], a.src_linesize[2], a.cw, a.ch, 1,
                             a.cfx, a.cwx, a.ctx_, x, a.cfy, a.cwy, a.cty, y);
        }
        const int c = 298 * (Y - 16), d = U - 128, e = V - 128;
        const int r = scale_clampi((c + 409 * e + 128) >> 8, 0, 255);
        const int g = scale_clampi((c - 100 * d - 208 * e + 128) >> 8, 0, 255);
        const int b = scale_clampi((c + 516 * d + 128) >> 8, 0, 255);
        out = 0xFF000000u | ((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b;
    }
    uint32_t *drow = reinterpret_cast<uint32_t *>(a.dst + (long long)k * a.dst_pic_stride + (long long)y * a.dst_stride);
    drow[x] = out;
}
#endif

}  // namespace cvs
#endif
