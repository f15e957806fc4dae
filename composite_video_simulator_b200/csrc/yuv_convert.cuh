// yuv_convert.cuh -- BGRA -> planar YUV 4:2:0 / 4:2:2 (BT.601, limited range) on the device.
//
// The step after the field loop of ffmpeg_ntsc: the BGRA picture goes through sws_scale() to the encoder's pixel
// format (ffmpeg_ntsc.cpp:2266-2274; context created at :2118-2131 with SWS_BILINEAR, frame tagged SMPTE170M /
// MPEG range at :2100-2101).  SURVEY section 8f-1.  libswscale is a third-party dependency that is absent here (no
// FFmpeg in this environment), so this is NOT pinned against the reference: the arithmetic below is the
// published 15-bit fixed-point form of the BT.601 limited-range matrix that swscale's C path uses for RGB input
// (coefficients c * 219/255 * 2^15 for luma and c * 224/255 * 2^15 for chroma, rounded to nearest), with
// the chroma sample taken from the mean of the 2x1 (4:2:2) or 2x2 (4:2:0) pixels it covers, rounded to nearest.
// swscale's bilinear chroma scaler and its chroma siting are not reproduced.  oracle/convert_oracle.c restates the
// conversion independently of this file; tests/test_gpu_yuv_convert.py compares the two bit for bit and checks
// both within +-1 of the real-valued BT.601 conversion.
#ifndef CVS_YUV_CONVERT_CUH
#define CVS_YUV_CONVERT_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace cvs {

constexpr int kYuvShift = 15;
struct YuvCoef { int ry, gy, by, ru, gu, bu, rv, gv, bv; };
inline YuvCoef yuv_coef_bt601() {
    YuvCoef c;
    // rounded to NEAREST, also the negative ones (so that the U and V rows each sum to zero: grey stays 128)
    auto q = [](double v) { const double t = v * (double)(1 << kYuvShift); return (int)(t < 0 ? -(long long)(-t + 0.5) : (long long)(t + 0.5)); };
    const double ys = 219.0 / 255.0, cs = 224.0 / 255.0;
    c.ry = q(0.299 * ys); c.gy = q(0.587 * ys); c.by = q(0.114 * ys);
    c.ru = q(-0.169 * cs); c.gu = q(-0.331 * cs); c.bu = q(0.500 * cs);
    c.rv = q(0.500 * cs); c.gv = q(-0.419 * cs); c.bv = q(-0.081 * cs);
    return c;
}

struct YuvArgs {
    const uint8_t *bgra; uint8_t *y, *u, *v;
    long long sp_bgra, sp_y, sp_u, sp_v;     // picture strides (bytes) of a batch
    int stride, ly, lu, lv;                  // row strides (bytes)
    int w, h, n, v420;                       // v420: chroma is subsampled vertically too
    YuvCoef c;
};

// one thread = 8 pixels x (2 rows for 4:2:0 / 1 row for 4:2:2): 32-byte loads, 8-byte luma stores, 4-byte chroma stores
__global__ void __launch_bounds__(256) k_bgra_to_yuv(const __grid_constant__ YuvArgs a) {
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;          // group of 8 pixels
    const int ry = blockIdx.y;                                     // chroma row
    const int k = blockIdx.z;
    const int x0 = gx * 8;
    if (x0 >= a.w) return;
    const int rows = a.v420 ? 2 : 1;
    const int y0 = ry * rows;
    const uint8_t *src = a.bgra + (long long)k * a.sp_bgra;
    int sr[4] = {0, 0, 0, 0}, sg[4] = {0, 0, 0, 0}, sb[4] = {0, 0, 0, 0};
    const bool full = x0 + 8 <= a.w;
    for (int r = 0; r < rows; r++) {
        const int yy = y0 + r < a.h ? y0 + r : a.h - 1;            // (odd heights: the last row stands for the missing one)
        const uint32_t *row = reinterpret_cast<const uint32_t *>(src + (long long)yy * a.stride);
        uint32_t px[8];
        if (full && ((reinterpret_cast<uintptr_t>(row + x0) & 15) == 0)) {
            const uint4 p0 = *reinterpret_cast<const uint4 *>(row + x0), p1 = *reinterpret_cast<const uint4 *>(row + x0 + 4);
            px[0] = p0.x; px[1] = p0.y; px[2] = p0.z; px[3] = p0.w; px[4] = p1.x; px[5] = p1.y; px[6] = p1.z; px[7] = p1.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) px[i] = row[x0 + i < a.w ? x0 + i : a.w - 1];
        }
        uint32_t yw[2] = {0, 0};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int b = px[i] & 0xFF, g = (px[i] >> 8) & 0xFF, rr = (px[i] >> 16) & 0xFF;
            const int yv = (a.c.ry * rr + a.c.gy * g + a.c.by * b + (16 << kYuvShift) + (1 << (kYuvShift - 1))) >> kYuvShift;
            yw[i >> 2] |= (uint32_t)yv << (8 * (i & 3));
            sr[i >> 1] += rr; sg[i >> 1] += g; sb[i >> 1] += b;
        }
        if (y0 + r < a.h) {
            uint8_t *yd = a.y + (long long)k * a.sp_y + (long long)(y0 + r) * a.ly + x0;
            if (full && ((reinterpret_cast<uintptr_t>(yd) & 7) == 0)) *reinterpret_cast<uint2 *>(yd) = make_uint2(yw[0], yw[1]);
            else
                for (int i = 0; i < 8 && x0 + i < a.w; i++) yd[i] = (uint8_t)(yw[i >> 2] >> (8 * (i & 3)));
        }
    }
    // chroma: mean of the 2 x rows covered pixels, rounded to nearest
    const int sh = kYuvShift + (a.v420 ? 2 : 1);
    const int nsum = a.v420 ? 4 : 2;
    uint32_t uw = 0, vw = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int uu = (a.c.ru * sr[j] + a.c.gu * sg[j] + a.c.bu * sb[j] + ((128 * nsum) << kYuvShift) + (1 << (sh - 1))) >> sh;
        const int vv = (a.c.rv * sr[j] + a.c.gv * sg[j] + a.c.bv * sb[j] + ((128 * nsum) << kYuvShift) + (1 << (sh - 1))) >> sh;
        uw |= (uint32_t)uu << (8 * j);
        vw |= (uint32_t)vv << (8 * j);
    }
    const int cw = (a.w + 1) / 2, cx0 = gx * 4;
    uint8_t *ud = a.u + (long long)k * a.sp_u + (long long)ry * a.lu + cx0;
    uint8_t *vd = a.v + (long long)k * a.sp_v + (long long)ry * a.lv + cx0;
    if (cx0 + 4 <= cw && ((reinterpret_cast<uintptr_t>(ud) & 3) == 0) && ((reinterpret_cast<uintptr_t>(vd) & 3) == 0)) {
        *reinterpret_cast<uint32_t *>(ud) = uw;
        *reinterpret_cast<uint32_t *>(vd) = vw;
    } else {
        for (int j = 0; j < 4 && cx0 + j < cw; j++) { ud[j] = (uint8_t)(uw >> (8 * j)); vd[j] = (uint8_t)(vw >> (8 * j)); }
    }
}

}  // namespace cvs
#endif
