// yuv_convert.cuh -- BGRA -> planar YUV 4:2:0 / 4:2:2 on the device, bit for bit what libswscale's C code produces.
//
// The step after the field loop of ffmpeg_ntsc: the BGRA picture goes through sws_scale() to the encoder's pixel
// format (ffmpeg_ntsc.cpp:2266-2274; context created at :2118-2131 with SWS_BILINEAR, same size both sides; frame
// tagged SMPTE170M / MPEG range at :2100-2101).  SURVEY section 8f-1.
//
// PINNED: oracle/convert_oracle.c restates what the library does for that call and is compared byte for byte with
// libswscale 9.1.100 itself (tests/test_swscale_pin.py, library C code = SWS_ACCURATE_RND | SWS_BITEXACT); the kernels
// below are compared with the oracle, with the library (where the GPU box has it) and with its committed outputs
// (tests/test_gpu_yuv_convert.py).  The arithmetic, all integer:
//   luma    y14 = (RY r + GY g + BY b + (16 << 15) + 256) >> 9;   s15 = min(2 y14, 32767);   Y = clip8((s15 + 64) >> 7)
//   chroma  even width: c14 = (RU (r0 + r1) + GU (g0 + g1) + BU (b0 + b1) + (256 << 15) + 512) >> 10 from the two pixels
//                       of a chroma sample;  s15 = min(2 c14, 32767)
//           odd width:  c14 = (RU r + GU g + BU b + (256 << 14) + 256) >> 9 per PIXEL, then the library's horizontal
//                       bilinear bank (14-bit weights, sws_filter.cpp):  s15 = min((sum c14 w) >> 13, 32767)
//           4:2:2:      C = clip8((s15 + 64) >> 7)
//           4:2:0:      the vertical bank over 2:1 (12-bit weights: 512 1536 1536 512 on rows 2cy-1 .. 2cy+2, folded at
//                       the picture's edges):  C = clip8(((64 << 12) + sum s15 w) >> 19)
//   matrix  R/G/B -> Y: (int)(c 219/255 2^15 + .5), -> U/V: c 224/255, negative ones negated after rounding (BT.601).
#ifndef CVS_YUV_CONVERT_CUH
#define CVS_YUV_CONVERT_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace cvs {

constexpr int kYuvShift = 15;
constexpr int kYuvMaxTaps = 8;
struct YuvCoef { int ry, gy, by, ru, gu, bu, rv, gv, bv; };
inline YuvCoef yuv_coef_bt601() {
    YuvCoef c;
    auto q = [](double v) { return (int)(v * (double)(1 << kYuvShift) + 0.5); };
    c.ry = q(0.299 * 219 / 255); c.gy = q(0.587 * 219 / 255); c.by = q(0.114 * 219 / 255);
    c.ru = -q(0.169 * 224 / 255); c.gu = -q(0.331 * 224 / 255); c.bu = q(0.500 * 224 / 255);
    c.rv = q(0.500 * 224 / 255); c.gv = -q(0.419 * 224 / 255); c.bv = -q(0.081 * 224 / 255);
    return c;
}

struct YuvArgs {
    const uint8_t *bgra; uint8_t *y, *u, *v;
    long long sp_bgra, sp_y, sp_u, sp_v;     // picture strides (bytes) of a batch
    int stride, ly, lu, lv;                  // row strides (bytes)
    int w, h, n, v420;                       // v420: chroma is subsampled vertically too
    const int32_t *vpos, *vcoef;             // vertical chroma bank: crows x vtaps
    const int32_t *hpos, *hcoef;             // horizontal chroma bank of odd widths: cw x htaps
    int vtaps, htaps;
    YuvCoef c;
};

__device__ __forceinline__ int yuv_clip8(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ int yuv_luma(const YuvCoef &c, uint32_t px) {
    const int b = px & 0xFF, g = (px >> 8) & 0xFF, r = (px >> 16) & 0xFF;
    const int y14 = (c.ry * r + c.gy * g + c.by * b + (16 << kYuvShift) + 256) >> 9;
    return yuv_clip8((min(2 * y14, 32767) + 64) >> 7);
}

// Even widths, any alignment, any bank (the general form; the two streaming kernels below take the common cases).
// One thread = 8 pixels x one chroma row: the rows its vertical taps cover (4 of them for 4:2:0, of which two are also
// its luma rows; 1 for 4:2:2) as 32-byte loads, 8-byte luma stores, 4-byte chroma stores.  Neighbouring chroma rows
// share source rows; the second read comes from L2.
__global__ void __launch_bounds__(256) k_bgra_to_yuv(const __grid_constant__ YuvArgs a) {
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;          // group of 8 pixels
    const int cy = blockIdx.y;                                     // chroma row
    const int k = blockIdx.z;
    const int x0 = gx * 8;
    if (x0 >= a.w) return;
    const int l0 = a.v420 ? 2 * cy : cy, l1 = a.v420 ? min(2 * cy + 1, a.h - 1) : cy;      // luma rows of this thread
    const int p = a.vpos[cy], nt = a.vtaps;
    const int32_t *vc = a.vcoef + (size_t)cy * nt;
    const uint8_t *src = a.bgra + (long long)k * a.sp_bgra;
    const bool full = x0 + 8 <= a.w;
    int au[4], av[4];
#pragma unroll
    for (int j = 0; j < 4; j++) au[j] = av[j] = nt > 1 ? (64 << 12) : 64;
    const int lo = min(p, l0), hi = max(p + nt - 1, l1);
    for (int r = lo; r <= hi; r++) {
        const int t = r - p;
        const int wgt = (t >= 0 && t < nt) ? vc[t] : 0;
        const bool luma = r == l0 || r == l1;
        if (wgt == 0 && !luma) continue;
        const uint32_t *row = reinterpret_cast<const uint32_t *>(src + (long long)r * a.stride);
        uint32_t px[8];
        if (full && ((reinterpret_cast<uintptr_t>(row + x0) & 15) == 0)) {
            const uint4 p0 = *reinterpret_cast<const uint4 *>(row + x0), p1 = *reinterpret_cast<const uint4 *>(row + x0 + 4);
            px[0] = p0.x; px[1] = p0.y; px[2] = p0.z; px[3] = p0.w; px[4] = p1.x; px[5] = p1.y; px[6] = p1.z; px[7] = p1.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) px[i] = x0 + i < a.w ? row[x0 + i] : 0u;
        }
        if (luma) {
            uint32_t yw[2] = {0, 0};
#pragma unroll
            for (int i = 0; i < 8; i++) yw[i >> 2] |= (uint32_t)yuv_luma(a.c, px[i]) << (8 * (i & 3));
            uint8_t *yd = a.y + (long long)k * a.sp_y + (long long)r * a.ly + x0;
            if (full && ((reinterpret_cast<uintptr_t>(yd) & 7) == 0)) *reinterpret_cast<uint2 *>(yd) = make_uint2(yw[0], yw[1]);
            else
                for (int i = 0; i < 8 && x0 + i < a.w; i++) yd[i] = (uint8_t)(yw[i >> 2] >> (8 * (i & 3)));
        }
        if (wgt != 0) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t q0 = px[2 * j], q1 = px[2 * j + 1];
                const int sb = (q0 & 0xFF) + (q1 & 0xFF), sg = ((q0 >> 8) & 0xFF) + ((q1 >> 8) & 0xFF), sr = ((q0 >> 16) & 0xFF) + ((q1 >> 16) & 0xFF);
                const int u14 = (a.c.ru * sr + a.c.gu * sg + a.c.bu * sb + (256 << kYuvShift) + 512) >> 10;
                const int v14 = (a.c.rv * sr + a.c.gv * sg + a.c.bv * sb + (256 << kYuvShift) + 512) >> 10;
                const int u15 = min(2 * u14, 32767), v15 = min(2 * v14, 32767);
                if (nt > 1) { au[j] += u15 * wgt; av[j] += v15 * wgt; }
                else { au[j] += u15; av[j] += v15; }
            }
        }
    }
    const int sh = nt > 1 ? 19 : 7;
    uint32_t uw = 0, vw = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uw |= (uint32_t)yuv_clip8(au[j] >> sh) << (8 * j);
        vw |= (uint32_t)yuv_clip8(av[j] >> sh) << (8 * j);
    }
    const int cw = a.w / 2, cx0 = gx * 4;
    uint8_t *ud = a.u + (long long)k * a.sp_u + (long long)cy * a.lu + cx0;
    uint8_t *vd = a.v + (long long)k * a.sp_v + (long long)cy * a.lv + cx0;
    if (cx0 + 4 <= cw && ((reinterpret_cast<uintptr_t>(ud) & 3) == 0) && ((reinterpret_cast<uintptr_t>(vd) & 3) == 0)) {
        *reinterpret_cast<uint32_t *>(ud) = uw;
        *reinterpret_cast<uint32_t *>(vd) = vw;
    } else {
        for (int j = 0; j < 4 && cx0 + j < cw; j++) { ud[j] = (uint8_t)(uw >> (8 * j)); vd[j] = (uint8_t)(vw >> (8 * j)); }
    }
}

// ---- the two common cases as streaming kernels ----------------------------------------------------------------------
// Both are bound by the integer ALU pipe (ncu, profiles/ncu_r2_convert.md), so they drop what the library's arithmetic
// cannot reach with this matrix: with 0 <= r, g, b <= 255 the 14-bit luma lies in [1024, 15040] and the 14-bit chroma of
// a pixel pair in [1024, 15360] (yuv_bounds_ok checks exactly that on the host, from the coefficients), so
// min(2 s14, 32767) never clips, the bytes (s15 + 64) >> 7 = (s14 + 32) >> 6 stay inside [16, 240], and so does a
// weighted mean of such samples with non-negative weights.  Same bytes, no min / max.
inline bool yuv_bounds_ok(const YuvCoef &c) {
    auto lo = [](int a, int b, int d) { return (a < 0 ? a : 0) + (b < 0 ? b : 0) + (d < 0 ? d : 0); };
    auto hi = [](int a, int b, int d) { return (a > 0 ? a : 0) + (b > 0 ? b : 0) + (d > 0 ? d : 0); };
    const long long y_lo = ((long long)lo(c.ry, c.gy, c.by) * 255 + (16 << kYuvShift) + 256) >> 9, y_hi = ((long long)hi(c.ry, c.gy, c.by) * 255 + (16 << kYuvShift) + 256) >> 9;
    const long long u_lo = ((long long)lo(c.ru, c.gu, c.bu) * 510 + (256 << kYuvShift) + 512) >> 10, u_hi = ((long long)hi(c.ru, c.gu, c.bu) * 510 + (256 << kYuvShift) + 512) >> 10;
    const long long v_lo = ((long long)lo(c.rv, c.gv, c.bv) * 510 + (256 << kYuvShift) + 512) >> 10, v_hi = ((long long)hi(c.rv, c.gv, c.bv) * 510 + (256 << kYuvShift) + 512) >> 10;
    auto fits = [](long long a, long long b) { return a >= 0 && 2 * b <= 32767 && ((2 * b + 64) >> 7) <= 255; };
    return fits(y_lo, y_hi) && fits(u_lo, u_hi) && fits(v_lo, v_hi);
}
// Shared by both: 8 pixels of one source row -> 8 luma bytes and the 15-bit chroma of the 4 samples they carry.
struct Yuv8 { uint32_t y0, y1; int u[4], v[4]; };
__device__ __forceinline__ Yuv8 yuv_row8(const YuvCoef &c, const uint32_t px[8], bool want_luma, bool want_chroma) {
    Yuv8 o;
    o.y0 = o.y1 = 0;
    int b[8], g[8], r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { b[i] = px[i] & 0xFF; g[i] = (px[i] >> 8) & 0xFF; r[i] = (px[i] >> 16) & 0xFF; }
    if (want_luma) {
        int y[8];
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = (((c.ry * r[i] + c.gy * g[i] + c.by * b[i] + (16 << kYuvShift) + 256) >> 9) + 32) >> 6;
        o.y0 = (uint32_t)(y[0] + (y[1] << 8) + (y[2] << 16) + (y[3] << 24));
        o.y1 = (uint32_t)(y[4] + (y[5] << 8) + (y[6] << 16) + (y[7] << 24));
    }
    if (want_chroma) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int sb = b[2 * j] + b[2 * j + 1], sg = g[2 * j] + g[2 * j + 1], sr = r[2 * j] + r[2 * j + 1];
            o.u[j] = 2 * ((c.ru * sr + c.gu * sg + c.bu * sb + (256 << kYuvShift) + 512) >> 10);
            o.v[j] = 2 * ((c.rv * sr + c.gv * sg + c.bv * sb + (256 << kYuvShift) + 512) >> 10);
        }
    }
    return o;
}
__device__ __forceinline__ void yuv_load8(const uint8_t *rowp, int x0, uint32_t px[8]) {
    const uint4 p0 = *reinterpret_cast<const uint4 *>(rowp + 4 * (size_t)x0), p1 = *reinterpret_cast<const uint4 *>(rowp + 4 * (size_t)x0 + 16);
    px[0] = p0.x; px[1] = p0.y; px[2] = p0.z; px[3] = p0.w; px[4] = p1.x; px[5] = p1.y; px[6] = p1.z; px[7] = p1.w;
}

// 4:2:2, widths that are multiples of 8, 16-byte aligned rows: nothing is shared between rows; one thread = 8 pixels.
__global__ void __launch_bounds__(256) k_bgra_to_yuv422_fast(const __grid_constant__ YuvArgs a) {
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y, k = blockIdx.z;
    const int x0 = gx * 8;
    if (x0 >= a.w) return;
    uint32_t px[8];
    yuv_load8(a.bgra + (long long)k * a.sp_bgra + (long long)r * a.stride, x0, px);
    const Yuv8 o = yuv_row8(a.c, px, true, true);
    *reinterpret_cast<uint2 *>(a.y + (long long)k * a.sp_y + (long long)r * a.ly + x0) = make_uint2(o.y0, o.y1);
    uint32_t uw = 0, vw = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uw |= (uint32_t)((o.u[j] + 64) >> 7) << (8 * j);
        vw |= (uint32_t)((o.v[j] + 64) >> 7) << (8 * j);
    }
    *reinterpret_cast<uint32_t *>(a.u + (long long)k * a.sp_u + (long long)r * a.lu + gx * 4) = uw;
    *reinterpret_cast<uint32_t *>(a.v + (long long)k * a.sp_v + (long long)r * a.lv + gx * 4) = vw;
}

// 4:2:0, same conditions.  A block owns 256 pixels x kTileC chroma rows: every source row the tile's vertical taps cover
// (2 kTileC + 2 of them away from the edges) is loaded ONCE, its luma stored and its 15-bit chroma parked in shared
// memory; after a barrier each thread weighs the rows of one chroma row.  Source rows are read 1 + 2 / (2 kTileC)
// times instead of twice, and the chroma matrix runs once per row.  (0.47 of the HBM peak against 0.36 for the general
// kernel, the same with tiles of 8 or of 7 chroma rows: the kernel is bound by the ALU pipe, 81 % busy, not by its passes.)
constexpr int kTileC = 7, kTileRows = 2 * kTileC + 4;    // 7 chroma rows need 16 source rows: two full passes of the block's 8 row slots
__global__ void __launch_bounds__(256, 6) k_bgra_to_yuv420_tiled(const __grid_constant__ YuvArgs a) {
    __shared__ uint4 su[kTileRows][32], sv[kTileRows][32];       // per row and 8-pixel group: u15[4], v15[4]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 groups x 8 rows
    const int gx = blockIdx.x * 32 + tx, x0 = gx * 8, k = blockIdx.z;
    const int crows = (a.h + 1) / 2;
    const int c0 = blockIdx.y * kTileC, c1 = min(c0 + kTileC, crows) - 1;
    const int nt = a.vtaps;
    const int l_lo = 2 * c0, l_hi = min(2 * c1 + 1, a.h - 1);    // luma rows of the tile
    const int lo = min(a.vpos[c0], l_lo), hi = max(a.vpos[c1] + nt - 1, l_hi);   // hi - lo < kTileRows: checked by the host
    const bool live = x0 < a.w;
    const uint8_t *src = a.bgra + (long long)k * a.sp_bgra;
    for (int r = lo + ty; r <= hi; r += 8) {
        if (!live) continue;
        uint32_t px[8];
        yuv_load8(src + (long long)r * a.stride, x0, px);
        const bool luma = r >= l_lo && r <= l_hi;
        const Yuv8 o = yuv_row8(a.c, px, luma, true);
        if (luma) *reinterpret_cast<uint2 *>(a.y + (long long)k * a.sp_y + (long long)r * a.ly + x0) = make_uint2(o.y0, o.y1);
        su[r - lo][tx] = make_uint4((uint32_t)o.u[0], (uint32_t)o.u[1], (uint32_t)o.u[2], (uint32_t)o.u[3]);
        sv[r - lo][tx] = make_uint4((uint32_t)o.v[0], (uint32_t)o.v[1], (uint32_t)o.v[2], (uint32_t)o.v[3]);
    }
    __syncthreads();
    const int cy = c0 + ty;
    if (!live || cy > c1) return;
    const int p = a.vpos[cy] - lo;
    const int32_t *vc = a.vcoef + (size_t)cy * nt;
    int au[4], av[4];
#pragma unroll
    for (int j = 0; j < 4; j++) au[j] = av[j] = 64 << 12;
    for (int t = 0; t < nt; t++) {
        const int wgt = vc[t];
        if (wgt == 0) continue;
        const uint4 u4 = su[p + t][tx], v4 = sv[p + t][tx];
        au[0] += (int)u4.x * wgt; au[1] += (int)u4.y * wgt; au[2] += (int)u4.z * wgt; au[3] += (int)u4.w * wgt;
        av[0] += (int)v4.x * wgt; av[1] += (int)v4.y * wgt; av[2] += (int)v4.z * wgt; av[3] += (int)v4.w * wgt;
    }
    uint32_t uw = 0, vw = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uw |= (uint32_t)(au[j] >> 19) << (8 * j);
        vw |= (uint32_t)(av[j] >> 19) << (8 * j);
    }
    *reinterpret_cast<uint32_t *>(a.u + (long long)k * a.sp_u + (long long)cy * a.lu + gx * 4) = uw;
    *reinterpret_cast<uint32_t *>(a.v + (long long)k * a.sp_v + (long long)cy * a.lv + gx * 4) = vw;
}

// Odd widths: the library keeps chroma at full width on the source side and resamples it horizontally, so every
// chroma sample is a small 2-D filter.  One thread = one chroma sample and the luma pixels under it.  (Encoders do
// not take odd widths; this path exists so that every size the library accepts gives the library's bytes.)
__global__ void __launch_bounds__(256) k_bgra_to_yuv_oddw(const __grid_constant__ YuvArgs a) {
    const int cx = blockIdx.x * blockDim.x + threadIdx.x;
    const int cy = blockIdx.y;
    const int k = blockIdx.z;
    const int cw = (a.w + 1) / 2;
    if (cx >= cw) return;
    const uint8_t *src = a.bgra + (long long)k * a.sp_bgra;
    const int l0 = a.v420 ? 2 * cy : cy, l1 = a.v420 ? min(2 * cy + 1, a.h - 1) : cy;
    for (int r = l0; r <= l1; r++) {
        const uint32_t *row = reinterpret_cast<const uint32_t *>(src + (long long)r * a.stride);
        uint8_t *yd = a.y + (long long)k * a.sp_y + (long long)r * a.ly;
        for (int x = 2 * cx; x < min(2 * cx + 2, a.w); x++) yd[x] = (uint8_t)yuv_luma(a.c, row[x]);
    }
    const int p = a.vpos[cy], nt = a.vtaps, hp = a.hpos[cx], ht = a.htaps;
    const int32_t *vc = a.vcoef + (size_t)cy * nt, *hc = a.hcoef + (size_t)cx * ht;
    int au = nt > 1 ? (64 << 12) : 64, av = au;
    for (int t = 0; t < nt; t++) {
        const int wgt = vc[t];
        if (wgt == 0) continue;
        const uint32_t *row = reinterpret_cast<const uint32_t *>(src + (long long)(p + t) * a.stride);
        int su = 0, sv = 0;
        for (int j = 0; j < ht; j++) {
            const int hw = hc[j];
            if (hw == 0) continue;
            const uint32_t px = row[hp + j];
            const int b = px & 0xFF, g = (px >> 8) & 0xFF, r = (px >> 16) & 0xFF;
            su += ((a.c.ru * r + a.c.gu * g + a.c.bu * b + (256 << (kYuvShift - 1)) + 256) >> 9) * hw;
            sv += ((a.c.rv * r + a.c.gv * g + a.c.bv * b + (256 << (kYuvShift - 1)) + 256) >> 9) * hw;
        }
        const int u15 = min(su >> 13, 32767), v15 = min(sv >> 13, 32767);
        if (nt > 1) { au += u15 * wgt; av += v15 * wgt; }
        else { au += u15; av += v15; }
    }
    const int sh = nt > 1 ? 19 : 7;
    a.u[(long long)k * a.sp_u + (long long)cy * a.lu + cx] = (uint8_t)yuv_clip8(au >> sh);
    a.v[(long long)k * a.sp_v + (long long)cy * a.lv + cx] = (uint8_t)yuv_clip8(av >> sh);
}

}  // namespace cvs
#endif
