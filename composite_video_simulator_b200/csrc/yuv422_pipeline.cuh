// yuv422_pipeline.cuh -- the reference's 4:2:2 scanline path, composite_video_process()
// (ffmpeg_to_composite.cpp:629-952), as a per-lane streaming pipeline.
//
// One lane owns one field scanline, a group of 31 consecutive rows plus a halo lane (the row above, for the
// vertical chroma blend) advances through x in lock-step, 8 luma pixels (4 chroma samples) per step.  This path
// quantises to 8 bits after EVERY stage (clampu8, :335-342) and works in place on the picture, so the stages
// communicate through small per-lane byte rings in shared memory that play the role of the reference's in-place
// planes -- including its quirks (the last `delay` samples of a delayed filter keep their unfiltered values because
// nobody overwrites them).  A stage with look-ahead or write delay simply runs on an older block; block lags are
// launch constants (lags_of).
//
// The stages of a row are shared out to FOUR ROLES, one warp each (see lags_of and yuv422_kernels.cuh): a warp then
// holds a quarter of the filter state (~95 registers instead of 210), five groups of four warps fit on an SM, and the roles of a
// group meet at a barrier after every step.
//
// Arithmetic: the reference's filters are IEEE doubles and their results are truncated to 8 bits, and a float
// evaluation does not stay within one LSB here (profiles/yuv422_fp32_experiment_r2.txt): every filter is evaluated
// in double, operation for operation as the reference does, and the output is BIT-EXACT.
//
// Two variants of every stage:
//   Pipe422<true>   "general": any width, any switch, partial blocks, line start / end; byte-wise ring writes.
//                   Used by the general kernel k_yuv422 for everything the fast kernel does not take.
//   Fast422         whole blocks, the common switches; line start and end are handled inside the same code (block 0
//                   of demod_w, the head-switch fill, roles skipping steps whose block does not exist).  Rolled filter
//                   loops, saturating packs, conversion-pipe conversions, word-wise delayed ring stores, the product
//                   shared between the poles of a cascade.  One code path per role keeps what an SM executes inside
//                   its 32 KB instruction cache, which decides everything else (scripts/probes/ifetch_probe.cu).
//
// Host/device portable like lane_pipeline.cuh: tests/emu422_harness.cpp runs both variants on the CPU against the
// oracle, with the roles of a step in either order (they must not depend on each other within a step).
//
// Stages (block lags: lags_of):
//   G0   load Y,U,V block into the rings (and the two luma bytes past the row, :496)
//   G1   input chroma lowpass, written at c-2 (U) / c-4 (V)                                  (:353-393)
//   G2   modulate chroma onto luma, pre-emphasis, luma noise, head-switch delay              (:434-477,:636-733)
//   G3a  Y/C separation + demodulation, chroma noise, phase noise; VHS luma lowpass + boost  (:480-553,:738-783,:810-828)
//   G3b  VHS chroma lowpass written at c-cd                                                  (:830-851)
//   G4   VHS luma sharpen; vertical chroma blend (lane above via shuffle), chroma sharpen    (:858-925)
//   G5   VHS re-modulation and second demodulation                                           (:927-930)
//   GE   chroma dropout, -yc-recomb rounds (one block of lag each)                           (:932-946)
//   GO   output chroma lowpass (full or lite)                                                (:948-951)
//   ST   store
#ifndef CVS_YUV422_PIPELINE_CUH
#define CVS_YUV422_PIPELINE_CUH

#include "lane_pipeline.cuh"

namespace cvs422 {

using cvs::LaneRng;
using cvs::draw_mod;
using cvs::noise_step;
using cvs::kRngBase;
using cvs::kRngSlots;
using cvs::umulhi32;

constexpr int kB = 8;                    // luma pixels per block
constexpr int kBC = 4;                   // chroma samples per block
constexpr int kRingY = 128;              // per-lane luma ring (bytes, power of two)
constexpr int kRingC = 64;               // per-lane chroma rings
constexpr int kMaxRecombine = 4;         // -yc-recomb rounds the rings have room for
constexpr int kRingA = 64;               // per-lane ring of the composite signal for the in-ring head-switch delay
constexpr int kHsMaxDelay = kRingA - 2 * kB;   // largest head-switch delay that ring can express
constexpr int kWarm = 64;                // noise warm-up length (samples), see cvs::warm_luma
constexpr int kPhaseMapMax = 15;         // largest -chroma-phase-noise whose rotation is tabulated (31 states x 512 bytes)

enum : uint32_t {
    G_IN_LP = 1u << 0,        // composite_in_chroma_lowpass
    G_OUT_FULL = 1u << 1,     // composite_out_chroma_lowpass
    G_OUT_LITE = 1u << 2,     // !full && composite_out_chroma_lowpass_lite
    G_PREEMPH = 1u << 3,
    G_NOCOLOR = 1u << 4,      // nocolor_subcarrier
    G_NOCOLOR_YC = 1u << 5,   // nocolor_subcarrier_after_yc_sep
    G_VHS = 1u << 6,
    G_VBLEND = 1u << 7,       // vhs_chroma_vert_blend && output_ntsc
    G_SVIDEO = 1u << 8,
    G_PHASE = 1u << 9,        // video_chroma_phase_noise != 0
    G_GENERAL = 1u << 10,     // a rarely used switch is on: the general kernel
    G_PHASE_MAP = 1u << 11,   // the phase-noise rotation of a byte is tabulated behind phase_lut (pnoise <= kPhaseMapMax)
};

// per-row flags in the host side table (same packing as cvs::rowinfo_pack)
enum : uint32_t {
    RG_DROPOUT = 1u << 0,
    RG_HEADSW_PRE = 1u << 1,  // row is rotated by the head switch; its composite luma comes from the pre-pass scratch row
    RG_HEADSW = 1u << 2,      // row is delayed by hs_delay pixels with fill 16 (:697-731): done in the ring
};

struct K422 {
    double a_in[2], a_inhp[2];            // input lowpass poles / its boost highpass, [0] = U, [1] = V   (:377-383)
    double a_out[2], a_outhp[2];          // output lowpass (full: as input; lite: rate/8, no boost)     (:412-416)
    double a_pre, preemph;                // composite pre-emphasis                                       (:642-647)
    double a_luma;                        // VHS luma lowpass and boost                                   (:817-821)
    double a_lsharp, sharpen;             // VHS luma sharpen (2x luma cut)                               (:894-899)
    double a_ch;                          // VHS chroma lowpass                                           (:837-841)
    double a_csharp, sharpen_c;           // VHS chroma sharpen (2x chroma cut)                           (:910-921)
    const double *phase_lut;              // [2*pnoise+1][2] = {cos, sin}(state*pi/100)                   (:765,:772-773)
                                          // G_PHASE_MAP: followed by [2*pnoise+1][2][256] bytes, the rotated U / V byte of
                                          // every input byte and state (the rotation's input is an 8-bit sample)
    uint32_t flags;
    int32_t d_in[2], d_out[2];            // write delays of the input / output lowpass (samples)
    int32_t cd;                           // VHS chroma delay 4 / 5 / 6                                   (:793-803)
    int32_t amp, amp_back;                // subcarrier_amplitude, _back
    int32_t vnoise, cnoise, pnoise;
    uint32_t vmagic, vshift, cmagic, cshift;
    int32_t ntsc, phase_shift, phase_offset;
    int32_t recombine;                    // video_yc_recombine (<= kMaxRecombine)
    int32_t lagV;                         // blocks between G3 and G4: 1 (cd = 4) or 2
    int32_t w, h, cw;
};

// Block lags of a configuration (blocks behind the load front: a stage works on block s - lag at step s).
// A row is worked on by FOUR warps at once, one per ROLE (a group of consecutive stages); the roles of a row meet at
// a barrier after every step, so a role reads what the previous role wrote in an EARLIER step: one extra block of
// lag at every role boundary.  (Five roles -- role 2 split after the luma sharpen -- were measured too: 3.65 ms
// against 3.53 ms per launch; more code for the same instruction cache, profiles/ab_variants_r2.txt.)
//   role 0: G0 G1 G2      load, input chroma lowpass, first modulation (luma noise, head switch)
//   role 1: G3a           Y/C separation, chroma / phase noise, VHS luma lowpass + boost
//   role 2: G3b G4        VHS chroma lowpass, VHS luma sharpen, vertical chroma blend, chroma sharpen
//   role 3: G5 GE GO ST   re-modulation, second demodulation, dropout, -yc-recomb, output chroma lowpass, store
constexpr int kRoles = 4;
struct Lags {
    int bM, bD, bC, bV, bR, bD2, bE, bF, bS;   // as offsets: block = s - lag
};
CVS_HD Lags lags_of(const K422 &K) {
    Lags L;
    const bool vhs = (K.flags & G_VHS) != 0, sv = (K.flags & G_SVIDEO) != 0;
    L.bM = 1;                             // the input lowpass writes up to 4 samples back
    L.bD = L.bM + 2;                      // demodulation reads two bytes of the next block; role boundary
    L.bC = L.bD + 1;                      // role boundary
    L.bV = L.bC + K.lagV;                 // the VHS chroma lowpass writes 4..6 samples back
    L.bR = L.bV + 1;                      // role boundary
    L.bD2 = L.bR + 1;                     // as bD
    L.bE = vhs ? (sv ? L.bR : L.bD2) : L.bD + 1;
    L.bF = L.bE + K.recombine;
    L.bS = L.bF + 1;
    return L;
}
CVS_HD int line_steps(const K422 &K) {
    const int nb = (K.w + 2 + kB - 1) / kB;          // blocks that carry data (incl. the two bytes past the row)
    return nb + lags_of(K).bS + 1;
}

// ---- arithmetic ------------------------------------------------------------------------------------
CVS_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
CVS_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
CVS_HD double dsub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
// exact (double)v for 0 <= v < 2^31 without a conversion instruction: splice v into the mantissa of 2^52
CVS_HD double u2d(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(__hiloint2double(0x43300000, (int)v), 4503599627370496.0);
#else
    return (double)v;
#endif
}
CVS_HD double i2d(int v) {               // small signed values (|v| < 2^31)
#if defined(__CUDA_ARCH__)
    return __dsub_rn(__hiloint2double(0x43300000, (int)((uint32_t)v ^ 0x80000000u)), 4503601774854144.0);   // 2^52 + 2^31
#else
    return (double)v;
#endif
}
CVS_HD int clamp8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
// clampu8((int)s): truncation and floor agree once the result is clamped at 0, and floor is one
// round-down add of 1.5 * 2^52 whose low word is the integer (|s| < 2^31)
CVS_HD int q8(double s) {
#if defined(__CUDA_ARCH__)
    return clamp8(__double2loint(__dadd_rd(s, 6755399441055744.0)));
#else
    return clamp8((int)s);
#endif
}
// LowpassFilter::lowpass (:114-118): prev = s*alpha + (prev - prev*alpha), three roundings
CVS_HD double pole(double &p, double s, double a) {
    const double t = dmul(s, a);
    const double u = dsub(p, dmul(p, a));
    p = dadd(t, u);
    return p;
}
// The same filter with the product shared (the fast kernel).  A pole's next state is rn(t + u) with t = rn(s a) and
// u = rn(p - rn(p a)); u depends on the OLD state only, so it is what a lane carries, and in a cascade of poles with
// one alpha the product rn(p' a) that forms pole k's next u is, operand for operand, also the t of pole k + 1:
// 1 + 3 K operations for K poles instead of 4 K, every one of them the reference's.
CVS_HD double carry_of(double p, double a) { return dsub(p, dmul(p, a)); }
template <int KP>
CVS_HD double cascade(double *u, double s, double a) {        // u[k] = carry of pole k; returns the last pole's output
    double t = dmul(s, a), p = 0;
    CVS_UNROLL
    for (int k = 0; k < KP; k++) {
        p = dadd(t, u[k]);
        t = dmul(p, a);
        u[k] = dsub(p, t);
    }
    return p;
}
CVS_HD int div_trunc(int v, int den, uint32_t magic, uint32_t shift) {     // C '/' for |v| < 2^31, den > 0
    const uint32_t a = (uint32_t)(v < 0 ? -v : v);
    const uint32_t q = umulhi32(a, magic) >> shift;
    (void)den;
    return v < 0 ? -(int)q : (int)q;
}
CVS_HD int div50(int v) {                // exact for every 32-bit magnitude
    const uint32_t a = (uint32_t)(v < 0 ? -v : v);
    const uint32_t q = umulhi32(a, 0x51EB851Fu) >> 4;
    return v < 0 ? -(int)q : (int)q;
}

// ---- interior-variant helpers (round 2) ------------------------------------------------------------------
// The interior loop is bound by the FP64 pipe (one warp instruction per two cycles per scheduler) and by the
// instruction cache (32 KB L1.5 per SM), so it (1) keeps the 8-bit <-> double conversions off the FP64 pipe,
// on the otherwise idle conversion pipe (I2F.F64 / F2I.F64.TRUNC), (2) clamps and packs with the saturating
// pack (I2IP.U8.S32.SAT: four values -> one word in two instructions), (3) runs its filters as ROLLED loops
// of CVS422_CU chroma samples / CVS422_LU luma pixels per iteration, and (4) writes delayed filter outputs as
// whole words (the bytes that belong to the next word wait in a carry register).
#ifndef CVS422_XU_CONV
#define CVS422_XU_CONV 1
#endif
#ifndef CVS422_CU
#define CVS422_CU 2                      // chroma samples per iteration of a rolled filter loop: 1, 2 or 4
#endif
#ifndef CVS422_CU_MID
#define CVS422_CU_MID CVS422_CU          // the same for the VHS chroma lowpass and the chroma sharpen (role 2, the longest role)
#endif
#ifndef CVS422_LU
#define CVS422_LU 8                      // luma pixels per iteration: 1, 2, 4 or 8
#endif
// Which role runs the VHS luma sharpen (G4a): 2 (with the chroma stages of G3b / G4: FP64 instructions per role and step
// 123 / 123 / 276 / 123) or 0 (with the load and the input lowpass: 221 / 123 / 178 / 123).  Same block, same step: the
// stage reads what role 1 wrote one step earlier and nobody else touches that luma block in between, whichever role runs
// it; both are bit-exact.  Measured (profiles/ab_variants_r2.txt section 8): 89.9 k fields/s with 2, 83.1 k with 0 -- the
// better FP64 balance loses, so 2 stays.
#ifndef CVS422_LSHARP_ROLE
#define CVS422_LSHARP_ROLE 2
#endif
#if defined(__CUDA_ARCH__)
#define CVS_ROLLED _Pragma("unroll 1")
#else
#define CVS_ROLLED
#endif

CVS_HD double fu2d(uint32_t v) {         // exact (double)v
#if defined(__CUDA_ARCH__) && CVS422_XU_CONV
    return __uint2double_rn(v);
#else
    return u2d(v);
#endif
}
// an integer whose clamp to 0..255 is clampu8((int)s): truncation or floor, they differ only below zero
CVS_HD int fq(double s) {
#if defined(__CUDA_ARCH__)
#if CVS422_XU_CONV
    return __double2int_rz(s);
#else
    return __double2loint(__dadd_rd(s, 6755399441055744.0));
#endif
#else
    return (int)s;
#endif
}
// (upper << 16) | clamp8(hi) << 8 | clamp8(lo)
CVS_HD uint32_t sat_pack2(int hi, int lo, uint32_t upper) {
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(upper));
    return d;
#else
    return (upper << 16) | ((uint32_t)clamp8(hi) << 8) | (uint32_t)clamp8(lo);
#endif
}
CVS_HD uint32_t sat4(int a, int b, int c, int d) { return sat_pack2(b, a, sat_pack2(d, c, 0u)); }   // a = byte 0
CVS_HD int sat1(int v) { return (int)sat_pack2(0, v, 0u); }                                        // clamp8 in one instruction
CVS_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) {             // low word of (hi:lo) >> sh, 0 < sh < 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, (uint32_t)sh);
#else
    return (lo >> sh) | (hi << (32 - sh));
#endif
}
// __byte_perm for selectors whose nibbles are 0..7
CVS_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t ab = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) r |= (uint32_t)((ab >> (8 * ((sel >> (4 * k)) & 7))) & 0xFFu) << (8 * k);
    return r;
#endif
}
// per-byte (a + b + 1) >> 1
CVS_HD uint32_t avg4_up(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) & 0xFEFEFEFEu) >> 1); }

// ---- per-row constants -------------------------------------------------------------------------------
struct Row422 {
    int xi;                // subcarrier phase index of the line (:449-459)
    int row;               // field row index
    uint32_t rflags;
    int hs_delay;          // RG_HEADSW: Y[x] = Y[x - hs_delay], 16 for x < hs_delay
    int mU[4], mV[4];      // carrier taps for (x & 3): Umult / Vmult rotated by xi (:438-439)
    int fl[4];             // 0xFF where the demodulator flips the sign of the chroma sample (x & 3), else 0 (:527-530)
    uint32_t flx;          // ~(fl[0..3] as bytes): chroma word ^ flx = 255 - (flipped) chroma for four pixels at once
    uint32_t selU, selV;   // byte selectors that pick the U / V samples out of eight demodulated pixels
    uint32_t selM;         // byte selector: the chroma sample (U or V) that rides on pixels 0..3 of a block; + 0x2222: 4..7
    int mX[4], mC[4];      // modulation of pixel (x & 3) without a multiply: y + (t ^ mX) + mC = y +- (t - 128)
    double cosp, sinp;     // phase noise rotation of this row
    const uint8_t *pmap;   // fast kernel: the row's 512-byte rotation table (U then V), or null
};

CVS_HD int line_phase(const K422 &K, unsigned long long fieldno, unsigned y) {
    if (!K.ntsc) return (int)((fieldno + y) & 3);
    if (K.phase_shift == 90) return (int)((fieldno + (unsigned long long)(long long)K.phase_offset + (y >> 1)) & 3);
    if (K.phase_shift == 180) return (int)((((fieldno + y) & 2) + (unsigned long long)(long long)K.phase_offset) & 3);
    if (K.phase_shift == 270) return (int)((fieldno + (unsigned long long)(long long)K.phase_offset - (y >> 1)) & 3);
    return 0;
}

CVS_HD void row_setup(const K422 &K, unsigned field, unsigned long long fieldno, int row, uint32_t rowinfo, Row422 &rc) {
    rc.row = row;
    rc.rflags = (rowinfo >> 16) & 0xFFu;
    rc.hs_delay = (rc.rflags & RG_HEADSW) ? (int)(rowinfo >> 24) : 0;
    rc.xi = line_phase(K, fieldno, field + 2u * (unsigned)row);
    CVS_UNROLL
    for (int j = 0; j < 4; j++) {
        const int ph = (rc.xi + j) & 3;
        rc.mU[j] = (ph == 0) ? 1 : ((ph == 2) ? -1 : 0);
        rc.mV[j] = (ph == 1) ? 1 : ((ph == 3) ? -1 : 0);
        rc.fl[j] = (((j + rc.xi + 2) & 3) < 2) ? 0xFF : 0;
    }
    rc.flx = ~((uint32_t)rc.fl[0] | ((uint32_t)rc.fl[1] << 8) | ((uint32_t)rc.fl[2] << 16) | ((uint32_t)rc.fl[3] << 24));
    rc.selM = 0;
    CVS_UNROLL
    for (int j = 0; j < 4; j++) {
        const int ph = (rc.xi + j) & 3;                       // 0: +U, 1: +V, 2: -U, 3: -V
        rc.selM |= (uint32_t)(((ph & 1) ? 4 : 0) + (j >> 1)) << (4 * j);
        rc.mX[j] = (ph & 2) ? -1 : 0;
        rc.mC[j] = (ph & 2) ? 129 : -128;
    }
    rc.selU = (rc.xi & 1) ? 0x7531u : 0x6420u;
    rc.selV = (rc.xi & 1) ? 0x6420u : 0x7531u;
    rc.pmap = nullptr;
    if (K.flags & G_PHASE) {
        const int st = (int)(int16_t)(rowinfo & 0xFFFFu);
        rc.cosp = K.phase_lut[2 * (st + K.pnoise)];
        rc.sinp = K.phase_lut[2 * (st + K.pnoise) + 1];
    } else {
        rc.cosp = 1;
        rc.sinp = 0;
    }
}

// the byte tables behind phase_lut (G_PHASE_MAP) and a row's table in a copy of them
CVS_HD const uint8_t *phase_maps(const K422 &K) { return reinterpret_cast<const uint8_t *>(K.phase_lut + 2 * (2 * K.pnoise + 1)); }
CVS_HD size_t phase_maps_bytes(const K422 &K) { return (K.flags & G_PHASE_MAP) ? (size_t)(2 * K.pnoise + 1) * 512 : 0; }
CVS_HD const uint8_t *row_phase_map(const K422 &K, const uint8_t *maps, uint32_t rowinfo) {
    return maps + (size_t)((int)(int16_t)(rowinfo & 0xFFFFu) + K.pnoise) * 512;
}

// ---- per-lane state -----------------------------------------------------------------------------------
struct Demod {               // the three luma samples before x+2 of the 4-tap box (:487-499)
    int o1, o2, o3;
    CVS_HD void reset(int y0, int y1) { o1 = 16; o2 = y0; o3 = y1; }
};

struct Lane422 {
    double inU[4], inV[4];   // [0] boost highpass, [1..3] lowpass
    double pre;
    double lum[4];           // three lowpass poles + boost
    double lsh[3];
    double chU[3], chV[3];
    double csU[3], csV[3];
    double outU[4], outV[4];
    Demod dm1, dm2;
    int nY, nU, nV;
    LaneRng rngL, rngC;
    uint32_t cIn[2], cCh[2], cOut[2];  // interior variant: bytes of a delayed filter that wait for their word (U, V)
    uint8_t *ry, *ru, *rv, *rya;     // lane-private rings (kRingY / kRingC / kRingC / kRingA bytes), 4-byte aligned
    int32_t *rcomb;                  // 3 ints per -yc-recomb round (lane-private, 3 * kMaxRecombine ints)

    CVS_HD void reset() {
        for (int i = 0; i < 4; i++) { inU[i] = inV[i] = 128; outU[i] = outV[i] = 128; lum[i] = 16; }
        for (int i = 0; i < 3; i++) { lsh[i] = 16; chU[i] = chV[i] = 128; csU[i] = csV[i] = 128; }
        pre = 16;
        nY = nU = nV = 0;
        cIn[0] = cIn[1] = cCh[0] = cCh[1] = cOut[0] = cOut[1] = 0;
        dm1.reset(16, 16);
        dm2.reset(16, 16);
    }
};

CVS_HD uint32_t pack4(int a, int b, int c, int d) {           // four values 0..255 -> one word
#if defined(__CUDA_ARCH__)
    return __byte_perm(__byte_perm((uint32_t)a, (uint32_t)b, 0x0040), __byte_perm((uint32_t)c, (uint32_t)d, 0x0040), 0x5410);
#else
    return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24);
#endif
}
CVS_HD int byte_of(uint32_t w, int k) {                       // k is a compile-time constant at every call site
#if defined(__CUDA_ARCH__)
    return (int)__byte_perm(w, 0u, 0x4440u + (uint32_t)k);
#else
    return (int)((w >> (8 * k)) & 0xFFu);
#endif
}

// Ring access.  Blocks are 8 luma bytes / 4 chroma bytes and the rings hold 16 of them, so a block never
// wraps: whole blocks move as 32-bit words (the rings are 4-byte aligned), only delayed writes are bytes.
CVS_HD uint8_t *yblk(uint8_t *ring, int b) { return ring + ((b & (kRingY / kB - 1)) * kB); }
CVS_HD uint8_t *cblk(uint8_t *ring, int b) { return ring + ((b & (kRingC / kBC - 1)) * kBC); }
CVS_HD uint32_t ldw(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
    return *reinterpret_cast<const uint32_t *>(p);
#else
    uint32_t v;
    __builtin_memcpy(&v, p, 4);
    return v;
#endif
}
CVS_HD void stw(uint8_t *p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint32_t *>(p) = v;
#else
    __builtin_memcpy(p, &v, 4);
#endif
}

// block counts of a row
struct Geo {
    int nb;      // blocks with pixels
    int nbl;     // blocks the ring must take in (the two bytes past the row included)
};
CVS_HD Geo geo_of(const K422 &K) {
    Geo g;
    g.nb = (K.w + kB - 1) / kB;
    g.nbl = (K.w + 2 + kB - 1) / kB;
    return g;
}

// the division constants of the two demodulator gains (amp_back for the first, amp for the others)
struct DivPair {
    uint32_t back_magic, back_shift, amp_magic, amp_shift;
};

// What one step exchanges with its caller: block s of the source row comes in, the pre-blend chroma of
// block bV goes to the lane below between front() and back(), block bS comes out.
struct StepIO {
    uint32_t y0, y1, u, v;        // in: block s (bytes beyond the data are don't-care); out: block bS
};

template <bool EDGE>
struct Pipe422 {
    static CVS_HD bool in(int x, int lim) { return !EDGE || x < lim; }

    // a delayed three-pole chroma lowpass on the four samples of block b (packed in wv): P[c - d] = lp(P[c]),
    // optionally preceded by the boost s += highpass(s)                                    (:353-431, :830-851)
    static CVS_HD void chroma_lp4(const K422 &K, uint8_t *ring, int b, uint32_t wv, double *hp, double lp[3],
                                  double a_lp, double a_hp, int d) {
        const int c0 = b * kBC, t = c0 - d;
        CVS_UNROLL
        for (int k = 0; k < kBC; k++) {
            const int c = c0 + k;
            if (in(c, K.cw)) {
                double s = u2d((uint32_t)byte_of(wv, k));
                if (hp) {
                    const double lpv = pole(*hp, s, a_hp);
                    s = dadd(s, dsub(s, lpv));
                }
                s = pole(lp[0], s, a_lp);
                s = pole(lp[1], s, a_lp);
                s = pole(lp[2], s, a_lp);
                if (!EDGE || c >= d) ring[(t + k) & (kRingC - 1)] = (uint8_t)q8(s);
            }
        }
    }

    // composite_video_yuv_to_ntsc for one pixel: y += chroma / 50                        (:434-477)
    static CVS_HD int modulate_px(const K422 &K, const Row422 &rc, int j, int u, int v, int y) {
        int q = (u - 128) * rc.mU[j & 3] + (v - 128) * rc.mV[j & 3];
        if (EDGE && K.amp != 50) q = div50(q * K.amp);
        return clamp8(y + q);
    }

    // G0 + G1: block s enters the rings; input chroma lowpass
    static CVS_HD void stage_load(const K422 &K, const Geo &G, Lane422 &ln, int s, const StepIO &io) {
        if (!EDGE || s < G.nbl) {
            uint8_t *yb = yblk(ln.ry, s);
            stw(yb, io.y0);
            stw(yb + 4, io.y1);
            stw(cblk(ln.ru, s), io.u);
            stw(cblk(ln.rv, s), io.v);
        }
        if ((K.flags & G_IN_LP) && (!EDGE || s < G.nb)) {
            chroma_lp4(K, ln.ru, s, io.u, &ln.inU[0], &ln.inU[1], K.a_in[0], K.a_inhp[0], K.d_in[0]);
            chroma_lp4(K, ln.rv, s, io.v, &ln.inV[0], &ln.inV[1], K.a_in[1], K.a_inhp[1], K.d_in[1]);
        }
    }

    // G2: block b of the composite signal.  FIRST: the full chain (pre-emphasis, noise, head switch);
    // !FIRST: plain re-modulation of a block whose luma is final (VHS recombine, -yc-recomb).
    template <bool FIRST>
    static CVS_HD void stage_modulate(const K422 &K, const Row422 &rc, Lane422 &ln, int b, bool warp_hs,
                                      const uint8_t *hsrow) {
        const int x0 = b * kB;
        uint8_t *yb = yblk(ln.ry, b), *ub = cblk(ln.ru, b), *vb = cblk(ln.rv, b);
        const uint32_t y0 = ldw(yb), y1 = ldw(yb + 4), uw = ldw(ub), vw = ldw(vb);
        uint32_t *grp = nullptr, *grp_next = nullptr;
        if (FIRST && K.vnoise != 0) {
            grp = ln.rngL.group_ptr(kRngBase + (uint32_t)x0);
            grp_next = ln.rngL.group_ptr(kRngBase + (uint32_t)x0 + kB);
        }
        int yo[kB];
        CVS_UNROLL
        for (int j = 0; j < kB; j++) {
            const int x = x0 + j;
            int y = byte_of(j < 4 ? y0 : y1, j & 3);
            if (in(x, K.w)) {
                y = modulate_px(K, rc, j, byte_of(uw, j >> 1), byte_of(vw, j >> 1), y);
                if (FIRST) {
                    if (EDGE && (K.flags & G_PREEMPH)) {                       // (:636-650)
                        double sv = u2d((uint32_t)y);
                        const double lpv = pole(ln.pre, sv, K.a_pre);
                        sv = dadd(sv, dmul(dsub(sv, lpv), K.preemph));
                        y = q8(sv);
                    }
                    if (K.vnoise != 0) {                                       // (:653-665)
                        y = clamp8(y + ln.nY);
                        const int d = draw_mod(ln.rngL.next_in_group(grp, grp_next, j, kB), (uint32_t)(2 * K.vnoise + 1), K.vmagic, K.vshift);
                        ln.nY = noise_step(ln.nY, d, K.vnoise);
                    }
                    if (warp_hs) {                                             // (:697-731) as a delay line
                        ln.rya[x & (kRingA - 1)] = (uint8_t)y;
                        if (rc.hs_delay > 0) y = (x >= rc.hs_delay) ? ln.rya[(x - rc.hs_delay) & (kRingA - 1)] : 16;
                    }
                    if (EDGE && hsrow) y = hsrow[x];                           // rotated row from the pre-pass
                }
            }
            yo[j] = y;                                                         // bytes past the row stay (demod reads two)
        }
        stw(yb, pack4(yo[0], yo[1], yo[2], yo[3]));
        stw(yb + 4, pack4(yo[4], yo[5], yo[6], yo[7]));
        if (EDGE && (K.flags & G_NOCOLOR)) {                                   // (:473-474); chroma past the row is never read
            stw(ub, 0x80808080u);
            stw(vb, 0x80808080u);
        }
    }

    // composite_ntsc_to_yuv on one block (:480-553).  y0,y1 = the block, y2 = the next block (two bytes used).
    static CVS_HD void demod8(const K422 &K, const Row422 &rc, Demod &dm, int b, uint32_t y0, uint32_t y1, uint32_t y2,
                              int divisor, uint32_t dmagic, uint32_t dshift, int Yn[kB], int U[kBC], int V[kBC]) {
        const int x0 = b * kB;
        int ch[kB];
        if (EDGE && b == 0) dm.reset(byte_of(y0, 0), byte_of(y0, 1));
        const int xflip0 = ((4 - rc.xi) & 3) + 2;          // first flipped index (:527)
        CVS_UNROLL
        for (int j = 0; j < kB; j++) {
            const int x = x0 + j;
            const int c = (j < 2) ? byte_of(y0, j + 2) : ((j < 6) ? byte_of(y1, j - 2) : byte_of(y2, j - 6));
            ch[j] = 128;
            Yn[j] = byte_of(j < 4 ? y0 : y1, j & 3);
            if (in(x, K.w)) {
                const int sum = dm.o1 + dm.o2 + dm.o3 + c;
                dm.o1 = dm.o2; dm.o2 = dm.o3; dm.o3 = c;
                const int yn = sum >> 2;
                int cv = clamp8(c + 128 - yn);
                if (EDGE && (K.flags & G_NOCOLOR_YC)) {
                    Yn[j] = cv;                                                // (:504-508)
                } else {
                    Yn[j] = yn;
                    if (!EDGE || x >= xflip0) cv ^= rc.fl[j & 3];
                    if (EDGE && divisor != 50) cv = clamp8(div_trunc((cv - 128) * 50, divisor, dmagic, dshift) + 128);
                }
                ch[j] = cv ^ 0xFF;                                             // 255 - chroma (:539-548)
            }
        }
        const bool odd = (rc.xi & 1) != 0;
        CVS_UNROLL
        for (int k = 0; k < kBC; k++) {
            if (EDGE && (K.flags & G_NOCOLOR_YC)) { U[k] = 128; V[k] = 128; }
            else {
                U[k] = odd ? ch[2 * k + 1] : ch[2 * k];
                V[k] = odd ? ch[2 * k] : ch[2 * k + 1];
            }
        }
    }

    // luma bytes of a block back into the ring; in the last block the bytes past the row keep their value
    static CVS_HD void store_luma(const K422 &K, uint8_t *yb, int x0, const int Y[kB], uint32_t y0, uint32_t y1) {
        int o[kB];
        CVS_UNROLL
        for (int j = 0; j < kB; j++) o[j] = in(x0 + j, K.w) ? Y[j] : byte_of(j < 4 ? y0 : y1, j & 3);
        stw(yb, pack4(o[0], o[1], o[2], o[3]));
        stw(yb + 4, pack4(o[4], o[5], o[6], o[7]));
    }

    // G3: first demodulation of block b and everything pointwise after it
    static CVS_HD void stage_separate(const K422 &K, const Row422 &rc, Lane422 &ln, int b,
                                      uint32_t amag, uint32_t ashift) {
        const int x0 = b * kB, c0 = b * kBC;
        uint8_t *yb = yblk(ln.ry, b), *ub = cblk(ln.ru, b), *vb = cblk(ln.rv, b);
        const uint32_t y0 = ldw(yb), y1 = ldw(yb + 4);
        int Yn[kB], U[kBC], V[kBC];
        if (!EDGE || !(K.flags & G_NOCOLOR)) {
            const uint32_t y2 = ldw(yblk(ln.ry, b + 1));
            demod8(K, rc, ln.dm1, b, y0, y1, y2, K.amp_back, amag, ashift, Yn, U, V);
        } else {
            const uint32_t uw = ldw(ub), vw = ldw(vb);
            CVS_UNROLL
            for (int j = 0; j < kB; j++) Yn[j] = byte_of(j < 4 ? y0 : y1, j & 3);
            CVS_UNROLL
            for (int k = 0; k < kBC; k++) { U[k] = byte_of(uw, k); V[k] = byte_of(vw, k); }
        }
        if (K.cnoise != 0) {                                                   // (:738-754)
            uint32_t *grp = ln.rngC.group_ptr(kRngBase + (uint32_t)(2 * c0));
            uint32_t *grp_next = ln.rngC.group_ptr(kRngBase + (uint32_t)(2 * c0) + 2 * kBC);
            CVS_UNROLL
            for (int k = 0; k < kBC; k++) {
                if (in(c0 + k, K.cw)) {
                    U[k] = clamp8(U[k] + ln.nU);
                    V[k] = clamp8(V[k] + ln.nV);
                    const int du = draw_mod(ln.rngC.next_in_group(grp, grp_next, 2 * k, 2 * kBC), (uint32_t)(2 * K.cnoise + 1), K.cmagic, K.cshift);
                    ln.nU = noise_step(ln.nU, du, K.cnoise);
                    const int dv = draw_mod(ln.rngC.next_in_group(grp, grp_next, 2 * k + 1, 2 * kBC), (uint32_t)(2 * K.cnoise + 1), K.cmagic, K.cshift);
                    ln.nV = noise_step(ln.nV, dv, K.cnoise);
                }
            }
        }
        if (K.flags & G_PHASE) {                                               // (:755-783), rotation as written there
            CVS_UNROLL
            for (int k = 0; k < kBC; k++) {
                const double u = i2d(U[k] - 128), v = i2d(V[k] - 128);
                const double u_ = dsub(dmul(u, rc.cosp), dmul(u, rc.sinp));
                const double v_ = dadd(dmul(v, rc.cosp), dmul(v, rc.sinp));
                U[k] = q8(dadd(u_, 128.0));
                V[k] = q8(dadd(v_, 128.0));
            }
        }
        // the unfiltered values go into the ring: a delayed filter never overwrites the last `delay` samples
        const uint32_t uw = pack4(U[0], U[1], U[2], U[3]), vw = pack4(V[0], V[1], V[2], V[3]);
        stw(ub, uw);
        stw(vb, vw);
        if (K.flags & G_VHS) {
            CVS_UNROLL
            for (int j = 0; j < kB; j++) {                                     // luma lowpass + boost (:810-828)
                if (in(x0 + j, K.w)) {
                    double s = u2d((uint32_t)Yn[j]);
                    s = pole(ln.lum[0], s, K.a_luma);
                    s = pole(ln.lum[1], s, K.a_luma);
                    s = pole(ln.lum[2], s, K.a_luma);
                    const double lpv = pole(ln.lum[3], s, K.a_luma);
                    Yn[j] = q8(dadd(s, dmul(dsub(s, lpv), 1.6)));
                }
            }
            store_luma(K, yb, x0, Yn, y0, y1);
        } else {
            store_luma(K, yb, x0, Yn, y0, y1);
        }
    }
    // G3b: VHS chroma lowpass of block b, written cd samples back (:830-851)
    static CVS_HD void stage_chroma_lp(const K422 &K, Lane422 &ln, int b) {
        const uint32_t uw = ldw(cblk(ln.ru, b)), vw = ldw(cblk(ln.rv, b));
        chroma_lp4(K, ln.ru, b, uw, nullptr, ln.chU, K.a_ch, 0.0, K.cd);
        chroma_lp4(K, ln.rv, b, vw, nullptr, ln.chV, K.a_ch, 0.0, K.cd);
    }
    // G4a: VHS luma sharpen of block b (:888-901); the lowpassed luma is an 8-bit plane in the reference too
    static CVS_HD void stage_luma_sharpen(const K422 &K, Lane422 &ln, int b) {
        const int x0 = b * kB;
        uint8_t *yb = yblk(ln.ry, b);
        const uint32_t y0 = ldw(yb), y1 = ldw(yb + 4);
        int Yn[kB];
        CVS_UNROLL
        for (int j = 0; j < kB; j++) {
            Yn[j] = byte_of(j < 4 ? y0 : y1, j & 3);
            if (in(x0 + j, K.w)) {
                const double y1d = u2d((uint32_t)Yn[j]);
                double ts = pole(ln.lsh[0], y1d, K.a_lsharp);
                ts = pole(ln.lsh[1], ts, K.a_lsharp);
                ts = pole(ln.lsh[2], ts, K.a_lsharp);
                Yn[j] = q8(dadd(y1d, dmul(dsub(y1d, ts), K.sharpen)));
            }
        }
        store_luma(K, yb, x0, Yn, y0, y1);
    }

    // G4 part 1: the lane's own chroma of block b before the vertical blend (what the lane below needs)
    static CVS_HD void blend_fetch(const Lane422 &ln, int b, uint32_t &pu, uint32_t &pv) {
        pu = ldw(cblk(ln.ru, b));
        pv = ldw(cblk(ln.rv, b));
    }
    // G4 part 2: blend with the row above (:858-883), chroma sharpen (:904-925), re-modulate (:927-930)
    static CVS_HD void stage_vhs_chroma(const K422 &K, const Row422 &rc, Lane422 &ln, int b, uint32_t pu, uint32_t pv,
                                        uint32_t au, uint32_t av) {
        const int c0 = b * kBC;
        int U[kBC], V[kBC];
        CVS_UNROLL
        for (int k = 0; k < kBC; k++) {
            U[k] = byte_of(pu, k);
            V[k] = byte_of(pv, k);
            if (in(c0 + k, K.cw)) {
                int u = U[k], v = V[k];
                if ((K.flags & G_VBLEND) && rc.row >= 1) {
                    // the delay line starts at 128 and row 0 never enters it
                    const int ua = (rc.row == 1) ? 128 : byte_of(au, k), va = (rc.row == 1) ? 128 : byte_of(av, k);
                    u = (ua + u + 1) >> 1;
                    v = (va + v + 1) >> 1;
                }
                double s = u2d((uint32_t)u), ts;
                ts = pole(ln.csU[0], s, K.a_csharp); ts = pole(ln.csU[1], ts, K.a_csharp); ts = pole(ln.csU[2], ts, K.a_csharp);
                U[k] = q8(dadd(s, dmul(dsub(s, ts), K.sharpen_c)));
                s = u2d((uint32_t)v);
                ts = pole(ln.csV[0], s, K.a_csharp); ts = pole(ln.csV[1], ts, K.a_csharp); ts = pole(ln.csV[2], ts, K.a_csharp);
                V[k] = q8(dadd(s, dmul(dsub(s, ts), K.sharpen_c)));
            }
        }
        stw(cblk(ln.ru, b), pack4(U[0], U[1], U[2], U[3]));
        stw(cblk(ln.rv, b), pack4(V[0], V[1], V[2], V[3]));
    }

    // a demodulation whose result goes straight back into the rings (VHS recombine, -yc-recomb)
    static CVS_HD void stage_redemod(const K422 &K, const Row422 &rc, Lane422 &ln, Demod &dm, int b,
                                     uint32_t amag, uint32_t ashift) {
        uint8_t *yb = yblk(ln.ry, b);
        const uint32_t y0 = ldw(yb), y1 = ldw(yb + 4), y2 = ldw(yblk(ln.ry, b + 1));
        int Yn[kB], U[kBC], V[kBC];
        demod8(K, rc, dm, b, y0, y1, y2, K.amp, amag, ashift, Yn, U, V);
        store_luma(K, yb, b * kB, Yn, y0, y1);
        stw(cblk(ln.ru, b), pack4(U[0], U[1], U[2], U[3]));
        stw(cblk(ln.rv, b), pack4(V[0], V[1], V[2], V[3]));
    }

    static CVS_HD void stage_dropout(Lane422 &ln, int b) {                      // (:932-941)
        stw(cblk(ln.ru, b), 0x80808080u);
        stw(cblk(ln.rv, b), 0x80808080u);
    }
};

// ---- the interior variant: whole blocks inside the row, common switches only --------------------------
// Same arithmetic, same ring contents at every step boundary as Pipe422<false>; what differs is the shape of
// the code (see "interior-variant helpers").  tests/test_yuv422_emu.py runs it on the CPU against the oracle.
struct Fast422 {
    // the fast kernel carries u = p - p a instead of the state p of every pole (cascade())
    static CVS_HD void prime(const K422 &K, Lane422 &ln) {
        ln.inU[0] = carry_of(ln.inU[0], K.a_inhp[0]);
        ln.inV[0] = carry_of(ln.inV[0], K.a_inhp[1]);
        ln.outU[0] = carry_of(ln.outU[0], K.a_outhp[0]);
        ln.outV[0] = carry_of(ln.outV[0], K.a_outhp[1]);
        for (int k = 1; k < 4; k++) {
            ln.inU[k] = carry_of(ln.inU[k], K.a_in[0]);
            ln.inV[k] = carry_of(ln.inV[k], K.a_in[1]);
            ln.outU[k] = carry_of(ln.outU[k], K.a_out[0]);
            ln.outV[k] = carry_of(ln.outV[k], K.a_out[1]);
        }
        for (int k = 0; k < 4; k++) ln.lum[k] = carry_of(ln.lum[k], K.a_luma);
        for (int k = 0; k < 3; k++) {
            ln.lsh[k] = carry_of(ln.lsh[k], K.a_lsharp);
            ln.chU[k] = carry_of(ln.chU[k], K.a_ch);
            ln.chV[k] = carry_of(ln.chV[k], K.a_ch);
            ln.csU[k] = carry_of(ln.csU[k], K.a_csharp);
            ln.csV[k] = carry_of(ln.csV[k], K.a_csharp);
        }
    }
    // a delayed filter's block: positions 4b - d + k.  The word that is complete now is stored; with d % 4 != 0
    // the bytes of the following word wait in `carry` (= the previous block's outputs)
    static CVS_HD void put_delayed(uint8_t *ring, int b, int d, uint32_t nw, uint32_t &carry) {
        const int r = d & 3, qd = d >> 2;
        if (r == 0) {
            stw(cblk(ring, b - qd), nw);
        } else {
            stw(cblk(ring, b - qd - 1), funnel_r(carry, nw, 8 * r));
            carry = nw;
        }
    }
    // after the row's last block b: the waiting bytes go to their places one by one
    static CVS_HD void carry_leave(uint8_t *ring, int b, int d, uint32_t carry) {
        const int r = d & 3, qd = d >> 2;
        if (r == 0) return;
        for (int k = r; k < 4; k++) ring[(kBC * (b - qd) + k - r) & (kRingC - 1)] = (uint8_t)((carry >> (8 * k)) & 0xFFu);
    }

    // three-pole lowpass (BOOST: preceded by s += s - highpass(s)) on the four samples of wu and of wv
    template <bool BOOST, int CU = CVS422_CU>
    static CVS_HD void lp_pair(uint32_t wu, uint32_t wv, double *hpU, double *lpU, double *hpV, double *lpV, double aU, double ahU,
                               double aV, double ahV, uint32_t &ou, uint32_t &ov) {
        ou = ov = 0;
        CVS_ROLLED
        for (int it = 0; it < kBC / CU; it++) {
            int qu[CU], qv[CU];
            CVS_UNROLL
            for (int j = 0; j < CU; j++) {
                double s = fu2d((uint32_t)byte_of(wu, j)), t = fu2d((uint32_t)byte_of(wv, j));
                if (BOOST) {
                    const double lu = cascade<1>(hpU, s, ahU), lv = cascade<1>(hpV, t, ahV);
                    s = dadd(s, dsub(s, lu));
                    t = dadd(t, dsub(t, lv));
                }
                s = cascade<3>(lpU, s, aU);
                t = cascade<3>(lpV, t, aV);
                qu[j] = fq(s);
                qv[j] = fq(t);
            }
            push_c<CU>(ou, qu);
            push_c<CU>(ov, qv);
            if (CU < 4) { wu >>= (8 * CU) & 31; wv >>= (8 * CU) & 31; }
        }
    }
    // CU clamped bytes enter a word from the top (after 4 / CU pushes the first one is byte 0)
    template <int CU>
    static CVS_HD void push_c(uint32_t &acc, const int *q) {
        if (CU == 4) acc = sat4(q[0], q[1 % CU], q[2 % CU], q[3 % CU]);
        else if (CU == 2) acc = funnel_r(acc, sat_pack2(q[1 % CU], q[0], 0u), 16);
        else acc = funnel_r(acc, (uint32_t)sat1(q[0]), 8);
    }

    // y + chroma on the carrier for the 8 pixels of a block (amp == 50), not yet clamped
    static CVS_HD void modulate8(const Row422 &rc, uint32_t y0, uint32_t y1, uint32_t uw, uint32_t vw, int yo[kB]) {
        const uint32_t t_lo = prmt(uw, vw, rc.selM), t_hi = prmt(uw, vw, rc.selM + 0x2222u);
        CVS_UNROLL
        for (int j = 0; j < kB; j++)
            yo[j] = byte_of(j < 4 ? y0 : y1, j & 3) + (byte_of(j < 4 ? t_lo : t_hi, j & 3) ^ rc.mX[j & 3]) + rc.mC[j & 3];
    }

    // G2, first modulation: + luma noise, + head-switch delay
    static CVS_HD void stage_modulate_first(const K422 &K, const Row422 &rc, Lane422 &ln, int b, bool warp_hs) {
        const int x0 = b * kB;
        uint8_t *yb = yblk(ln.ry, b);
        int yo[kB];
        modulate8(rc, ldw(yb), ldw(yb + 4), ldw(cblk(ln.ru, b)), ldw(cblk(ln.rv, b)), yo);
        if (K.vnoise != 0) {                                                   // (:653-665)
            uint32_t *grp = ln.rngL.group_ptr(kRngBase + (uint32_t)x0);
            const uint32_t *grp_next = ln.rngL.group_ptr(kRngBase + (uint32_t)x0 + kB);
            CVS_UNROLL
            for (int j = 0; j < kB; j++) {
                yo[j] = sat1(yo[j]) + ln.nY;
                ln.nY = cvs::noise_draw(ln.nY, ln.rngL.next_in_group(grp, grp_next, j, kB), (uint32_t)(2 * K.vnoise + 1), K.vmagic, K.vshift, K.vnoise);
            }
        }
        uint32_t w0 = sat4(yo[0], yo[1], yo[2], yo[3]), w1 = sat4(yo[4], yo[5], yo[6], yo[7]);
        if (warp_hs) {                                                         // (:697-731) as a delay line
            uint8_t *ab = ln.rya + (x0 & (kRingA - 1));
            stw(ab, w0);
            stw(ab + 4, w1);
            if (rc.hs_delay > 0) {
                const int p = x0 - rc.hs_delay, sh = 8 * (p & 3);
                const uint32_t a0 = ldw(ln.rya + (p & (kRingA - 4))), a1 = ldw(ln.rya + ((p + 4) & (kRingA - 4))),
                               a2 = ldw(ln.rya + ((p + 8) & (kRingA - 4)));
                w0 = sh ? funnel_r(a0, a1, sh) : a0;
                w1 = sh ? funnel_r(a1, a2, sh) : a1;
                const int n16 = -p;                                            // pixels of this block before x = delay: 16
                if (n16 > 0) {
                    const uint32_t m0 = n16 >= 4 ? 0xFFFFFFFFu : ((1u << (8 * n16)) - 1u);
                    const uint32_t m1 = n16 >= 8 ? 0xFFFFFFFFu : (n16 > 4 ? ((1u << (8 * (n16 - 4))) - 1u) : 0u);
                    w0 = (w0 & ~m0) | (0x10101010u & m0);
                    w1 = (w1 & ~m1) | (0x10101010u & m1);
                }
            }
        }
        stw(yb, w0);
        stw(yb + 4, w1);
    }
    // re-modulation of a block whose luma is final (VHS recombine)
    static CVS_HD void stage_remodulate(const Row422 &rc, Lane422 &ln, int b) {
        uint8_t *yb = yblk(ln.ry, b);
        int yo[kB];
        modulate8(rc, ldw(yb), ldw(yb + 4), ldw(cblk(ln.ru, b)), ldw(cblk(ln.rv, b)), yo);
        stw(yb, sat4(yo[0], yo[1], yo[2], yo[3]));
        stw(yb + 4, sat4(yo[4], yo[5], yo[6], yo[7]));
    }

    // 8 luma pixels in yw0:yw1 -> the same words through a rolled per-pixel loop; F(double) -> double is the filter
    template <class F>
    static CVS_HD void luma8(uint32_t &yw0, uint32_t &yw1, F f) {
        uint32_t o0 = 0, o1 = 0;
        CVS_ROLLED
        for (int it = 0; it < kB / CVS422_LU; it++) {
            int q[CVS422_LU];
            CVS_UNROLL
            for (int j = 0; j < CVS422_LU; j++) q[j] = fq(f(fu2d((uint32_t)byte_of(j < 4 ? yw0 : yw1, j & 3))));
            if (CVS422_LU == 8) {
                o0 = sat4(q[0], q[1 % CVS422_LU], q[2 % CVS422_LU], q[3 % CVS422_LU]);
                o1 = sat4(q[4 % CVS422_LU], q[5 % CVS422_LU], q[6 % CVS422_LU], q[7 % CVS422_LU]);
            } else if (CVS422_LU == 4) {
                o0 = o1;
                o1 = sat4(q[0], q[1 % CVS422_LU], q[2 % CVS422_LU], q[3 % CVS422_LU]);
                yw0 = yw1;
            } else {
                const int sh = (8 * CVS422_LU) & 31;
                const uint32_t nb = (CVS422_LU == 2) ? sat_pack2(q[1 % CVS422_LU], q[0], 0u) : (uint32_t)sat1(q[0]);
                o0 = funnel_r(o0, o1, sh);
                o1 = funnel_r(o1, nb, sh);
                yw0 = funnel_r(yw0, yw1, sh);
                yw1 >>= sh;
            }
        }
        yw0 = o0;
        yw1 = o1;
    }
    struct LumaLp {                       // VHS luma lowpass + boost (:810-828)
        Lane422 &ln; const K422 &K;
        CVS_HD double operator()(double s) const {
            // four poles of one alpha: the boost's lowpass continues the cascade (its input is the third pole's output)
            double t = dmul(s, K.a_luma), p = 0, p2 = 0;
            CVS_UNROLL
            for (int k = 0; k < 4; k++) {
                p = dadd(t, ln.lum[k]);
                t = dmul(p, K.a_luma);
                ln.lum[k] = dsub(p, t);
                if (k == 2) p2 = p;
            }
            return dadd(p2, dmul(dsub(p2, p), 1.6));
        }
    };
    struct LumaSharpen {                  // VHS luma sharpen (:888-901)
        Lane422 &ln; const K422 &K;
        CVS_HD double operator()(double y1d) const {
            const double ts = cascade<3>(ln.lsh, y1d, K.a_lsharp);
            return dadd(y1d, dmul(dsub(y1d, ts), K.sharpen));
        }
    };
    // G4a
    static CVS_HD void stage_luma_sharpen(const K422 &K, Lane422 &ln, int b) {
        uint8_t *yb = yblk(ln.ry, b);
        uint32_t yw0 = ldw(yb), yw1 = ldw(yb + 4);
        luma8(yw0, yw1, LumaSharpen{ln, K});
        stw(yb, yw0);
        stw(yb + 4, yw1);
    }

    // composite_ntsc_to_yuv on one whole block (:480-553), words in, words out.  Block 0 starts the box filter from the
    // row's first two samples and leaves the pixels before the first flip index alone (:487-499, :527).
    static CVS_HD void demod_w(const Row422 &rc, Demod &dm, int b, uint32_t y0, uint32_t y1, uint32_t y2,
                               uint32_t &yw0, uint32_t &yw1, uint32_t &uw, uint32_t &vw) {
        uint32_t f_lo = rc.flx, f_hi = rc.flx;
        if (b == 0) {
            dm.reset(byte_of(y0, 0), byte_of(y0, 1));
            const int xflip0 = ((4 - rc.xi) & 3) + 2;                         // 2..5
            f_lo |= (xflip0 >= 4) ? 0xFFFFFFFFu : ((1u << (8 * xflip0)) - 1u);
            f_hi |= (xflip0 > 4) ? 0xFFu : 0u;
        }
        int yn[kB], cv[kB];
        CVS_UNROLL
        for (int j = 0; j < kB; j++) {
            const int c = (j < 2) ? byte_of(y0, j + 2) : ((j < 6) ? byte_of(y1, j - 2) : byte_of(y2, j - 6));
            const int sum = dm.o1 + dm.o2 + dm.o3 + c;
            dm.o1 = dm.o2; dm.o2 = dm.o3; dm.o3 = c;
            yn[j] = sum >> 2;
            cv[j] = c + 128 - yn[j];
        }
        yw0 = sat4(yn[0], yn[1], yn[2], yn[3]);
        yw1 = sat4(yn[4], yn[5], yn[6], yn[7]);
        const uint32_t c_lo = sat4(cv[0], cv[1], cv[2], cv[3]) ^ f_lo, c_hi = sat4(cv[4], cv[5], cv[6], cv[7]) ^ f_hi;
        uw = prmt(c_lo, c_hi, rc.selU);                                       // 255 - chroma (:539-548)
        vw = prmt(c_lo, c_hi, rc.selV);
    }

    // G3
    static CVS_HD void stage_separate(const K422 &K, const Row422 &rc, Lane422 &ln, int b) {
        const int c0 = b * kBC;
        uint8_t *yb = yblk(ln.ry, b), *ub = cblk(ln.ru, b), *vb = cblk(ln.rv, b);
        uint32_t yw0, yw1, uw, vw;
        demod_w(rc, ln.dm1, b, ldw(yb), ldw(yb + 4), ldw(yblk(ln.ry, b + 1)), yw0, yw1, uw, vw);
        if (K.cnoise != 0 || (K.flags & G_PHASE)) {
            int U[kBC], V[kBC];
            CVS_UNROLL
            for (int k = 0; k < kBC; k++) { U[k] = byte_of(uw, k); V[k] = byte_of(vw, k); }
            if (K.cnoise != 0) {                                               // (:738-754)
                uint32_t *grp = ln.rngC.group_ptr(kRngBase + (uint32_t)(2 * c0));
                const uint32_t *grp_next = ln.rngC.group_ptr(kRngBase + (uint32_t)(2 * c0) + 2 * kBC);
                CVS_UNROLL
                for (int k = 0; k < kBC; k++) {
                    U[k] = sat1(U[k] + ln.nU);
                    V[k] = sat1(V[k] + ln.nV);
                    ln.nU = cvs::noise_draw(ln.nU, ln.rngC.next_in_group(grp, grp_next, 2 * k, 2 * kBC), (uint32_t)(2 * K.cnoise + 1), K.cmagic, K.cshift, K.cnoise);
                    ln.nV = cvs::noise_draw(ln.nV, ln.rngC.next_in_group(grp, grp_next, 2 * k + 1, 2 * kBC), (uint32_t)(2 * K.cnoise + 1), K.cmagic, K.cshift, K.cnoise);
                }
            }
            if ((K.flags & G_PHASE) && rc.pmap) {                              // (:755-783), tabulated
                CVS_UNROLL
                for (int k = 0; k < kBC; k++) {
                    U[k] = rc.pmap[U[k]];
                    V[k] = rc.pmap[256 + V[k]];
                }
            } else if (K.flags & G_PHASE) {
                CVS_UNROLL
                for (int k = 0; k < kBC; k++) {
                    const double u = i2d(U[k] - 128), v = i2d(V[k] - 128);
                    const double u_ = dsub(dmul(u, rc.cosp), dmul(u, rc.sinp));
                    const double v_ = dadd(dmul(v, rc.cosp), dmul(v, rc.sinp));
                    U[k] = fq(dadd(u_, 128.0));
                    V[k] = fq(dadd(v_, 128.0));
                }
            }
            uw = sat4(U[0], U[1], U[2], U[3]);
            vw = sat4(V[0], V[1], V[2], V[3]);
        }
        stw(ub, uw);                                                           // unfiltered: see Pipe422::stage_separate
        stw(vb, vw);
        if (K.flags & G_VHS) {
            luma8(yw0, yw1, LumaLp{ln, K});
        }
        stw(yb, yw0);
        stw(yb + 4, yw1);
    }
    // G3b
    static CVS_HD void stage_chroma_lp(const K422 &K, Lane422 &ln, int b) {
        uint32_t ou, ov;
        lp_pair<false, CVS422_CU_MID>(ldw(cblk(ln.ru, b)), ldw(cblk(ln.rv, b)), nullptr, ln.chU, nullptr, ln.chV, K.a_ch, 0.0, K.a_ch, 0.0, ou, ov);
        put_delayed(ln.ru, b, K.cd, ou, ln.cCh[0]);
        put_delayed(ln.rv, b, K.cd, ov, ln.cCh[1]);
    }

    // G4: vertical blend (:858-883), chroma sharpen (:904-925)
    static CVS_HD void stage_vhs_chroma(const K422 &K, const Row422 &rc, Lane422 &ln, int b, uint32_t pu, uint32_t pv,
                                        uint32_t au, uint32_t av) {
        if ((K.flags & G_VBLEND) && rc.row >= 1) {
            if (rc.row == 1) au = av = 0x80808080u;                            // the delay line starts at 128
            pu = avg4_up(au, pu);
            pv = avg4_up(av, pv);
        }
        uint32_t ou = 0, ov = 0;
        CVS_ROLLED
        for (int it = 0; it < kBC / CVS422_CU_MID; it++) {
            int qu[CVS422_CU_MID], qv[CVS422_CU_MID];
            CVS_UNROLL
            for (int j = 0; j < CVS422_CU_MID; j++) {
                const double s = fu2d((uint32_t)byte_of(pu, j)), t = fu2d((uint32_t)byte_of(pv, j));
                const double ts = cascade<3>(ln.csU, s, K.a_csharp), tt = cascade<3>(ln.csV, t, K.a_csharp);
                qu[j] = fq(dadd(s, dmul(dsub(s, ts), K.sharpen_c)));
                qv[j] = fq(dadd(t, dmul(dsub(t, tt), K.sharpen_c)));
            }
            push_c<CVS422_CU_MID>(ou, qu);
            push_c<CVS422_CU_MID>(ov, qv);
            if (CVS422_CU_MID < 4) { pu >>= (8 * CVS422_CU_MID) & 31; pv >>= (8 * CVS422_CU_MID) & 31; }
        }
        stw(cblk(ln.ru, b), ou);
        stw(cblk(ln.rv, b), ov);
    }

    // G5
    static CVS_HD void stage_redemod(const Row422 &rc, Lane422 &ln, int b) {
        uint8_t *yb = yblk(ln.ry, b);
        uint32_t yw0, yw1, uw, vw;
        demod_w(rc, ln.dm2, b, ldw(yb), ldw(yb + 4), ldw(yblk(ln.ry, b + 1)), yw0, yw1, uw, vw);
        stw(yb, yw0);
        stw(yb + 4, yw1);
        stw(cblk(ln.ru, b), uw);
        stw(cblk(ln.rv, b), vw);
    }

    // GO
    static CVS_HD void stage_out(const K422 &K, Lane422 &ln, int b) {
        const uint32_t uw = ldw(cblk(ln.ru, b)), vw = ldw(cblk(ln.rv, b));
        uint32_t ou, ov;
        if (K.flags & G_OUT_FULL) lp_pair<true>(uw, vw, &ln.outU[0], &ln.outU[1], &ln.outV[0], &ln.outV[1], K.a_out[0], K.a_outhp[0], K.a_out[1], K.a_outhp[1], ou, ov);
        else lp_pair<false>(uw, vw, nullptr, &ln.outU[1], nullptr, &ln.outV[1], K.a_out[0], 0.0, K.a_out[1], 0.0, ou, ov);
        put_delayed(ln.ru, b, K.d_out[0], ou, ln.cOut[0]);
        put_delayed(ln.rv, b, K.d_out[1], ov, ln.cOut[1]);
    }
};

// ---- roles: the general kernel's steps -----------------------------------------------------------------------
// One step of each role in the general variant (any block, any switch); the fast kernel's steps follow.
// role 0: block s of the source row comes in
CVS_HD void role0_step(const K422 &K, const Lags &L, const Geo &G, const Row422 &rc, Lane422 &ln, int s, const StepIO &in,
                       bool warp_hs, const uint8_t *hsrow) {
    const int bM = s - L.bM;
    Pipe422<true>::stage_load(K, G, ln, s, in);
    if (bM >= 0 && bM < G.nb) Pipe422<true>::template stage_modulate<true>(K, rc, ln, bM, warp_hs, hsrow);
#if CVS422_LSHARP_ROLE == 0
    const int bC = s - L.bC;
    if ((K.flags & G_VHS) && bC >= 0 && bC < G.nb) Pipe422<true>::stage_luma_sharpen(K, ln, bC);
#endif
}
// role 1
CVS_HD void role1_step(const K422 &K, const Lags &L, const Geo &G, const DivPair &dv, const Row422 &rc, Lane422 &ln, int s) {
    const int bD = s - L.bD;
    if (bD >= 0 && bD < G.nb) Pipe422<true>::stage_separate(K, rc, ln, bD, dv.back_magic, dv.back_shift);
}
// role 2, front half: chroma lowpass, luma sharpen; (pu, pv) = the lane's chroma of block s - bV before the vertical
// blend, which the lane below needs between the two halves
CVS_HD void role2_front(const K422 &K, const Lags &L, const Geo &G, Lane422 &ln, int s, uint32_t &pu, uint32_t &pv) {
    pu = pv = 0;
    if (!(K.flags & G_VHS)) return;
    const int bC = s - L.bC, bV = s - L.bV;
    if (bC >= 0 && bC < G.nb) {
        Pipe422<true>::stage_chroma_lp(K, ln, bC);
#if CVS422_LSHARP_ROLE == 2
        Pipe422<true>::stage_luma_sharpen(K, ln, bC);
#endif
    }
    if (bV >= 0 && bV < G.nb) Pipe422<true>::blend_fetch(ln, bV, pu, pv);
}
CVS_HD void role2_back(const K422 &K, const Lags &L, const Geo &G, const Row422 &rc, Lane422 &ln, int s, uint32_t pu, uint32_t pv,
                       uint32_t au, uint32_t av) {
    if (!(K.flags & G_VHS)) return;
    const int bV = s - L.bV;
    if (bV >= 0 && bV < G.nb) Pipe422<true>::stage_vhs_chroma(K, rc, ln, bV, pu, pv, au, av);
}
// role 3.  Returns true when `out` holds block bs of the finished row.
CVS_HD bool role3_step(const K422 &K, const Lags &L, const Geo &G, const DivPair &dv, const Row422 &rc, Lane422 &ln, int s,
                       StepIO &out, int &bs) {
    typedef Pipe422<true> P;
    const int nb = G.nb;
    const bool redemod = (K.flags & G_VHS) && !(K.flags & G_SVIDEO);
    const int bR = s - L.bR, b2 = s - L.bD2, bE = s - L.bE, bF = s - L.bF;
    bs = s - L.bS;
    if (redemod && bR >= 0 && bR < nb) P::template stage_modulate<false>(K, rc, ln, bR, false, nullptr);   // (:927-930)
    if (redemod && b2 >= 0 && b2 < nb) P::stage_redemod(K, rc, ln, ln.dm2, b2, dv.amp_magic, dv.amp_shift);
    if ((rc.rflags & RG_DROPOUT) && bE >= 0 && bE < nb) P::stage_dropout(ln, bE);
    for (int i = 0; i < K.recombine; i++) {                // -yc-recomb: rare, state lives in shared memory
        const int bm = bE - i, bd = bE - i - 1;
        if (bm >= 0 && bm < nb) P::template stage_modulate<false>(K, rc, ln, bm, false, nullptr);
        if (bd >= 0 && bd < nb) {
            Demod dm;
            dm.o1 = ln.rcomb[3 * i]; dm.o2 = ln.rcomb[3 * i + 1]; dm.o3 = ln.rcomb[3 * i + 2];
            P::stage_redemod(K, rc, ln, dm, bd, dv.amp_magic, dv.amp_shift);
            ln.rcomb[3 * i] = dm.o1; ln.rcomb[3 * i + 1] = dm.o2; ln.rcomb[3 * i + 2] = dm.o3;
        }
    }
    if (bF >= 0 && bF < nb && (K.flags & (G_OUT_FULL | G_OUT_LITE))) {
        const uint32_t uw = ldw(cblk(ln.ru, bF)), vw = ldw(cblk(ln.rv, bF));
        if (K.flags & G_OUT_FULL) {
            P::chroma_lp4(K, ln.ru, bF, uw, &ln.outU[0], &ln.outU[1], K.a_out[0], K.a_outhp[0], K.d_out[0]);
            P::chroma_lp4(K, ln.rv, bF, vw, &ln.outV[0], &ln.outV[1], K.a_out[1], K.a_outhp[1], K.d_out[1]);
        } else {
            P::chroma_lp4(K, ln.ru, bF, uw, nullptr, &ln.outU[1], K.a_out[0], 0.0, K.d_out[0]);
            P::chroma_lp4(K, ln.rv, bF, vw, nullptr, &ln.outV[1], K.a_out[1], 0.0, K.d_out[1]);
        }
    }
    if (bs < 0 || bs >= nb) return false;
    const uint8_t *yb = yblk(ln.ry, bs);
    out.y0 = ldw(yb);
    out.y1 = ldw(yb + 4);
    out.u = ldw(cblk(ln.ru, bs));
    out.v = ldw(cblk(ln.rv, bs));
    return true;
}

// ---- the fast kernel's steps ------------------------------------------------------------------------------
// Rows whose width is a multiple of 8 and that use the common switches only (fast_row_ok) never need the general
// variant: every block is whole, the line start is block 0 of the interior code (demod_w, the head-switch fill) and
// a role simply skips the steps in which its block does not exist.  One code path per role keeps what an SM
// executes inside its 32 KB instruction cache (scripts/probes/ifetch_probe.cu: 1.0 instruction per cycle and
// scheduler below that, 0.4 above) -- the general variant's line-end code alone is larger than that.
CVS_HD bool fast_row_ok(const K422 &K) { return !(K.flags & G_GENERAL) && (K.w % kB) == 0; }
CVS_HD bool blk_ok(int b, int nb) { return (unsigned)b < (unsigned)nb; }

CVS_HD void frow0_step(const K422 &K, const Lags &L, int nb, const Row422 &rc, Lane422 &ln, int s, const StepIO &in, bool warp_hs) {
    if (s <= nb) {                        // block nb: the two luma bytes past the row (:496)
        uint8_t *yb = yblk(ln.ry, s);
        stw(yb, in.y0);
        stw(yb + 4, in.y1);
        stw(cblk(ln.ru, s), in.u);
        stw(cblk(ln.rv, s), in.v);
    }
    if (s < nb && (K.flags & G_IN_LP)) {
        uint32_t ou, ov;
        Fast422::lp_pair<true>(in.u, in.v, &ln.inU[0], &ln.inU[1], &ln.inV[0], &ln.inV[1], K.a_in[0], K.a_inhp[0], K.a_in[1], K.a_inhp[1], ou, ov);
        Fast422::put_delayed(ln.ru, s, K.d_in[0], ou, ln.cIn[0]);
        Fast422::put_delayed(ln.rv, s, K.d_in[1], ov, ln.cIn[1]);
        if (s == nb - 1) {
            Fast422::carry_leave(ln.ru, s, K.d_in[0], ln.cIn[0]);
            Fast422::carry_leave(ln.rv, s, K.d_in[1], ln.cIn[1]);
        }
    }
    const int bM = s - L.bM;
    if (blk_ok(bM, nb)) Fast422::stage_modulate_first(K, rc, ln, bM, warp_hs);
#if CVS422_LSHARP_ROLE == 0
    if ((K.flags & G_VHS) && blk_ok(s - L.bC, nb)) Fast422::stage_luma_sharpen(K, ln, s - L.bC);
#endif
}
CVS_HD void frow1_step(const K422 &K, const Lags &L, int nb, const Row422 &rc, Lane422 &ln, int s) {
    const int b = s - L.bD;
    if (!blk_ok(b, nb)) return;
    Fast422::stage_separate(K, rc, ln, b);
}
CVS_HD void frow2_front(const K422 &K, const Lags &L, int nb, Lane422 &ln, int s, uint32_t &pu, uint32_t &pv) {
    pu = pv = 0;
    if (!(K.flags & G_VHS)) return;
    const int bC = s - L.bC, b = s - L.bV;
    if (blk_ok(bC, nb)) {
        Fast422::stage_chroma_lp(K, ln, bC);
        if (bC == nb - 1) {
            Fast422::carry_leave(ln.ru, bC, K.cd, ln.cCh[0]);
            Fast422::carry_leave(ln.rv, bC, K.cd, ln.cCh[1]);
        }
#if CVS422_LSHARP_ROLE == 2
        Fast422::stage_luma_sharpen(K, ln, bC);
#endif
    }
    if (blk_ok(b, nb)) Pipe422<false>::blend_fetch(ln, b, pu, pv);
}
CVS_HD void frow2_back(const K422 &K, const Lags &L, int nb, const Row422 &rc, Lane422 &ln, int s, uint32_t pu, uint32_t pv,
                       uint32_t au, uint32_t av) {
    const int b = s - L.bV;
    if (!(K.flags & G_VHS) || !blk_ok(b, nb)) return;
    Fast422::stage_vhs_chroma(K, rc, ln, b, pu, pv, au, av);
}
CVS_HD bool frow3_step(const K422 &K, const Lags &L, int nb, const Row422 &rc, Lane422 &ln, int s, StepIO &out, int &bs) {
    if ((K.flags & G_VHS) && !(K.flags & G_SVIDEO)) {
        const int bR = s - L.bR, b2 = s - L.bD2;
        if (blk_ok(bR, nb)) Fast422::stage_remodulate(rc, ln, bR);
        if (blk_ok(b2, nb)) Fast422::stage_redemod(rc, ln, b2);
    }
    const int bE = s - L.bE, bF = s - L.bF;
    if ((rc.rflags & RG_DROPOUT) && blk_ok(bE, nb)) Pipe422<false>::stage_dropout(ln, bE);
    if ((K.flags & (G_OUT_FULL | G_OUT_LITE)) && blk_ok(bF, nb)) {
        Fast422::stage_out(K, ln, bF);
        if (bF == nb - 1) {
            Fast422::carry_leave(ln.ru, bF, K.d_out[0], ln.cOut[0]);
            Fast422::carry_leave(ln.rv, bF, K.d_out[1], ln.cOut[1]);
        }
    }
    bs = s - L.bS;
    if (!blk_ok(bs, nb)) return false;
    const uint8_t *yb = yblk(ln.ry, bs);
    out.y0 = ldw(yb);
    out.y1 = ldw(yb + 4);
    out.u = ldw(cblk(ln.ru, bs));
    out.v = ldw(cblk(ln.rv, bs));
    return true;
}

// Head-switch pre-pass (:697-731) for rotations the in-ring delay cannot express (shift to the left, or
// further than the ring / the 10 % padding): the row's composite luma (G0..G2, with its exact noise) is
// computed up front and written, already rotated, into a scratch row: dest[(x - shif) mod twidth] =
// Y[x], everything not hit is the reference's padding value 16.
CVS_HD void headswitch_row(const K422 &K, const Row422 &rc_in, Lane422 &ln, const uint8_t *yrow, const uint8_t *urow,
                           const uint8_t *vrow, uint8_t *scratch, int shif) {
    typedef Pipe422<true> P;
    Row422 rc = rc_in;
    rc.rflags &= ~(uint32_t)(RG_HEADSW | RG_HEADSW_PRE);
    rc.hs_delay = 0;
    const Geo G = geo_of(K);
    const int w = K.w, tw = w + w / 10, nb = G.nb;
    for (int x = 0; x < w; x++) scratch[x] = 16;
    for (int s = 0; s <= nb; s++) {
        if (s < nb) {
            StepIO io;
            io.y0 = io.y1 = io.u = io.v = 0;
            for (int j = 0; j < kB; j++)
                if (s * kB + j < w) {
                    if (j < 4) io.y0 |= (uint32_t)yrow[s * kB + j] << (8 * j);
                    else io.y1 |= (uint32_t)yrow[s * kB + j] << (8 * (j - 4));
                }
            for (int k = 0; k < kBC; k++)
                if (s * kBC + k < K.cw) {
                    io.u |= (uint32_t)urow[s * kBC + k] << (8 * k);
                    io.v |= (uint32_t)vrow[s * kBC + k] << (8 * k);
                }
            P::stage_load(K, G, ln, s, io);
        }
        const int bM = s - 1;
        if (bM >= 0 && bM < nb) {
            P::template stage_modulate<true>(K, rc, ln, bM, false, nullptr);
            for (int j = 0; j < kB; j++) {
                const int x = bM * kB + j;
                if (x < w) {
                    int xd = (x - shif) % tw;
                    if (xd < 0) xd += tw;
                    if (xd < w) scratch[xd] = yblk(ln.ry, bM)[j];
                }
            }
        }
    }
}

}  // namespace cvs422
#endif
