// lane_pipeline.cuh -- the fused scanline pipeline, one lane = one field scanline.
//
// What it computes: every stage of the reference's composite_layer()
// (ffmpeg_ntsc.cpp:1570-1921) for ONE scanline, as a single streaming pass in x.
// The reference sweeps three int32 heap planes ~14 times per field; here a scanline never
// leaves registers between the BGRA load and the BGRA store (8 algorithmic bytes/pixel).
//
// Mapping: a warp owns 32 consecutive field rows (lane 0 is a halo row for
// the vertical chroma blend), all lanes advance through x in lock-step in blocks of kT (4 or 8)
// pixels, so every x-dependent condition (line start/end quirks) is warp-uniform, the
// vertical blend (ffmpeg_ntsc.cpp:1843-1863) is one __shfl_up per value, and the one-pole
// IIR cascades run in the reference's own sequential order (no re-association).  Stage
// look-ahead (QAM demodulation needs C[x+7], the VHS chroma delay is 9..14 samples, ...)
// is absorbed by running later stages on older blocks; all block lags are compile-time
// so every delay line is a statically indexed register array.
//
// The file is host/device portable on purpose: tests/ compile it with g++ and run all 32
// lanes of a warp in lock-step on the CPU (tests/emu_harness.cpp), where R = double gives
// the reference's arithmetic bit for bit and R = float is the production arithmetic.
//
// Timeline of step s (B(k) = pixels [kT k, kT k + kT); kLB = ceil(7/kT) blocks of demodulation
// look-ahead, LD = ceil(chroma delay/kT); in brackets the block for kT = 8 / kT = 4 with delay 9):
//   load  B(s)                BGRA block (prefetched one step ahead)
//   A1    B(s)                RGB->YIQ, input chroma lowpass cascades             (:1375, :1429)
//   A2    B(s-1)              composite C = Y + QAM(I',Q'), pre-emphasis, luma noise (:1460,:1613,:1631)
//   HS    B(s-1)              head switch: per-lane delay ring, or the pre-shifted scratch row (:1646)
//   B     B(kB), kB = s-1-kLB [s-2 / s-3]   Y/C separation + demod, chroma noise, phase noise (:1497,:1718,:1736)
//         VHS:                luma lowpass+HF boost, sharpen -> Y3 B(kB); chroma lowpass      (:1793-1883)
//   B2    B(kD), kD = kB-LD   [s-4 / s-6]   (VHS) delayed chroma block complete -> vertical blend -> remodulate C2
//   C     B(kC), kC = kD-kLB  [s-5 / s-8]   (VHS) second demod                               (:1885-1888)
//   F     B(kf)               dropout, output chroma lowpass, YIQ->RGB             (:1891-1916)
//                             kf = kC (VHS) / kD (VHS s-video) / kB (composite only)
//   store B(kf-1)
#ifndef CVS_LANE_PIPELINE_CUH
#define CVS_LANE_PIPELINE_CUH

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define CVS_HD __host__ __device__ __forceinline__
#define CVS_UNROLL _Pragma("unroll")
#else
#define CVS_HD inline __attribute__((always_inline))
#define CVS_UNROLL _Pragma("GCC unroll 16")
#endif

namespace cvs {

// Measured on B200 (1080p VHS-SP, 256 fields): kT = 8 / 2 CTAs per SM (243 registers) 2.55 ms;
// kT = 4 / 3 CTAs per SM (164 registers, no spills, 15 KB interior loop) 2.26 ms.
#ifndef CVS_KT
#define CVS_KT 4
#endif
#ifndef CVS_DEMOD_CARRY
#define CVS_DEMOD_CARRY 1        // interior steps reuse the overlapping half of the previous demodulation's box filter
#endif
// The conversion (XU) pipe runs at 1/8 of the FP32 rate, so the hot loop keeps clear of it (PRMT / directed-rounding
// tricks instead of I2F / F2I / FRND) -- except for a few conversions that take work off the FMA pipe, which is the
// busiest unit of this kernel.  Measured on B200, 1080p VHS-SP, fields/s (profiles/ab_variants_r1.txt):
//   round 1: unpack by PRMT+FADD 121.4 k; unpack by I2F.U8 120.8 k; + CVS_XU_TRUNC=1 126.2 k; CVS_XU_TRUNC=2 125.0 k
//   round 2 (profiles/ab_variants_r2.txt): TRUNC=1 141.1 k; TRUNC=3 145.0 k; + PAIR_TRUNC=1 148.7 k
#ifndef CVS_XU_UNPACK
#define CVS_XU_UNPACK 1          // BGRA bytes -> float by I2F.U8 (3 per pixel) instead of PRMT + FADD
#endif
#ifndef CVS_XU_TRUNC
#define CVS_XU_TRUNC 3           // 1: trunc() of a ready fp32 value by F2I.TRUNC + I2FP instead of LOP3 + FADD.RZ + FADD;
                                 // 2: also trunc(x * 2^k) (box filter / 4) by FMUL + F2I.TRUNC + I2FP;
                                 // 3: both by FRND.TRUNC, one conversion-pipe instruction (round 2: with ~20 % fewer
                                 //    instructions per step than in round 1 the conversion pipe has the room)
#endif
#ifndef CVS_XU_PAIR_TRUNC
#define CVS_XU_PAIR_TRUNC 1      // truncation of ready (I, Q) pairs by two FRND.TRUNC instead of the packed magic-number add
#endif
#ifndef CVS_RNG_PAIRED
#define CVS_RNG_PAIRED 0         // generator ring as 64-bit pairs (LaneRngP): half the LDS / STS, but the raw words of a
                                 // group are live together -- the register pressure costs more than the 12 instructions
                                 // save (profiles/ab_variants_r2.txt: VHS-EP 123 k against 138 k fields/s)
#endif
#ifndef CVS_MERGED_LUMA
#define CVS_MERGED_LUMA 1        // fp32 VHS luma chain on scaled states with merged boost / sharpen (Num<float>::vhs_luma)
#endif
#ifndef CVS_DIRECT_IQ
#define CVS_DIRECT_IQ 1          // fp32 RGB -> I, Q straight from the channels (three packed operations)
#endif
constexpr int kT = CVS_KT;       // pixels per step (8 or 4); every block lag below is derived from it
static_assert(kT == 4 || kT == 8, "kT must be 4 or 8");
constexpr int kLB = (7 + kT - 1) / kT;        // a demodulation of B(k) reads C up to x+7: kLB blocks of look-ahead
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr int lag_blocks(int px) { return (px + kT - 1) / kT; }
constexpr int kRngSlots = 32;    // per-lane ring of raw generator words (>= 31)
constexpr int kTailSlots = 16;   // per-lane stash of pre-filter chroma for the delay tail (>= max chroma delay)
constexpr uint32_t kRngBase = 128;  // ring index of the first in-line draw (multiple of 32, > max warm-up + 31)

// ---- flags (uniform per launch) ------------------------------------------------------------
enum : uint32_t {
    F_IN_LP = 1u << 0,        // composite_in_chroma_lowpass
    F_OUT_LP = 1u << 1,       // composite_out_chroma_lowpass
    F_PREEMPH = 1u << 2,      // composite_preemphasis != 0 && cut > 0
    F_NOCOLOR = 1u << 3,      // nocolor_subcarrier
    F_VBLEND = 1u << 4,       // vhs_chroma_vert_blend && output_ntsc
    F_SVIDEO = 1u << 5,       // vhs_svideo_out
    F_PHASE = 1u << 6,        // video_chroma_phase_noise != 0
    F_GENERAL = 1u << 7,      // any non-default switch: run the general (edge) variant everywhere
    F_NOISE_FAST = 1u << 8,   // per-pixel luma/chroma noise from a per-row counter generator instead of the exact
                              // rand() replay (cvs_set_noise_mode; the per-line draws stay exact)
};

// per-row flags (host side table)
enum : uint32_t {
    RF_DROPOUT = 1u << 0,     // chroma dropout hit this row (:1896)
    RF_HEADSW = 1u << 1,      // row is rotated by the head switch; C comes from the scratch row (pre-pass)
    RF_HEADSW_INLINE = 1u << 2,   // row is rotated by a small right shift with zero fill: delayed in-kernel
};
#ifndef CVS_HS_RING
#define CVS_HS_RING 44
#endif
constexpr int kHsRing = CVS_HS_RING;      // per-lane delay ring of the in-kernel head switch: a multiple of kT; not a power
                                          // of two on purpose (3 CTAs/SM must fit in shared memory).  The default 1080p
                                          // shift is -22 +-4 px of jitter; bigger shifts take the pre-pass.
constexpr int kHsSlots = kHsRing + kT;    // + one block: block 0 is stored twice so that a read of kT samples never wraps
constexpr int kHsMaxDelay = kHsRing - kT; // largest shift the ring can express

// ---- pairs --------------------------------------------------------------------------------------
// The two chroma planes (I/Q, later U/V) go through identical arithmetic, so they travel as a pair.
// On sm_100 an fp32 pair in an aligned register pair is one operand of the packed FFMA2/FADD2/FMUL2
// instructions (CUDA 12.9 __ffma2_rn/_rz/_rd, __fadd2_*, __fmul2_*): one issue slot does both planes.
// The kernel is issue-bound (profiles/ncu_kfields_r1.md), so this is where Blackwell's packed fp32
// pays.  Host code and the fp64 path see a plain struct and operate per component.
template <typename R> struct alignas(2 * sizeof(R)) V2 { R x, y; };
template <typename R> CVS_HD V2<R> mk2(R a, R b) { V2<R> v; v.x = a; v.y = b; return v; }

// ---- launch-uniform constants -------------------------------------------------------------
template <typename R>
struct KConst {
    // per filter: a = alpha, b = 1 - alpha, c = alpha^3 (gain of the scaled fp32 cascade, see Num<float>)
    V2<R> a_in, b_in, c_in;                       // input chroma lowpass: x = I 1.3 MHz, y = Q 0.6 MHz (:1442)
    R a_pre, b_pre, preemph;                      // composite pre-emphasis                    (:1621)
    R a_luma, b_luma, c_luma;                     // VHS luma lowpass                          (:1800)
    V2<R> a_ch, b_ch, c_ch;                       // VHS chroma lowpass (U and V alike)         (:1821)
    R a_sharp, b_sharp, c_sharp, sharpen;         // VHS sharpen lowpass (4x luma cut), gain   (:1874,:1880)
    // merged forms of the fp32 luma chain (Num<float>::vhs_luma): boost = 2.6 sv - 1.6 lp on the scaled states,
    // sharpen = (1 + 2g) y2 - 2g ts
    R k_boost_r, k_boost_p, k_sharp_y, k_sharp_t;
    // RGB -> I, Q straight from the channels (x 256): {r, g, b} coefficients of I (.x) and Q (.y), :1381-1382
    V2<R> k_iq_r, k_iq_g, k_iq_b;
    V2<R> a_out, b_out, c_out;                    // output chroma lowpass: x = I, y = Q        (:1411 / :1442)
    const R *phase_lut;                   // [2*pnoise+1][2] = {sin, cos} of state*pi/100, state = -p..p (:1746-1749)
    uint32_t flags;
    int32_t pnoise;                       // video_chroma_phase_noise
    int32_t phase_shift, phase_offset;    // video_scanline_phase_shift(_offset)
    int32_t amp, amp_back;                // subcarrier_amplitude, _back
    int32_t vnoise, cnoise;               // video_noise, video_chroma_noise
    uint32_t vmagic, vshift, cmagic, cshift;   // exact n % (2v+1) for 31-bit n: q = umulhi(n,magic) >> shift
    int32_t w, h;
};

// ---- numeric policy -------------------------------------------------------------------------
// R = double: the reference's arithmetic, operation for operation (three roundings per pole
// step, ffmpeg_ntsc.cpp:90-94; colour matrices as written).  Needs no FMA contraction
// (host: -ffp-contract=off; device: explicit __dmul_rn/__dadd_rn).
// R = float: production arithmetic; explicit fmaf where fused, everything else unfused, so the
// CPU emulation and the GPU produce identical bits.
template <typename R> struct Num;

template <> struct Num<double> {
    static CVS_HD double mul(double a, double b) {
#if defined(__CUDA_ARCH__)
        return __dmul_rn(a, b);
#else
        return a * b;
#endif
    }
    static CVS_HD double add(double a, double b) {
#if defined(__CUDA_ARCH__)
        return __dadd_rn(a, b);
#else
        return a + b;
#endif
    }
    static CVS_HD double sub(double a, double b) { return add(a, -b); }
    static CVS_HD double fma_(double a, double b, double c) { return add(mul(a, b), c); }   // never fused
    static CVS_HD double trunc_(double a) { return ::trunc(a); }
    static CVS_HD double floor_(double a) { return ::floor(a); }
    // lowpass(), :90-94
    static CVS_HD double pole(double &prev, double s, double alpha, double /*beta*/) {
        const double stage1 = mul(s, alpha);
        const double stage2 = sub(prev, mul(prev, alpha));
        prev = add(stage1, stage2);
        return prev;
    }
    // three cascaded poles; no truncation between stages (:1420/:1451/:1807/:1828/:1878)
    static CVS_HD void cascade_reset(double p[3], double v, double /*alpha*/) { p[0] = v; p[1] = v; p[2] = v; }
    static CVS_HD double cascade3(double p[3], double s, double a, double b, double /*c*/) {
        s = pole(p[0], s, a, b);
        s = pole(p[1], s, a, b);
        return pole(p[2], s, a, b);
    }
    static CVS_HD double cascade3_trunc(double p[3], double s, double a, double b, double c) {
        return trunc_(cascade3(p, s, a, b, c));
    }
    // pair forms: per component, in the reference's operation order
    typedef V2<double> P;
    static CVS_HD void cascade_reset2(P p[3], double v) { p[0] = mk2(v, v); p[1] = p[0]; p[2] = p[0]; }
    static CVS_HD P cascade3_trunc2(P p[3], P s, P a, P b, P /*c*/) {
        double x = pole(p[0].x, s.x, a.x, b.x); x = pole(p[1].x, x, a.x, b.x); x = pole(p[2].x, x, a.x, b.x);
        double y = pole(p[0].y, s.y, a.y, b.y); y = pole(p[1].y, y, a.y, b.y); y = pole(p[2].y, y, a.y, b.y);
        return mk2(trunc_(x), trunc_(y));
    }
    static CVS_HD P add2(P a, P b) { return mk2(add(a.x, b.x), add(a.y, b.y)); }
    static CVS_HD P scale2(P a, double k) { return mk2(mul(a.x, k), mul(a.y, k)); }
    static CVS_HD P axpy2(P a, double k, P c) { return mk2(add(mul(a.x, k), c.x), add(mul(a.y, k), c.y)); }   // a k + c
    static CVS_HD P floor_half2(P s) { return mk2(floor_half(s.x), floor_half(s.y)); }
    static CVS_HD P rot2(P uv, double c, double s, double /*ns*/) { return mk2(rot_a(uv.x, c, uv.y, s), rot_b(uv.x, s, uv.y, c)); }
    static CVS_HD void rgb2yiq2(uint32_t px, double &Y, P &IQ, const KConst<double> &) { rgb2yiq(px, Y, IQ.x, IQ.y); }
    static CVS_HD void rgb2yiq(uint32_t px, double &Y, double &I, double &Q, const KConst<double> &) { rgb2yiq(px, Y, I, Q); }
    static CVS_HD double floor_half(double s) { return ::floor(mul(s, 0.5)); }          // C int '>> 1'
    static CVS_HD double trunc_quarter(double s) { return ::trunc(mul(s, 0.25)); }     // C int '/ 4'
    static CVS_HD void unpack_rgb(uint32_t px, int &r, int &g, int &b) {
        r = (int)((px >> 16) & 0xFF); g = (int)((px >> 8) & 0xFF); b = (int)(px & 0xFF);
    }
    static CVS_HD void rgb2yiq(uint32_t px, double &Y, double &I, double &Q) {
        int r, g, b;
        unpack_rgb(px, r, g, b);
        rgb2yiq(r, g, b, Y, I, Q);
    }
    static CVS_HD uint32_t yiq2bgra(double Y, double I, double Q) {                    // :1385-1396, :1914
        const int r = clamp_(mix3(Y, 0.956, I, 0.621, Q));
        const int g = clamp_(mix3(Y, -0.272, I, -0.647, Q));
        const int b = clamp_(mix3(Y, -1.106, I, 1.703, Q));
        return ((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b;
    }
    static CVS_HD int clamp_(double v) { const int i = (int)v; return i < 0 ? 0 : (i > 255 ? 255 : i); }
    static CVS_HD void rgb2yiq(int r, int g, int b, double &Y, double &I, double &Q) {   // :1375-1383
        const double dY = add(add(mul(0.30, r), mul(0.59, g)), mul(0.11, b));
        Y = trunc_(mul(256, dY));
        I = trunc_(mul(256, add(mul(-0.27, sub(b, dY)), mul(0.74, sub(r, dY)))));
        Q = trunc_(mul(256, add(mul(0.41, sub(b, dY)), mul(0.48, sub(r, dY)))));
    }
    static CVS_HD double mix3(double Y, double ci, double I, double cq, double Q) {      // :1387-1389
        return trunc_(add(add(mul(1.000, Y), mul(ci, I)), mul(cq, Q)) / 256);
    }
    static CVS_HD double rot_a(double u, double c, double v, double s) { return trunc_(sub(mul(u, c), mul(v, s))); }  // :1756
    static CVS_HD double rot_b(double u, double s, double v, double c) { return trunc_(add(mul(u, s), mul(v, c))); }  // :1757
    static CVS_HD double boost(double s, double hp, double g) { return trunc_(add(s, mul(hp, g))); }                  // :1808-1810
    static CVS_HD double sharpen(double s, double ts, double g) { return trunc_(add(s, mul(mul(sub(s, ts), g), 2))); } // :1880
    static CVS_HD double preemph(double s, double hp, double g) { return trunc_(add(s, mul(hp, g))); }                // :1626-1627
    // VHS luma of one sample: three poles, + 1.6 x its own highpass (:1793-1812), then the sharpener (:1865-1883)
    static CVS_HD double pre_reset(double v, const KConst<double> &) { return v; }
    static CVS_HD double vhs_luma(double pL[3], double &pLpre, double pS[3], double y, const KConst<double> &K) {
        const double sv = cascade3(pL, y, K.a_luma, K.b_luma, K.c_luma);
        const double lp = pole(pLpre, sv, K.a_luma, K.b_luma);
        const double y2 = boost(sv, sub(sv, lp), 1.6);
        const double ts = cascade3(pS, y2, K.a_sharp, K.b_sharp, K.c_sharp);
        return sharpen(y2, ts, K.sharpen);
    }
};

// fp32 notes.
//  * Truncation / floor without the conversion (XU) pipe, which runs at 1/8 of the FP32 rate on
//    sm_100: for |x| < 2^22, x + 1.5*2^23 lands where the fp32 ulp is 1, so an add with directed
//    rounding IS the integer rounding:  floor(x) = (x +rd M) - M,  trunc(x) = (x +rz Ms) - Ms with
//    Ms = copysign(M, x).  The product/sum inside an FFMA is exact before that single rounding, so
//    trunc_mul(a, b) truncates the exact product; the host emulation reproduces it in double.
//  * Scaled cascade: with P_k = p_k / alpha^k the recurrence p_k' = beta p_k + alpha p_(k-1)' becomes
//    P_k' = beta P_k + P_(k-1)'  (one FFMA per pole) and the output is alpha^3 P_3' (one FMUL per
//    cascade instead of one per pole).  Same sequential order as the reference, no re-association.
template <> struct Num<float> {
    static constexpr float kMagic = 12582912.0f;        // 1.5 * 2^23
    static CVS_HD float mul(float a, float b) {
#if defined(__CUDA_ARCH__)
        return __fmul_rn(a, b);
#else
        return a * b;
#endif
    }
    static CVS_HD float add(float a, float b) {
#if defined(__CUDA_ARCH__)
        return __fadd_rn(a, b);
#else
        return a + b;
#endif
    }
    static CVS_HD float sub(float a, float b) { return add(a, -b); }
    static CVS_HD float fma_(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
        return __fmaf_rn(a, b, c);
#else
        return ::fmaf(a, b, c);
#endif
    }
    // trunc(x) for an fp32 value
    static CVS_HD float trunc_(float a) {
#if defined(__CUDA_ARCH__) && CVS_XU_TRUNC == 3
        return truncf(a);                                // FRND.TRUNC: one instruction on the conversion pipe
#elif defined(__CUDA_ARCH__) && CVS_XU_TRUNC
        return (float)__float2int_rz(a);                 // F2I.TRUNC (conversion pipe) + I2FP (ALU): |a| < 2^24
#elif defined(__CUDA_ARCH__)
        const float ms = __uint_as_float((__float_as_uint(a) & 0x80000000u) | 0x4B400000u);
        return __fadd_rn(__fadd_rz(a, ms), -ms);
#else
        return ::truncf(a);
#endif
    }
    // trunc(a * b) of the EXACT product, b > 0
    static CVS_HD float trunc_mul_pos(float a, float b) {
#if defined(__CUDA_ARCH__)
        const float ms = __uint_as_float((__float_as_uint(a) & 0x80000000u) | 0x4B400000u);
        return __fadd_rn(__fmaf_rz(a, b, ms), -ms);
#else
        return (float)::trunc((double)a * (double)b);
#endif
    }
    static CVS_HD float floor_(float a) {
#if defined(__CUDA_ARCH__)
        return __fadd_rn(__fadd_rd(a, kMagic), -kMagic);
#else
        return ::floorf(a);
#endif
    }
    // trunc(a * b) of the exact product for a, b >= 0: truncation is floor, no sign to look at
    static CVS_HD float trunc_mul_nonneg(float a, float b) {
#if defined(__CUDA_ARCH__)
        return __fadd_rn(__fmaf_rd(a, b, kMagic), -kMagic);
#else
        return (float)::trunc((double)a * (double)b);
#endif
    }
    static CVS_HD float floor_half(float s) {            // floor(s / 2), C int '>> 1'
#if defined(__CUDA_ARCH__)
        return __fadd_rn(__fmaf_rd(s, 0.5f, kMagic), -kMagic);
#else
        return ::floorf(s * 0.5f);
#endif
    }
    // trunc(a * b) for b a power of two (the product is exact in fp32)
    static CVS_HD float trunc_pow2(float a, float b) {
#if defined(__CUDA_ARCH__) && CVS_XU_TRUNC == 3
        return truncf(__fmul_rn(a, b));                   // FMUL + FRND.TRUNC
#elif defined(__CUDA_ARCH__) && CVS_XU_TRUNC >= 2
        return (float)__float2int_rz(__fmul_rn(a, b));    // FMUL + F2I.TRUNC + I2FP: one FMA-pipe slot instead of two
#else
        return trunc_mul_pos(a, b);
#endif
    }
    static CVS_HD float trunc_quarter(float s) { return trunc_pow2(s, 0.25f); }   // C int '/ 4'

    static CVS_HD float pole(float &prev, float s, float alpha, float beta) {
        prev = fma_(beta, prev, mul(alpha, s));
        return prev;
    }
    static CVS_HD void cascade_reset(float p[3], float v, float alpha) {
        p[0] = v / alpha;
        p[1] = p[0] / alpha;
        p[2] = p[1] / alpha;
    }
    static CVS_HD float cascade3_raw(float p[3], float s, float b) {
        p[0] = fma_(b, p[0], s);
        p[1] = fma_(b, p[1], p[0]);
        p[2] = fma_(b, p[2], p[1]);
        return p[2];
    }
    static CVS_HD float cascade3(float p[3], float s, float /*a*/, float b, float c) { return mul(cascade3_raw(p, s, b), c); }
    static CVS_HD float cascade3_trunc(float p[3], float s, float /*a*/, float b, float c) {
        return trunc_mul_pos(cascade3_raw(p, s, b), c);
    }
    // ---- pair forms: packed FFMA2/FADD2/FMUL2 on the device, the same per-component arithmetic on the host
    typedef V2<float> P;
#if defined(__CUDA_ARCH__)
    static __device__ __forceinline__ float2 f2(P a) { return make_float2(a.x, a.y); }
    static __device__ __forceinline__ P pp(float2 a) { return mk2(a.x, a.y); }
    static __device__ __forceinline__ float2 msign2(float2 a) {      // copysign(1.5*2^23, a) per component
        return make_float2(__uint_as_float((__float_as_uint(a.x) & 0x80000000u) | 0x4B400000u),
                           __uint_as_float((__float_as_uint(a.y) & 0x80000000u) | 0x4B400000u));
    }
#endif
#if defined(__CUDA_ARCH__)
    // trunc of both components of a ready pair: packed directed-rounding add (4 instructions, FMA + ALU pipes), or
    // two FRND.TRUNC on the conversion pipe (CVS_XU_PAIR_TRUNC: A/B switch)
    static __device__ __forceinline__ float2 trunc2_ready(float2 r) {
#if CVS_XU_PAIR_TRUNC
        return make_float2(truncf(r.x), truncf(r.y));
#else
        const float2 ms = msign2(r);
        return __fadd2_rn(__fadd2_rz(r, ms), make_float2(-ms.x, -ms.y));
#endif
    }
#endif
    static CVS_HD void cascade_reset2(P p[3], float v) { (void)v; p[0] = mk2(0.0f, 0.0f); p[1] = p[0]; p[2] = p[0]; }   // chroma resets to 0
    static CVS_HD P cascade3_trunc2(P p[3], P s, P /*a*/, P b, P c) {
#if defined(__CUDA_ARCH__)
        const float2 bb = f2(b);
        float2 q0 = __ffma2_rn(bb, f2(p[0]), f2(s));
        float2 q1 = __ffma2_rn(bb, f2(p[1]), q0);
        float2 q2 = __ffma2_rn(bb, f2(p[2]), q1);
        p[0] = pp(q0); p[1] = pp(q1); p[2] = pp(q2);
        // (FMUL2 + two FRND.TRUNC here as well measured slower: 145.6 k against 148.7 k fields/s, profiles/ab_variants_r2.txt)
        const float2 ms = msign2(q2);                               // sign(q2 * c) == sign(q2): c > 0
        return pp(__fadd2_rn(__ffma2_rz(q2, f2(c), ms), make_float2(-ms.x, -ms.y)));
#else
        float px_[3] = {p[0].x, p[1].x, p[2].x}, py_[3] = {p[0].y, p[1].y, p[2].y};
        const float x = trunc_mul_pos(cascade3_raw(px_, s.x, b.x), c.x);
        const float y = trunc_mul_pos(cascade3_raw(py_, s.y, b.y), c.y);
        for (int k = 0; k < 3; k++) p[k] = mk2(px_[k], py_[k]);
        return mk2(x, y);
#endif
    }
    static CVS_HD P add2(P a, P b) {
#if defined(__CUDA_ARCH__)
        return pp(__fadd2_rn(f2(a), f2(b)));
#else
        return mk2(a.x + b.x, a.y + b.y);
#endif
    }
    static CVS_HD P scale2(P a, float k) {                          // both components times k
#if defined(__CUDA_ARCH__)
        return pp(__fmul2_rn(f2(a), make_float2(k, k)));
#else
        return mk2(a.x * k, a.y * k);
#endif
    }
    static CVS_HD P axpy2(P a, float k, P c) {                      // a k + c per component (one rounding)
#if defined(__CUDA_ARCH__)
        return pp(__ffma2_rn(f2(a), make_float2(k, k), f2(c)));
#else
        return mk2(::fmaf(a.x, k, c.x), ::fmaf(a.y, k, c.y));
#endif
    }
    static CVS_HD P floor_half2(P s) {
#if defined(__CUDA_ARCH__)
        const float2 m = make_float2(kMagic, kMagic);
        return pp(__fadd2_rn(__ffma2_rd(f2(s), make_float2(0.5f, 0.5f), m), make_float2(-kMagic, -kMagic)));
#else
        return mk2(floor_half(s.x), floor_half(s.y));
#endif
    }
    // (u, v) -> trunc(u c - v s, u s + v c); ns = -s  (:1756-1757)
    static CVS_HD P rot2(P uv, float c, float s, float ns) {
#if defined(__CUDA_ARCH__)
        const float2 t = __fmul2_rn(make_float2(uv.y, uv.y), make_float2(ns, c));
        const float2 r = __ffma2_rn(make_float2(uv.x, uv.x), make_float2(c, s), t);
        return pp(trunc2_ready(r));
#else
        (void)ns;
        return mk2(rot_a(uv.x, c, uv.y, s), rot_b(uv.x, s, uv.y, c));
#endif
    }
    // RGB_to_YIQ (:1375-1383).  Y = trunc(256 dY) with dY >= 0; I and Q are linear in (r, g, b), so they come
    // straight from the channels with the x256 folded into the coefficients (K.k_iq_*: three packed operations
    // for both planes instead of forming b - dY and r - dY first).
    static CVS_HD void rgb2yiq2(uint32_t px, float &Y, P &IQ, const KConst<float> &K) {
#if defined(__CUDA_ARCH__)
#if CVS_XU_UNPACK
        // I2F.U8 with a byte selector: one instruction per channel on the (otherwise idle) conversion pipe
        const float rf = (float)((px >> 16) & 0xFFu), gf = (float)((px >> 8) & 0xFFu), bf = (float)(px & 0xFFu);
#else
        const float rf = __fadd_rn(__uint_as_float(__byte_perm(px, 0x4B000000u, 0x7442)), -8388608.0f);
        const float gf = __fadd_rn(__uint_as_float(__byte_perm(px, 0x4B000000u, 0x7441)), -8388608.0f);
        const float bf = __fadd_rn(__uint_as_float(__byte_perm(px, 0x4B000000u, 0x7440)), -8388608.0f);
#endif
        const float dY = fma_(0.11f, bf, fma_(0.59f, gf, mul(0.30f, rf)));
        Y = trunc_mul_nonneg(dY, 256.0f);
#if CVS_DIRECT_IQ
        float2 iq = __fmul2_rn(f2(K.k_iq_r), make_float2(rf, rf));
        iq = __ffma2_rn(f2(K.k_iq_g), make_float2(gf, gf), iq);
        iq = __ffma2_rn(f2(K.k_iq_b), make_float2(bf, bf), iq);
        IQ = pp(trunc2_ready(iq));
#else
        const float bd = sub(bf, dY), rd = sub(rf, dY);
        const float2 t = __fmul2_rn(make_float2(-0.27f, 0.41f), make_float2(bd, bd));
        const float2 iq = __ffma2_rn(make_float2(0.74f, 0.48f), make_float2(rd, rd), t);
        const float2 ms = msign2(iq);
        IQ = pp(__fadd2_rn(__ffma2_rz(iq, make_float2(256.0f, 256.0f), ms), make_float2(-ms.x, -ms.y)));
#endif
#else
        rgb2yiq(px, Y, IQ.x, IQ.y, K);
#endif
    }
    static CVS_HD void rgb2yiq(uint32_t px, float &Y, float &I, float &Q, const KConst<float> &K) {
#if defined(__CUDA_ARCH__)
        // 0x4B0000bb is the float 2^23 + bb: one PRMT + one FADD per channel instead of an I2F
        const float rf = __fadd_rn(__uint_as_float(__byte_perm(px, 0x4B000000u, 0x7442)), -8388608.0f);
        const float gf = __fadd_rn(__uint_as_float(__byte_perm(px, 0x4B000000u, 0x7441)), -8388608.0f);
        const float bf = __fadd_rn(__uint_as_float(__byte_perm(px, 0x4B000000u, 0x7440)), -8388608.0f);
#else
        const float rf = (float)((px >> 16) & 0xFF), gf = (float)((px >> 8) & 0xFF), bf = (float)(px & 0xFF);
#endif
        const float dY = fma_(0.11f, bf, fma_(0.59f, gf, mul(0.30f, rf)));
        Y = trunc_mul_nonneg(dY, 256.0f);
#if CVS_DIRECT_IQ
        I = trunc_(fma_(K.k_iq_b.x, bf, fma_(K.k_iq_g.x, gf, mul(K.k_iq_r.x, rf))));
        Q = trunc_(fma_(K.k_iq_b.y, bf, fma_(K.k_iq_g.y, gf, mul(K.k_iq_r.y, rf))));
#else
        const float bd = sub(bf, dY), rd = sub(rf, dY);
        I = trunc_mul_pos(fma_(0.74f, rd, mul(-0.27f, bd)), 256.0f);
        Q = trunc_mul_pos(fma_(0.48f, rd, mul(0.41f, bd)), 256.0f);
#endif
    }
    // YIQ_to_RGB (:1385-1396) + the store's packing (:1914): trunc((Y + ci I + cq Q) / 256) clamped to 0..255 per
    // channel, alpha 0.  The operands carry the 1/256 (exact: a power of two), so each channel is two FFMAs, and
    // on sm_100 the truncation, the clamp AND the byte packing are the packed conversion F2IP.U8.F32.TRUNC
    // (PTX cvt.rzi.u8.f32 / cvt.pack.sat.u8.s32.b32 on a float-to-int pair): two instructions per pixel where the
    // round-1 kernel needed a saturating FFMA, an FFMA.RZ, an integer min per channel and two PRMTs.
    static CVS_HD uint32_t yiq2bgra(float Y, float I, float Q) {
        constexpr float k = 1.0f / 256.0f;
        const float Yp = mul(Y, k);
        const float r = fma_(0.621f * k, Q, fma_(0.956f * k, I, Yp));
        const float g = fma_(-0.647f * k, Q, fma_(-0.272f * k, I, Yp));
        const float b = fma_(1.703f * k, Q, fma_(-1.106f * k, I, Yp));
#if defined(__CUDA_ARCH__)
        uint32_t hi, px;
        asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(hi) : "f"(r));                       // 0x000000RR
        asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;"                             // (hi << 16) | G << 8 | B
            : "=r"(px) : "r"(__float2int_rz(g)), "r"(__float2int_rz(b)), "r"(hi));
        return px;
#else
        return ((uint32_t)chan8(r) << 16) | ((uint32_t)chan8(g) << 8) | (uint32_t)chan8(b);
#endif
    }
    static CVS_HD int chan8(float v) { return (v < 0.0f) ? 0 : (v > 255.0f ? 255 : (int)v); }
    static CVS_HD float rot_a(float u, float c, float v, float s) { return trunc_(fma_(u, c, -mul(v, s))); }
    static CVS_HD float rot_b(float u, float s, float v, float c) { return trunc_(fma_(u, s, mul(v, c))); }
    static CVS_HD float boost(float s, float hp, float g) { return trunc_(fma_(hp, g, s)); }
    static CVS_HD float sharpen(float s, float ts, float g) { return trunc_(fma_(mul(sub(s, ts), g), 2.0f, s)); }
    static CVS_HD float preemph(float s, float hp, float g) { return trunc_(fma_(hp, g, s)); }
    // VHS luma of one sample on scaled states (see the fp32 notes): r = lowpassed luma / alpha^3; the pre-emphasis
    // pole keeps lp / alpha^4, so its update is one FFMA on r; the boost sv + 1.6 (sv - lp) and the sharpener
    // y2 + 2g (y2 - ts) are each one FMUL + one FFMA on those states (8 + 7 instructions instead of 10 + 9).
#if !CVS_MERGED_LUMA
    static CVS_HD float pre_reset(float v, const KConst<float> &) { return v; }
    static CVS_HD float vhs_luma(float pL[3], float &pLpre, float pS[3], float y, const KConst<float> &K) {
        const float sv = cascade3(pL, y, K.a_luma, K.b_luma, K.c_luma);
        const float lp = pole(pLpre, sv, K.a_luma, K.b_luma);
        const float y2 = boost(sv, sub(sv, lp), 1.6f);
        const float ts = cascade3(pS, y2, K.a_sharp, K.b_sharp, K.c_sharp);
        return sharpen(y2, ts, K.sharpen);
    }
#else
    static CVS_HD float pre_reset(float v, const KConst<float> &K) { return v / (K.a_luma * K.c_luma); }
    static CVS_HD float vhs_luma(float pL[3], float &pLpre, float pS[3], float y, const KConst<float> &K) {
        const float r = cascade3_raw(pL, y, K.b_luma);
        pLpre = fma_(K.b_luma, pLpre, r);
        const float y2 = trunc_(fma_(pLpre, K.k_boost_p, mul(r, K.k_boost_r)));
        const float t = cascade3_raw(pS, y2, K.b_sharp);
        return trunc_(fma_(t, K.k_sharp_t, mul(y2, K.k_sharp_y)));
    }
#endif
};

// integer helpers on integer-valued reals (all plane values are integers, |v| < 2^24)
template <typename R> CVS_HD R div4_trunc(R s) { return Num<R>::trunc_quarter(s); }   // C int '/ 4'
template <typename R> CVS_HD R shr1_floor(R s) { return Num<R>::floor_half(s); }      // C int '>> 1'


// (v * num) / den on C ints (truncation toward zero); |v*num| < 2^31 by construction
CVS_HD int muldiv_trunc(int v, int num, int den) { return (v * num) / den; }

// ---- per-lane RNG + noise ---------------------------------------------------------------------
// Ring layout in shared memory: word (slot, lane) at ring[slot * stride + lane]; stride = number
// of lanes sharing the buffer, so a warp access is conflict-free.  On the host stride = 1.
struct LaneRng {
    uint32_t *ring;      // already offset by the lane index
    int stride;
    uint32_t l1, l2, l3; // q[n-1], q[n-2], q[n-3] (short-lag taps stay in registers)

    // hist[i] = q[n0 - 31 + i], i = 0..30: the 31 raw words preceding ring index n0
    CVS_HD void init(uint32_t *ring_, int stride_, const uint32_t hist[31], uint32_t n0) {
        ring = ring_;
        stride = stride_;
        for (int i = 0; i < 31; i++) ring[((n0 - 31 + i) & (kRngSlots - 1)) * stride] = hist[i];
        l1 = hist[30]; l2 = hist[29]; l3 = hist[28];
    }
    CVS_HD void init(uint32_t *base, int lane, int stride_, const uint32_t hist[31], uint32_t n0) { init(base + lane, stride_, hist, n0); }
    // q[n] = q[n-31] + q[n-3].  The caller supplies n (same for every lane of a warp, so the slot
    // arithmetic is warp-uniform); draws must be requested in increasing n without gaps.
    CVS_HD uint32_t next_raw(uint32_t n) {
        const uint32_t v = ring[((n + 1) & (kRngSlots - 1)) * stride] + l3;     // (n-31) & 31 == (n+1) & 31
        ring[(n & (kRngSlots - 1)) * stride] = v;
        l3 = l2; l2 = l1; l1 = v;
        return v;
    }
    // Group form used by the pipeline: a step draws `count` consecutive words starting at a ring index
    // that is a multiple of `count` (8 luma / 16 chroma draws per 8-pixel step, kRngBase % 32 == 0), so
    // the group never wraps: word i is written at grp + i*stride and its long-lag tap q[n-31] sits one
    // slot further, i.e. at grp + (i+1)*stride, or at the start of the NEXT group for the last word.
    // grp/grp_next are computed once per step; every access is base + compile-time offset.
    CVS_HD uint32_t *group_ptr(uint32_t n_first) const { return ring + (n_first & (kRngSlots - 1)) * stride; }
    CVS_HD uint32_t next_in_group(uint32_t *grp, const uint32_t *grp_next, int i, int count) {
        const uint32_t v = ((i + 1 < count) ? grp[(i + 1) * stride] : grp_next[0]) + l3;
        grp[i * stride] = v;
        l3 = l2; l2 = l1; l1 = v;
        return v;
    }
};

// The same generator with the ring stored as PAIRS of consecutive words (k_fields): pair p of a lane is the 64-bit
// word at ((p * stride) + lane) * 2, so a warp's 64-bit access covers 256 contiguous bytes (conflict-free) and a
// group of COUNT consecutive draws costs COUNT/2 loads and COUNT/2 stores instead of COUNT each.  The long-lag tap
// of draw n is slot (n + 1) & 31, i.e. the taps of an aligned group straddle the pair grid by one word: the first
// tap is the high word of the group's own first pair -- carried over from the previous group's last load -- and the
// last tap is the low word of the NEXT group's first pair, whose high word becomes the new carry.
struct LaneRngP {
    uint32_t *ring;      // already offset by 2 * lane words
    int stride;          // lanes sharing the buffer
    uint32_t l1, l2, l3; // q[n-1], q[n-2], q[n-3]
    uint32_t carry;      // the word at slot (n + 1) & 31 for the next group's first draw n

    CVS_HD uint32_t &word(uint32_t slot) const { return ring[(size_t)((slot & (kRngSlots - 1)) >> 1) * 2 * stride + (slot & 1)]; }
    CVS_HD void init(uint32_t *base, int lane, int stride_, const uint32_t hist[31], uint32_t n0) {
        ring = base + 2 * lane;
        stride = stride_;
        for (int i = 0; i < 31; i++) word(n0 - 31 + (uint32_t)i) = hist[i];
        l1 = hist[30]; l2 = hist[29]; l3 = hist[28];
        carry = 0;
    }
    CVS_HD uint32_t next_raw(uint32_t n) {               // one draw, any n (warm-up)
        const uint32_t v = word(n + 1) + l3;
        word(n) = v;
        l3 = l2; l2 = l1; l1 = v;
        return v;
    }
    CVS_HD void begin_groups(uint32_t n_first) { carry = word(n_first + 1); }
    // COUNT consecutive draws from n_first (a multiple of COUNT; groups must be drawn in order without gaps)
    template <int COUNT>
    CVS_HD void draw_group(uint32_t n_first, uint32_t out[COUNT]) {
        static_assert(COUNT % 2 == 0 && kRngSlots % COUNT == 0, "aligned groups of pairs");
        uint32_t *grp = ring + (size_t)((n_first & (kRngSlots - 1)) >> 1) * 2 * stride;
        const uint32_t *nxt = ring + (size_t)(((n_first + COUNT) & (kRngSlots - 1)) >> 1) * 2 * stride;
        uint32_t tap[COUNT];
        tap[0] = carry;
        CVS_UNROLL
        for (int p = 1; p < COUNT / 2; p++) ld2(grp + (size_t)p * 2 * stride, tap[2 * p - 1], tap[2 * p]);
        ld2(nxt, tap[COUNT - 1], carry);
        CVS_UNROLL
        for (int i = 0; i < COUNT; i++) {
            const uint32_t v = tap[i] + l3;
            l3 = l2; l2 = l1; l1 = v;
            out[i] = v;
        }
        CVS_UNROLL
        for (int p = 0; p < COUNT / 2; p++) st2(grp + (size_t)p * 2 * stride, out[2 * p], out[2 * p + 1]);
    }
    static CVS_HD void ld2(const uint32_t *p, uint32_t &a, uint32_t &b) {
#if defined(__CUDA_ARCH__)
        const uint2 t = *reinterpret_cast<const uint2 *>(p);
        a = t.x; b = t.y;
#else
        a = p[0]; b = p[1];
#endif
    }
    static CVS_HD void st2(uint32_t *p, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<uint2 *>(p) = make_uint2(a, b);
#else
        p[0] = a; p[1] = b;
#endif
    }
};

// n % m for n < 2^31 and 1 < m < 2^16 via q = umulhi(n, magic) >> shift (exact)
CVS_HD uint32_t umulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
CVS_HD int draw_mod(uint32_t raw, uint32_t m, uint32_t magic, uint32_t shift) {
    const uint32_t n = raw >> 1;                       // rand() = q >> 1
    const uint32_t q = umulhi32(n, magic) >> shift;
    return (int)(n - q * m);
}
// noise += rand() % (2v+1) - v; noise /= 2   (:1640-1641, :1729-1732)
CVS_HD int noise_step(int noise, int d, int v) {
    int t = noise + d - v;
    t += (int)((uint32_t)t >> 31);                     // C '/ 2' truncates toward zero
    return t >> 1;
}
// draw + update in one expression: noise + n - (q m + v) is one IADD3 fed by one IMAD on the device
CVS_HD int noise_draw(int noise, uint32_t raw, uint32_t m, uint32_t magic, uint32_t shift, int v) {
    const uint32_t n = raw >> 1;                       // rand() = q >> 1
    const uint32_t q = umulhi32(n, magic) >> shift;
    int t = noise + (int)n - (int)(q * m + (uint32_t)v);
    t += (int)((uint32_t)t >> 31);
    return t >> 1;
}

// "Fast" per-pixel noise (the per-pixel noise amplitudes are sub-LSB in the x256 domain, so any
// generator stays within +-1 LSB of the reference; the per-LINE draws are never substituted).  One 32-bit LCG per
// lane and stream, seeded per (field, row, stream); a draw is uniform in [0, m): umulhi(state, m).
CVS_HD uint32_t lcg_seed(unsigned long long fieldno, unsigned field, int row, uint32_t stream) {
    uint32_t h = (uint32_t)fieldno * 0x9E3779B1u + (uint32_t)(fieldno >> 32) * 0x7F4A7C15u;
    h ^= ((uint32_t)row * 2u + field) * 0x85EBCA77u + stream * 0xC2B2AE3Du;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
CVS_HD int noise_draw_fast(int noise, uint32_t &state, uint32_t m, int v) {
    state = state * 1664525u + 1013904223u;
    int t = noise + (int)umulhi32(state, m) - v;
    t += (int)((uint32_t)t >> 31);
    return t >> 1;
}

// ---- per-row constants --------------------------------------------------------------------------
template <typename R>
struct RowConst {
    R sM;                // QAM carrier sign of x = 0 (mod 4): +1 when (xi & 2) == 0 (Umult/Vmult = {1,0,-1,0}/{0,1,0,-1}
                         // rotated by the line phase xi, :1465-1466); the taps of x & 3 = j follow from xi and sM
    R sgA;               // demod sign for even x with (x & 2) == 0; the (x & 2) != 0 sign is -sgA
    R sinp, cosp, nsinp; // chroma phase noise rotation of this row (:1748-1749); nsinp = -sinp
    int xi;              // subcarrier phase index (:1473-1480)
    uint32_t rflags;
    int row;             // field row index (y = field + 2*row)
    int hs_delay;        // RF_HEADSW_INLINE: C'[x] = C[x - hs_delay], 0 for x < hs_delay
    R bl_above, bl_own;  // vertical chroma blend as (above * bl_above + own * bl_own + 1) >> 1 (:1843-1863): (1, 1) for
                         // rows >= field+4, (0, 1) for row field+2 (blends with the zeroed delay line), and (0, 2) where
                         // nothing is blended (row `field`, or the blend is off): (2 own + 1) >> 1 == own
    bool odd_any;        // warp-uniform: some row of this warp has an odd phase index (set by the caller; default true)
};

template <typename R>
CVS_HD void rowconst_set_phase(RowConst<R> &rc, int xi) {
    rc.xi = xi;
    rc.sM = (R)((xi & 2) ? -1 : 1);
    // I[x] = -flip(x+xi) * chroma[x+xi]; for even x with (x&3)==0 the carrier is flipped iff xi is odd
    rc.sgA = (R)((xi & 1) ? 1 : -1);
}

// ---- the lane ----------------------------------------------------------------------------------
template <typename R, bool VHS, int CD, bool OUTFULL>
struct Lane {
    static constexpr int OD = OUTFULL ? 4 : 1;    // output lowpass: max(delayI, delayQ)
    static constexpr int ODI = OUTFULL ? 2 : 1;   // I delay
    static constexpr int ODQ = OUTFULL ? 4 : 1;   // Q delay
    static constexpr int LD = lag_blocks(CD);     // blocks by which the delayed (VHS) chroma lags its filter input
    static constexpr int LAG = VHS ? 2 + 2 * kLB + LD : 2 + kLB;   // steps between loading B(s) and storing B(s-LAG)
    static_assert(OD <= kT, "output lowpass delay must fit one block");

    // --- carried state (all statically indexed) ---
    // A
    R Yprev[kT];                     // Y of B(s-1)
    V2<R> oIQprev[kT];               // input-lowpass cascade outputs (I, Q) for t in B(s-1)
    V2<R> pIQ[3];                    // cascade poles
    R pPre;                          // pre-emphasis pole
    int nY;                          // luma noise accumulator (:1633)
    // B
    R Cm1;                           // C[kT*kB - 1], kB = s-1-kLB
    R Cwin[kLB][kT];                 // C of B(kB) .. B(s-2)
    R chc1[5], chc2[5];              // chroma residues carried between consecutive demodulations (demod_block)
    int nU, nV;                      // chroma noise accumulators (:1720)
    R pL[3], pLpre;                  // VHS luma poles (:1800-1805)
    R pS[3];                         // sharpen poles (:1873-1876)
    V2<R> pUV[3];                    // VHS chroma poles (:1821-1826)
    V2<R> oUVprev[kT];               // chroma cascade outputs (U, V) for t in B(s-3)
    R Y3hist[LD][kT];                // Y3 of B(kB-LD) .. B(kB-1)   ([0] is the oldest)
    // B2 / C
    R C2m1;                          // C2[kT*kC - 1], kC = kD - kLB, kD = kB - LD
    R C2win[kLB][kT];                // C2 of B(kC) .. B(kD-1)
    // F
    V2<R> pOIQ[3];                   // output lowpass poles
    R Ytail[OD];                     // last OD values of Y4 of the previous F block
    V2<R> IQtail[OD], oOIQtail[OD];  // ... of the raw (I4, Q4) and of the cascade outputs
    uint32_t outprev[kT - OD > 0 ? kT - OD : 1];   // packed pixels of positions [kT(k-1), kT k - OD)

#if CVS_RNG_PAIRED
    typedef LaneRngP Gen;
#else
    typedef LaneRng Gen;
#endif
    Gen rngL, rngC;
    // (fast noise mode keeps its two generator states in rngL.l1 / rngC.l1: the exact generators are idle then)
    R *tailU, *tailV;                // per-lane stash (kTailSlots each), stride tail_stride
    int tail_stride;

    CVS_HD void reset(const KConst<R> &K) {
        CVS_UNROLL
        for (int j = 0; j < kT; j++) {
            Yprev[j] = 0; oIQprev[j] = mk2((R)0, (R)0);
            oUVprev[j] = mk2((R)0, (R)0);
            for (int i = 0; i < kLB; i++) { Cwin[i][j] = 0; C2win[i][j] = 0; }
            for (int i = 0; i < LD; i++) Y3hist[i][j] = 0;
        }
        CVS_UNROLL
        for (int k = 0; k < 3; k++) pS[k] = 0;    // :1875
        Num<R>::cascade_reset2(pIQ, (R)0);        // resetFilter(0), :1447
        Num<R>::cascade_reset2(pUV, (R)0);        // :1823,:1825
        Num<R>::cascade_reset2(pOIQ, (R)0);       // :1416
        Num<R>::cascade_reset(pL, (R)16, K.a_luma);   // resetFilter(16), :1802
        pLpre = Num<R>::pre_reset((R)16, K);    // :1805
        pPre = 16;                          // :1622
        Cm1 = 0; C2m1 = 0;
        for (int m = 0; m < 5; m++) { chc1[m] = 0; chc2[m] = 0; }
        for (int j = 0; j < (kT - OD > 0 ? kT - OD : 1); j++) outprev[j] = 0;
        CVS_UNROLL
        for (int k = 0; k < OD; k++) { Ytail[k] = 0; IQtail[k] = mk2((R)0, (R)0); oOIQtail[k] = mk2((R)0, (R)0); }
    }
};

// QAM modulation of one sample (chroma_into_luma, :1486-1490): Y += (I amp Umult[ph] + Q amp Vmult[ph]) / 50 with
// ph = (xi + x) & 3, Umult = {1,0,-1,0}, Vmult = {0,1,0,-1}: exactly one of the two taps is +-1, so the sample is
// Y + sg * (ph odd ? Q : I) with sg = -1 for ph >= 2 (exact; amp == 50 makes (v*50)/50 == v).
// MODE_FAST_EVEN: the whole warp has an even line phase (the default 180-degree phase), so the I/Q choice is
// the compile-time parity of j and the sign is the lane's rc.sM: one FFMA per sample, no per-row tap table.
template <typename R, int MODE>
CVS_HD R modulate(const RowConst<R> &rc, int j, R Yv, R Iv, R Qv, int amp) {
    if (MODE < 0) {                                   // MODE_FAST_EVEN
        const R v = (j & 1) ? Qv : Iv;
        return Num<R>::fma_(v, (j & 2) ? -rc.sM : rc.sM, Yv);
    }
    const int ph = (rc.xi + j) & 3;
    const R v = (ph & 1) ? Qv : Iv;
    const R sg = (ph & 2) ? (R)-1 : (R)1;
    if (MODE != 2 || amp == 50) return Num<R>::fma_(v, sg, Yv);
    const int chroma = (int)v * amp * (int)sg;
    return Num<R>::add(Yv, (R)(chroma / 50));
}

// Y/C separation + QAM demodulation of block B(k) (chroma_from_luma, :1497-1567).
//   cm1      = C[8k-1]
//   c[i]     = C[kT k + i], i < (kLB+1) kT   (read up to kT+6; zero beyond the line end)
// outputs Yb (box-filtered luma), Ib, Qb for the 8 pixels of the block.
// chc[] carries the chroma residues ch[kT .. kT+4] of this block = ch[0..4] of the next one, so that an interior
// step only filters the kT new positions (the other box values are recovered exactly: box[m] = C[m+2] - ch[m],
// all integers); the edge variants recompute the whole window and only refresh the carry.
template <typename R, int MODE>
CVS_HD void demod_block(const RowConst<R> &rc, int k, int w, int amp, R cm1, const R c[(kLB + 1) * kT],
                        R Yb[kT], V2<R> IQb[kT], R chc[5]) {
    constexpr bool EDGE = MODE >= 1, GEN = MODE == 2;
    const int x0 = k * kT;
    // box[m] = (C[m-1] + C[m] + C[m+1] + C[m+2]) / 4 ; chroma[m] = C[m+2] - box[m], m = 0..12
    // (all values are integers < 2^24, so the running sum is exact in any order)
    constexpr int NCH = kT + 5;       // chroma samples needed: up to (first even pixel of the next block) + xi + 1
    R ch[NCH];
    if (!EDGE && CVS_DEMOD_CARRY) {
        CVS_UNROLL
        for (int m = 0; m < 5; m++) ch[m] = chc[m];
        CVS_UNROLL
        for (int m = 0; m < kT && m < 5; m++) Yb[m] = Num<R>::sub(c[m + 2], ch[m]);
        R sum = Num<R>::add(Num<R>::add(c[4], c[5]), Num<R>::add(c[6], c[7]));
        CVS_UNROLL
        for (int m = 5; m < NCH; m++) {
            if (m > 5) sum = Num<R>::add(Num<R>::sub(sum, c[m - 2]), c[m + 2]);
            const R box = div4_trunc<R>(sum);
            if (m < kT) Yb[m] = box;
            ch[m] = Num<R>::sub(c[m + 2], box);
        }
    } else {
        R sum = Num<R>::add(Num<R>::add(cm1, c[0]), Num<R>::add(c[1], c[2]));
        CVS_UNROLL
        for (int m = 0; m < NCH; m++) {
            if (m > 0) sum = Num<R>::add(Num<R>::sub(sum, (m == 1) ? cm1 : c[m - 2]), c[m + 2]);
            const R box = div4_trunc<R>(sum);
            if (m < kT) Yb[m] = box;
            ch[m] = Num<R>::sub(c[m + 2], box);
        }
    }
    CVS_UNROLL
    for (int m = 0; m < 5; m++) chc[m] = ch[kT + m];
    if (EDGE) {
        // carrier sign flips with their line-end / line-start conditions (:1539-1542), then the
        // amplitude rescale (:1544-1546).  In the interior the flip folds into sgA below.
        CVS_UNROLL
        for (int m = 0; m < NCH; m++) {
            const int q = x0 + m;
            const int ph = (q + rc.xi) & 3;
            const int g = q - ph;                      // start of this carrier period
            const bool flip = (ph >= 2) && (g >= 0) && (g + 3 < w);
            R v = flip ? -ch[m] : ch[m];
            if (GEN && amp != 50) v = (R)muldiv_trunc((int)v, 50, amp);
            ch[m] = v;
        }
    }
    // even pixels: I = -chroma[x+xi], Q = -chroma[x+xi+1]  (:1549-1552); 5 even positions: 4 in
    // the block plus the first of the next block (needed by the odd-pixel interpolation)
    const bool x1 = (rc.xi & 1) != 0, x2 = (rc.xi & 2) != 0;
    // With the default 180-degree line phase every row has xi in {0, 2}: the first level of the 4-way
    // selection is then the identity for the whole warp (MODE_FAST_EVEN).
    constexpr int NE = kT / 2 + 1;    // even positions: kT/2 in the block plus the first of the next block
    R s01[2 * NE], s23[2 * NE];      // [2e] = I candidate, [2e+1] = Q candidate
    if (MODE < 0) {                     // MODE_FAST_EVEN
        CVS_UNROLL
        for (int e = 0; e < NE; e++) {
            s01[2 * e] = ch[2 * e]; s23[2 * e] = ch[2 * e + 2];
            s01[2 * e + 1] = ch[2 * e + 1]; s23[2 * e + 1] = ch[2 * e + 3];
        }
    } else {
        CVS_UNROLL
        for (int e = 0; e < NE; e++) {
            const int j = 2 * e;
            s01[2 * e] = x1 ? ch[j + 1] : ch[j];
            s23[2 * e] = x1 ? ch[j + 3] : ch[j + 2];
            s01[2 * e + 1] = x1 ? ch[j + 2] : ch[j + 1];
            s23[2 * e + 1] = x1 ? ch[j + 4] : ch[j + 3];
        }
    }
    V2<R> IQe[NE];
    CVS_UNROLL
    for (int e = 0; e < NE; e++) {
        const int j = 2 * e;
        R iv = x2 ? s23[2 * e] : s01[2 * e];
        R qv = x2 ? s23[2 * e + 1] : s01[2 * e + 1];
        if (EDGE) {
            const int x = x0 + j;
            const bool ok = (x + rc.xi + 1) < w;        // :1549, else zero (:1553-1556)
            IQe[e] = mk2(ok ? -iv : (R)0, ok ? -qv : (R)0);
        } else {
            const R sg = (j & 2) ? -rc.sgA : rc.sgA;    // -(flip ? -1 : 1) folded
            IQe[e] = Num<R>::scale2(mk2(iv, qv), sg);
        }
    }
    // odd pixels: average of the even neighbours, arithmetic >> 1 (:1557-1560)
    CVS_UNROLL
    for (int e = 0; e < kT / 2; e++) {
        IQb[2 * e] = IQe[e];
        IQb[2 * e + 1] = Num<R>::floor_half2(Num<R>::add2(IQe[e], IQe[e + 1]));
    }
    if (EDGE) {
        // after the interpolation the last pixels lose their chroma (:1561-1564); an odd pixel is
        // only interpolated while x+1 < w (:1557)
        const int zstart = (w - 1) & ~1;
        CVS_UNROLL
        for (int j = 0; j < kT; j++) {
            const int x = x0 + j;
            if (x >= zstart || ((j & 1) && (x + 1 >= w))) IQb[j] = mk2((R)0, (R)0);
        }
    }
}

// ---- code variants ---------------------------------------------------------------------------------
// Every stage is instantiated three ways (template int MODE):
//   MODE_FAST     interior steps of a default-switch configuration: no per-pixel predicates at all
//   MODE_EDGE     steps that touch a line start/end, default switches
//   MODE_GENERAL  any non-default switch (F_GENERAL): every run-time flag honoured, on every step
// Keeping MODE_EDGE free of the rarely used switches keeps the code that every line start/end pulls
// through the instruction cache small (the first ncu capture showed 17% of the kernel time waiting on
// instruction fetch in the then-combined edge+general body).
// MODE_FAST_EVEN is MODE_FAST for warps whose rows all have an even subcarrier phase index (always, with the
// default 180-degree line phase): the first level of the demodulator's 4-way selection and the I/Q choice of the
// modulator are then compile-time.  The kernel picks the loop per warp (rc.odd_any).
enum { MODE_FAST_EVEN = -1, MODE_FAST = 0, MODE_EDGE = 1, MODE_GENERAL = 2 };

// ---- the step ------------------------------------------------------------------------------------
// Exchange buffer for the vertical chroma blend: each lane publishes its pre-blend chroma block and
// receives the block of the row above (lane - 1).  On the GPU this is 16 __shfl_up_sync.
template <typename R>
struct BlendXchg {
    V2<R> uv[kT];
};

// slide a window of kLB carried blocks by one block: m1 <- last sample leaving, win <- win[1..], newest
template <typename R>
CVS_HD void window_push(R &m1, R win[kLB][kT], const R newest[kT]) {
    m1 = win[0][kT - 1];
    CVS_UNROLL
    for (int i = 0; i + 1 < kLB; i++) {
        CVS_UNROLL
        for (int j = 0; j < kT; j++) win[i][j] = win[i + 1][j];
    }
    CVS_UNROLL
    for (int j = 0; j < kT; j++) win[kLB - 1][j] = newest[j];
}

template <typename R, bool VHS, int CD, bool OUTFULL>
struct Pipeline {
    typedef Lane<R, VHS, CD, OUTFULL> L;
    typedef Num<R> N;

    // ---- A1 + A2 (+ head-switch substitution): returns C block B(s-1) in Cnew --------------------
    // pxprev = BGRA of B(s-1): read only on the tail path (raw chroma of the last `delay` pixels), so
    // the fast variant never carries it; the edge variants re-read it from memory (an L2 hit).
    // NFT: "fast noise" (a kernel instantiation of its own, so that the exact kernels carry none of it)
    template <int MODE, bool NFT = false>
    static CVS_HD void stage_a(const KConst<R> &K, const RowConst<R> &rc, L &ln, int s,
                               const uint32_t px[kT], const uint32_t pxprev[kT], const int32_t *hs_row, R Cnew[kT]) {
        constexpr bool EDGE = MODE >= 1, GEN = MODE == 2;
        constexpr bool nfast = NFT;
        const int w = K.w;
        const int p = s * kT;
        R Ycur[kT];
        V2<R> oIQcur[kT];
        // A1: RGB -> YIQ and the input chroma lowpass cascades on B(s)
        CVS_UNROLL
        for (int j = 0; j < kT; j++) {
            const int t = p + j;
            if (!EDGE || t < w) {
                R y;
                V2<R> iq;
                N::rgb2yiq2(px[j], y, iq, K);
                Ycur[j] = y;
                oIQcur[j] = N::cascade3_trunc2(ln.pIQ, iq, K.a_in, K.b_in, K.c_in);   // P[x-delay] = s, :1453
            } else {
                Ycur[j] = 0; oIQcur[j] = mk2((R)0, (R)0);
            }
        }
        // A2: composite block B(s-1)
        if (!EDGE || s >= 1) {
            const bool in_lp = !GEN || (K.flags & F_IN_LP);
#if CVS_RNG_PAIRED
            uint32_t rawL[kT];                                                       // luma draws of B(s-1)
            if (!nfast && (!GEN || K.vnoise != 0)) ln.rngL.template draw_group<kT>(kRngBase + (uint32_t)(p - kT), rawL);
#define CVS_RAW_L(j) rawL[j]
#else
            uint32_t *gL = ln.rngL.group_ptr(kRngBase + (uint32_t)(p - kT));          // luma draws of B(s-1)
            const uint32_t *gLn = ln.rngL.group_ptr(kRngBase + (uint32_t)p);
#define CVS_RAW_L(j) ln.rngL.next_in_group(gL, gLn, j, kT)
#endif
            CVS_UNROLL
            for (int j = 0; j < kT; j++) {
                const int x = p - kT + j;
                R c = 0;
                if (!EDGE || x < w) {
                    // filtered chroma arrives early by the filter delay; the last `delay` samples of
                    // a line keep their unfiltered values (:1453)
                    R iv = (j + 2 < kT) ? ln.oIQprev[(j + 2) % kT].x : oIQcur[(j + 2) % kT].x;
                    R qv = (j + 4 < kT) ? ln.oIQprev[(j + 4) % kT].y : oIQcur[(j + 4) % kT].y;
                    if (EDGE) {
                        const bool rawI = !in_lp || (x + 2 >= w), rawQ = !in_lp || (x + 4 >= w);
                        if (rawI || rawQ) {
                            R y, i, q;
                            N::rgb2yiq(pxprev[j], y, i, q, K);
                            if (rawI) iv = i;
                            if (rawQ) qv = q;
                        }
                    }
                    c = modulate<R, MODE>(rc, j, ln.Yprev[j], iv, qv, K.amp);                        // :1611
                    if (GEN && (K.flags & F_PREEMPH)) {                                              // :1613-1629
                        const R lp = N::pole(ln.pPre, c, K.a_pre, K.b_pre);
                        c = N::preemph(c, N::sub(c, lp), K.preemph);
                    }
                    if (!GEN || K.vnoise != 0) {                                                     // :1631-1644
                        c = N::add(c, (R)ln.nY);
                        if (nfast) ln.nY = noise_draw_fast(ln.nY, ln.rngL.l1, (uint32_t)(2 * K.vnoise + 1), K.vnoise);
                        else ln.nY = noise_draw(ln.nY, CVS_RAW_L(j), (uint32_t)(2 * K.vnoise + 1), K.vmagic, K.vshift, K.vnoise);
                    }
                    if (EDGE && hs_row && (rc.rflags & RF_HEADSW)) c = (R)hs_row[x];                 // :1646-1713
                }
                Cnew[j] = c;
            }
        } else {
            CVS_UNROLL
            for (int j = 0; j < kT; j++) Cnew[j] = 0;
        }
        CVS_UNROLL
        for (int j = 0; j < kT; j++) {
            ln.Yprev[j] = Ycur[j];
            ln.oIQprev[j] = oIQcur[j];
        }
    }

    // ---- B: demod of B(s-2), noise, phase; VHS luma/chroma filters ---------------------------------
    // Outputs: Yb/Ib/Qb of B(s-2) for the non-VHS path; for VHS publishes the completed delayed chroma
    // block B(s-4) in xo (pre-blend) and leaves Y3 of B(s-2) in y3new.
    template <int MODE, bool NFT = false>
    static CVS_HD void stage_b(const KConst<R> &K, const RowConst<R> &rc, L &ln, int s, const R Cnew[kT],
                               R Yb[kT], V2<R> IQb[kT], BlendXchg<R> &xo) {
        constexpr bool EDGE = MODE >= 1, GEN = MODE == 2;
        constexpr bool nfast = NFT;
        const int w = K.w;
        constexpr int LD = L::LD;
        const int k = s - 1 - kLB;                                 // kB
        if (EDGE && k < 0) {
            CVS_UNROLL
            for (int j = 0; j < kT; j++) { Yb[j] = 0; IQb[j] = mk2((R)0, (R)0); xo.uv[j] = mk2((R)0, (R)0); }
            window_push<R>(ln.Cm1, ln.Cwin, Cnew);
            return;
        }
        R c[(kLB + 1) * kT];                                       // C of B(kB) .. B(s-1)
        CVS_UNROLL
        for (int i = 0; i < kLB; i++) {
            CVS_UNROLL
            for (int j = 0; j < kT; j++) c[i * kT + j] = ln.Cwin[i][j];
        }
        CVS_UNROLL
        for (int j = 0; j < kT; j++) c[kLB * kT + j] = Cnew[j];
        if (GEN && (K.flags & F_NOCOLOR)) {                        // :1715: no demod, chroma stays zero
            CVS_UNROLL
            for (int j = 0; j < kT; j++) { Yb[j] = c[j]; IQb[j] = mk2((R)0, (R)0); }
        } else {
            demod_block<R, MODE>(rc, k, w, K.amp_back, ln.Cm1, c, Yb, IQb, ln.chc1);   // :1716
        }
        window_push<R>(ln.Cm1, ln.Cwin, Cnew);

        const int x0 = k * kT;
        if (GEN ? (K.cnoise != 0) : VHS) {                         // :1718-1735
#if CVS_RNG_PAIRED
            uint32_t rawC[2 * kT];                                                   // chroma draws of B(kB): U, V per pixel
            if (!nfast) ln.rngC.template draw_group<2 * kT>(kRngBase + 2u * (uint32_t)x0, rawC);
#define CVS_RAW_C(i) rawC[i]
#else
            uint32_t *gC = ln.rngC.group_ptr(kRngBase + 2u * (uint32_t)x0);           // chroma draws of B(kB): U, V per pixel
            const uint32_t *gCn = ln.rngC.group_ptr(kRngBase + 2u * (uint32_t)(x0 + kT));
#define CVS_RAW_C(i) ln.rngC.next_in_group(gC, gCn, i, 2 * kT)
#endif
            CVS_UNROLL
            for (int j = 0; j < kT; j++) {
                if (!EDGE || x0 + j < w) {
                    IQb[j] = N::add2(IQb[j], mk2((R)ln.nU, (R)ln.nV));
                    const uint32_t m = (uint32_t)(2 * K.cnoise + 1);
                    if (nfast) {
                        ln.nU = noise_draw_fast(ln.nU, ln.rngC.l1, m, K.cnoise);
                        ln.nV = noise_draw_fast(ln.nV, ln.rngC.l1, m, K.cnoise);
                    } else {
                        ln.nU = noise_draw(ln.nU, CVS_RAW_C(2 * j), m, K.cmagic, K.cshift, K.cnoise);
                        ln.nV = noise_draw(ln.nV, CVS_RAW_C(2 * j + 1), m, K.cmagic, K.cshift, K.cnoise);
                    }
                }
            }
        }
        if (GEN ? ((K.flags & F_PHASE) != 0) : VHS) {              // :1736-1764
            CVS_UNROLL
            for (int j = 0; j < kT; j++) {
                IQb[j] = N::rot2(IQb[j], rc.cosp, rc.sinp, rc.nsinp);
            }
        }
        if constexpr (VHS) {

        // VHS luma: 3 poles @ luma_cut reset 16, + 1.6 x highpass, then sharpen (:1793-1812, :1865-1883)
        // VHS chroma: 3 poles @ chroma_cut, output CD samples early (:1814-1836)
        V2<R> oUVcur[kT];
        CVS_UNROLL
        for (int j = 0; j < kT; j++) {
            if (!EDGE || x0 + j < w) {
                Yb[j] = N::vhs_luma(ln.pL, ln.pLpre, ln.pS, Yb[j], K);
                oUVcur[j] = N::cascade3_trunc2(ln.pUV, IQb[j], K.a_ch, K.b_ch, K.c_ch);
                if (EDGE && (x0 + j >= w - CD) && (x0 + j - (w - CD)) < kTailSlots) {
                    // the last CD samples keep their pre-filter values (:1830,:1834): stash them
                    ln.tailU[(x0 + j - (w - CD)) * ln.tail_stride] = IQb[j].x;
                    ln.tailV[(x0 + j - (w - CD)) * ln.tail_stride] = IQb[j].y;
                }
            } else {
                Yb[j] = 0; oUVcur[j] = mk2((R)0, (R)0);
            }
        }
        // delayed chroma block B(kD), kD = kB - LD: position x' = kT kD + j was produced at time x' + CD,
        // which lies in B(kB) (this step's outputs) or B(kB-1) (carried)
        CVS_UNROLL
        for (int j = 0; j < kT; j++) {
            constexpr int LDT = LD * kT;
            const int rel = j + CD - LDT;                   // index into the current outputs (time in B(kB))
            V2<R> uv = (rel >= 0) ? oUVcur[(rel + LDT) % kT] : ln.oUVprev[(rel + kT) % kT];
            if (EDGE) {
                const int xp = (k - LD) * kT + j;
                if (xp >= w - CD && xp < w && xp >= 0 && (xp - (w - CD)) < kTailSlots) {
                    uv = mk2(ln.tailU[(xp - (w - CD)) * ln.tail_stride], ln.tailV[(xp - (w - CD)) * ln.tail_stride]);
                }
            }
            xo.uv[j] = uv;
        }
        CVS_UNROLL
        for (int j = 0; j < kT; j++) ln.oUVprev[j] = oUVcur[j];
        }
    }

    // ---- B2 + C: vertical blend of B(s-4), remodulate, second demod of B(s-5) -----------------------
    // xi_ = block received from the row above (pre-blend).  Y3new = Y3 of B(s-2) (from stage_b).
    // Produces the F-stage input block: B(s-5) (recombine) or B(s-4) (s-video).
    template <int MODE>
    static CVS_HD void stage_c(const KConst<R> &K, const RowConst<R> &rc, L &ln, int s, const R Y3new[kT],
                               const BlendXchg<R> &own, const BlendXchg<R> &above,
                               R Yf[kT], V2<R> IQf[kT], int &kf) {
        constexpr bool EDGE = MODE >= 1, GEN = MODE == 2;
        constexpr int LD = L::LD;
        const int w = K.w;
        const int kD = s - 1 - kLB - LD;                              // delayed chroma / C2 block
        V2<R> UV[kT];
        CVS_UNROLL
        for (int j = 0; j < kT; j++)                                  // (delay + cur + 1) >> 1, :1857-1858 (all integers: exact)
            UV[j] = N::floor_half2(N::axpy2(above.uv[j], rc.bl_above, N::axpy2(own.uv[j], rc.bl_own, mk2((R)1, (R)1))));
        const bool svideo = GEN && (K.flags & F_SVIDEO);
        if (svideo) {                                                 // :1885: no recombine
            kf = kD;
            CVS_UNROLL
            for (int j = 0; j < kT; j++) { Yf[j] = ln.Y3hist[0][j]; IQf[j] = UV[j]; }
        } else {
            kf = kD - kLB;                                            // kC
            // C2 of B(kD) = Y3 + QAM(U,V) with subcarrier_amplitude (:1886)
            R c[(kLB + 1) * kT], c2new[kT];
            CVS_UNROLL
            for (int i = 0; i < kLB; i++) {
                CVS_UNROLL
                for (int j = 0; j < kT; j++) c[i * kT + j] = ln.C2win[i][j];
            }
            CVS_UNROLL
            for (int j = 0; j < kT; j++) {
                const int x = kD * kT + j;
                R cv = 0;
                if (!EDGE || (x >= 0 && x < w))
                    cv = modulate<R, MODE>(rc, j, ln.Y3hist[0][j], UV[j].x, UV[j].y, K.amp);
                c[kLB * kT + j] = cv;
                c2new[j] = cv;
            }
            if (!EDGE || kf >= 0) {
                demod_block<R, MODE>(rc, kf, w, K.amp, ln.C2m1, c, Yf, IQf, ln.chc2);      // :1887
            } else {
                CVS_UNROLL
                for (int j = 0; j < kT; j++) { Yf[j] = 0; IQf[j] = mk2((R)0, (R)0); }
            }
            window_push<R>(ln.C2m1, ln.C2win, c2new);
        }
        CVS_UNROLL
        for (int i = 0; i + 1 < LD; i++) {
            CVS_UNROLL
            for (int j = 0; j < kT; j++) ln.Y3hist[i][j] = ln.Y3hist[i + 1][j];
        }
        CVS_UNROLL
        for (int j = 0; j < kT; j++) ln.Y3hist[LD - 1][j] = Y3new[j];
    }

    // ---- F: dropout, output chroma lowpass, YIQ -> RGB; completes output block B(kf-1) ----------------
    // Returns true when `out` holds a complete block B(kf-1) (kf >= 1).
    // NODROP: the caller knows that no row of the warp lost its chroma (the lean interior loop)
    template <int MODE, bool NODROP = false>
    static CVS_HD bool stage_f(const KConst<R> &K, const RowConst<R> &rc, L &ln, int kf,
                               R Yf[kT], V2<R> IQf[kT], uint32_t out[kT]) {
        constexpr bool EDGE = MODE >= 1, GEN = MODE == 2;
        constexpr int OD = L::OD, ODI = L::ODI, ODQ = L::ODQ;
        const int w = K.w;
        if (EDGE && kf < 0) return false;
        const int x0 = kf * kT;
        if (!NODROP) {
            const R keep = (rc.rflags & RF_DROPOUT) ? (R)0 : (R)1;   // :1891-1901: the row loses its chroma
            CVS_UNROLL
            for (int j = 0; j < kT; j++) IQf[j] = N::scale2(IQf[j], keep);
        }
        V2<R> oIQ[kT];
        CVS_UNROLL
        for (int j = 0; j < kT; j++) {
            if (!EDGE || x0 + j < w) {
                oIQ[j] = N::cascade3_trunc2(ln.pOIQ, IQf[j], K.a_out, K.b_out, K.c_out);   // :1420-1422
            } else {
                oIQ[j] = mk2((R)0, (R)0);
            }
        }
        const bool out_lp = !GEN || (K.flags & F_OUT_LP);
        // pixels [x0-OD, x0+8-OD): Y and raw chroma delayed by OD, filtered I by OD-ODI, filtered Q by OD-ODQ
        uint32_t pk[kT];
        CVS_UNROLL
        for (int j = 0; j < kT; j++) {
            const int pos = x0 - OD + j;
            // value of an array at absolute offset (j - OD + d) relative to the current block
            const int iy = j - OD;                // Y / raw chroma index
            const int ii = j - OD + ODI;          // filtered I index (time = pos + ODI)
            const int iq = j - OD + ODQ;
            const R yv = (iy >= 0) ? Yf[(iy + kT) % kT] : ln.Ytail[(iy + OD) % OD];
            R iv = (ii >= 0) ? oIQ[(ii + kT) % kT].x : ln.oOIQtail[(ii + OD) % OD].x;
            R qv = (iq >= 0) ? oIQ[(iq + kT) % kT].y : ln.oOIQtail[(iq + OD) % OD].y;
            if (EDGE) {
                const R ir = (iy >= 0) ? IQf[(iy + kT) % kT].x : ln.IQtail[(iy + OD) % OD].x;
                const R qr = (iy >= 0) ? IQf[(iy + kT) % kT].y : ln.IQtail[(iy + OD) % OD].y;
                if (!out_lp || pos + ODI >= w) iv = ir;       // tail keeps pre-filter values (:1422)
                if (!out_lp || pos + ODQ >= w) qv = qr;
            }
            // YIQ -> RGB (:1385-1396), alpha = 0 (:1914)
            pk[j] = N::yiq2bgra(yv, iv, qv);
        }
        // out block B(kf-1): slots 0..7-OD carried from the previous step, slots 8-OD..7 are pk[0..OD-1]
        CVS_UNROLL
        for (int j = 0; j < kT; j++) out[j] = (j < kT - OD) ? ln.outprev[j % (kT - OD > 0 ? kT - OD : 1)] : pk[(j - (kT - OD) + kT) % kT];
        CVS_UNROLL
        for (int j = 0; j < kT - OD; j++) ln.outprev[j] = pk[j + OD];
        CVS_UNROLL
        for (int d = 0; d < OD; d++) {
            ln.Ytail[d] = Yf[kT - OD + d];
            ln.IQtail[d] = IQf[kT - OD + d];
            ln.oOIQtail[d] = oIQ[kT - OD + d];
        }
        return kf >= 1;
    }
};

// Head-switch substitution of a finished composite block B(k) (fast path; the general path does it
// inline): rows rotated by the head switch read their composite signal from the scratch row.
template <typename R>
CVS_HD void headswitch_substitute(const RowConst<R> &rc, const int32_t *hs_row, int k, R C[kT]) {
    if (hs_row && (rc.rflags & RF_HEADSW)) {
        CVS_UNROLL
        for (int j = 0; j < kT; j++) C[j] = (R)hs_row[k * kT + j];
    }
}

// In-kernel head switch for the common case (ffmpeg_ntsc.cpp:1683-1700 with a small negative shift
// whose wrapped-in part is entirely zero padding): Y'[x] = Y[x - d] for x >= d, 0 below.  The finished
// composite block B(k) goes through a per-lane ring (ring[slot * stride], already offset by the lane) and
// comes back delayed by the lane's own d (0 for rows that are not rotated).  The ring holds kHsRing samples
// plus a second copy of its first block behind the last one, so the kT samples a lane reads are always
// contiguous: one wrap test per lane and step instead of one modulo per sample (the block index is
// warp-uniform, so the write position is uniform arithmetic).
template <typename R>
CVS_HD void headswitch_delay_block(R *ring, int stride, int k, int w, int d, R C[kT]) {
    static_assert(kHsRing % kT == 0, "a block never wraps inside the ring");
    constexpr int kBlocks = kHsRing / kT;
    const int x0 = k * kT;
    const int b0 = (k % kBlocks) * kT;                  // k >= 0 at every call site
    CVS_UNROLL
    for (int j = 0; j < kT; j++) ring[(b0 + j) * stride] = C[j];
    if (b0 == 0) {
        CVS_UNROLL
        for (int j = 0; j < kT; j++) ring[(kHsRing + j) * stride] = C[j];
    }
    int r0 = b0 - d;                                    // 0 <= d <= kHsMaxDelay < kHsRing: one correction
    r0 += (r0 < 0) ? kHsRing : 0;
    CVS_UNROLL
    for (int j = 0; j < kT; j++) {
        const R v = ring[(r0 + j) * stride];
        C[j] = (x0 + j >= d && x0 + j < w) ? v : (R)0;   // the composite signal is zero before the line and beyond its end
    }
}

// number of steps a line of width w takes: its blocks plus the pipeline depth
template <bool VHS, int CD>
CVS_HD int line_steps(int w) { return (w + kT - 1) / kT + (VHS ? 2 + 2 * kLB + lag_blocks(CD) : 2 + kLB); }

// [s_lo, s_hi): steps whose every stage works on an interior block, so the fast variant is valid.
// Block indices at step s: kB = s-1-kLB (first demod, VHS filters), kD = kB-LD (delayed chroma, C2),
// kC = kD-kLB (second demod).  Line start: the first block of each demodulation (index 0) has the
// carrier-period condition g >= 0, so kB >= 1 and, with VHS, kC >= 1.  Line end, for step s: A1 reads
// B(s) whole and the loop prefetches B(s+1): kT(s+2) <= w;  A2 builds B(s-1) and must stay clear of the raw-chroma tail (x+4 >= w):
// kT s + 3 < w;  the VHS chroma stash starts at x >= w-CD for x in B(kB): kT(s-kLB) - 1 < w - CD.
// Everything downstream is older and weaker.
template <bool VHS, int CD>
CVS_HD void interior_steps(int w, int &s_lo, int &s_hi) {
    constexpr int LD = lag_blocks(CD);
    s_lo = VHS ? 2 + 2 * kLB + LD : 2 + kLB;
    // the interior loop prefetches B(s+1) with an unguarded load, so B(s+1) must lie inside the row as
    // well: kT(s+2) <= w (a row that ends the picture ends the caller's buffer)
    int hi = w / kT - 1;
    const int a2 = (w - 3 + kT - 1) / kT;                // kT s < w - 3
    if (a2 < hi) hi = a2;
    if (VHS) {
        const int st = (w - CD + 1 + kT - 1) / kT + kLB; // kT (s - kLB) < w - CD + 1
        if (st < hi) hi = st;
    }
    s_hi = hi < s_lo ? s_lo : hi;
}

// ---- row / lane set-up ------------------------------------------------------------------------------
constexpr int kWarmPx = 64;   // noise warm-up length in pixels (luma: 64 draws, chroma: 128 draws)

// packed per-row side info written by the host (field_plan.cpp): phase-noise state in the low 16 bits
// (two's complement), row flags in bits 16..23, in-kernel head-switch delay in bits 24..31
CVS_HD uint32_t rowinfo_pack(int phase_state, uint32_t rflags, int hs_delay) {
    return ((uint32_t)phase_state & 0xFFFFu) | (rflags << 16) | ((uint32_t)hs_delay << 24);
}

// subcarrier phase index of a scanline, ffmpeg_ntsc.cpp:1473-1480
CVS_HD int line_phase(int phase_shift, int off, unsigned long long fieldno, unsigned y) {
    if (phase_shift == 90) return (int)((fieldno + (unsigned long long)(long long)off + (y >> 1)) & 3);
    if (phase_shift == 180) return (int)((((fieldno + y) & 2) + (unsigned long long)(long long)off) & 3);
    if (phase_shift == 270) return (int)((fieldno + (unsigned long long)(long long)off - (y >> 1)) & 3);
    return off & 3;
}

template <typename R>
CVS_HD void row_setup(const KConst<R> &K, unsigned field, unsigned long long fieldno, int row, uint32_t rowinfo,
                      RowConst<R> &rc) {
    rc.row = row;
    rc.rflags = (rowinfo >> 16) & 0xFFu;
    rc.hs_delay = (rc.rflags & RF_HEADSW_INLINE) ? (int)(rowinfo >> 24) : 0;
    rc.odd_any = true;
    const bool blend = (K.flags & F_VBLEND) && row >= 1;     // the loop starts at field+2 (:1849) ...
    rc.bl_above = (R)((blend && row >= 2) ? 1 : 0);           // ... where the delay line is still zero (:1847-1848)
    rc.bl_own = (R)(blend ? 1 : 2);
    rowconst_set_phase<R>(rc, line_phase(K.phase_shift, K.phase_offset, fieldno, field + 2u * (unsigned)row));
    if (K.flags & F_PHASE) {
        const int st = (int)(int16_t)(rowinfo & 0xFFFFu);
        rc.sinp = K.phase_lut[2 * (st + K.pnoise)];
        rc.cosp = K.phase_lut[2 * (st + K.pnoise) + 1];
    } else {
        rc.sinp = 0;
        rc.cosp = 1;
    }
    rc.nsinp = -rc.sinp;
}

// hist_out[k] = sum_i poly[i] * window[k + i]: the 31 raw words preceding position base + J, where
// poly = x^J and window[i] = q[base - 31 + i]   (glibc_rand.h)
CVS_HD void rng_rebase(const uint32_t *window, const uint32_t *poly, uint32_t hist_out[31]) {
    for (int k = 0; k < 31; k++) {
        uint32_t acc = 0;
        for (int i = 0; i < 31; i++) acc += poly[i] * window[k + i];
        hist_out[k] = acc;
    }
}

// The noise accumulators carry across rows (:1633,:1720) but each lane starts its row cold.  The
// update n' = trunc((n + d)/2) is monotone in n, so running it from both ends of the state range
// [-v, v] over the draws preceding the row brackets the true state; after a few dozen draws the two
// runs coincide (they merge with probability 1/2 per draw once adjacent) and the state is exact.
// Returns false if they did not merge (the host then retries from further back).
// `bracket` (optional): a [lo, hi] pair narrower than [-v, v], from rewarm_bracket().
CVS_HD void rng_begin_groups(LaneRng &, uint32_t) {}
CVS_HD void rng_begin_groups(LaneRngP &g, uint32_t n_first) { g.begin_groups(n_first); }
template <typename G>
CVS_HD bool warm_luma(const uint32_t mod, uint32_t magic, uint32_t shift, int v, G &g, int ndraws,
                      bool from_field_start, int &state, const int *bracket = nullptr) {
    int lo = from_field_start ? 0 : -v, hi = from_field_start ? 0 : v;
    if (bracket) { lo = bracket[0]; hi = bracket[1]; }
    for (int i = 0; i < ndraws; i++) {
        const int d = draw_mod(g.next_raw(kRngBase - (uint32_t)ndraws + (uint32_t)i), mod, magic, shift);
        lo = noise_step(lo, d, v);
        hi = noise_step(hi, d, v);
    }
    state = lo;
    rng_begin_groups(g, kRngBase);
    return lo == hi;
}
template <typename G>
CVS_HD bool warm_chroma(const uint32_t mod, uint32_t magic, uint32_t shift, int v, G &g, int npx,
                        bool from_field_start, int &su, int &sv, const int *bracket = nullptr) {
    int lou = from_field_start ? 0 : -v, hiu = from_field_start ? 0 : v, lov = lou, hiv = hiu;
    if (bracket) { lou = bracket[0]; hiu = bracket[1]; lov = bracket[2]; hiv = bracket[3]; }
    for (int i = 0; i < npx; i++) {
        const uint32_t n = kRngBase - 2u * (uint32_t)npx + 2u * (uint32_t)i;
        const int du = draw_mod(g.next_raw(n), mod, magic, shift);
        const int dv = draw_mod(g.next_raw(n + 1), mod, magic, shift);
        lou = noise_step(lou, du, v); hiu = noise_step(hiu, du, v);
        lov = noise_step(lov, dv, v); hiv = noise_step(hiv, dv, v);
    }
    su = lou; sv = lov;
    rng_begin_groups(g, kRngBase);
    return lou == hiu && lov == hiv;
}

// Second chance of the warm-up (cold path; the first attempt fails with p < 2^-58 per row): the generator is
// invertible, q[n-31] = q[n] - q[n-3], so from the 31 words `hist` preceding the warm-up the lane steps BACKWARD by
// streams * extra_px draws, then forward again over them with the two bracketing runs.  The returned bracket(s)
// [lo, hi] per stream (1 = luma, 2 = chroma U, V interleaved) then seed the regular warm-up, which leaves the ring
// exactly as the first attempt did.  extra_px must not reach before the field's first draw of the stream; when it
// reaches it exactly (at_field_start) the state there is the known 0.
constexpr int kRewarmPx = 2048;
CVS_HD void rewarm_bracket(const uint32_t hist[31], int streams, int extra_px, bool at_field_start, uint32_t mod,
                           uint32_t magic, uint32_t shift, int v, int bracket[4]) {
    uint32_t b[31];                       // slot k holds q[t0 - 31 + k + 31 j] for the j that is current
    for (int k = 0; k < 31; k++) b[k] = hist[k];
    const int n = streams * extra_px;
    int slot = 0;                         // slot of q[t0 - i], i = 1..n, walking down: (31 - i) mod 31
    for (int i = 1; i <= n; i++) {
        slot = slot == 0 ? 30 : slot - 1;
        const int s3 = slot >= 3 ? slot - 3 : slot + 28;
        b[slot] = b[slot] - b[s3];        // q[t0 - i - 31] = q[t0 - i] - q[t0 - i - 3], stored where q[t0 - i] was
    }
    for (int k = 0; k < 4; k++) bracket[k] = (k & 1) ? (at_field_start ? 0 : v) : (at_field_start ? 0 : -v);
    for (int i = 0; i < n; i++) {         // forward again: position t0 - n + i lives in the slot the walk is at
        const int s3 = slot >= 3 ? slot - 3 : slot + 28;
        const uint32_t qv = b[slot] + b[s3];
        b[slot] = qv;
        slot = slot == 30 ? 0 : slot + 1;
        const int d = draw_mod(qv, mod, magic, shift);
        const int st = (streams == 2) ? (i & 1) : 0;
        bracket[2 * st] = noise_step(bracket[2 * st], d, v);
        bracket[2 * st + 1] = noise_step(bracket[2 * st + 1], d, v);
    }
}

// 8 BGRA pixels of B(k) of a source row; zero beyond the line end
CVS_HD void load_block_scalar(const uint32_t *srow, int k, int w, uint32_t px[kT]) {
    CVS_UNROLL
    for (int j = 0; j < kT; j++) {
        const int x = k * kT + j;
        px[j] = (x < w) ? srow[x] : 0u;
    }
}

// Head-switch pre-pass (ffmpeg_ntsc.cpp:1683-1700): a rotated row needs random access to the whole
// composite row, so rows with a non-zero shift get their composite signal (stages A1/A2, including
// the exact luma noise) computed up front and written, already rotated, to a scratch row; the main
// pass substitutes it for its own C.  dest[(k - shif) mod twidth] = C[k]; everything else is the
// zero padding of the reference's tmp[] ring.
template <typename R, bool NFT = false>
CVS_HD void headswitch_row(const KConst<R> &K, const RowConst<R> &rc_in, Lane<R, false, 9, false> &ln,
                           const uint32_t *srow, int32_t *scratch, int shif) {
    typedef Pipeline<R, false, 9, false> P;
    const int w = K.w;
    const int tw = w + w / 10;
    RowConst<R> rc = rc_in;
    rc.rflags &= ~(uint32_t)RF_HEADSW;                 // compute, do not substitute
    for (int x = 0; x < w; x++) scratch[x] = 0;
    const int nb = (w + kT - 1) / kT;
    uint32_t pxprev[kT];
    CVS_UNROLL
    for (int j = 0; j < kT; j++) pxprev[j] = 0;
    for (int s = 0; s <= nb; s++) {
        uint32_t px[kT];
        load_block_scalar(srow, s, w, px);
        R C[kT];
        P::template stage_a<MODE_GENERAL, NFT>(K, rc, ln, s, px, pxprev, (const int32_t *)0, C);
        CVS_UNROLL
        for (int j = 0; j < kT; j++) pxprev[j] = px[j];
        if (s >= 1) {
            for (int j = 0; j < kT; j++) {
                const int x = (s - 1) * kT + j;
                if (x < w) {
                    int xd = (x - shif) % tw;
                    if (xd < 0) xd += tw;
                    if (xd < w) scratch[xd] = (int32_t)C[j];
                }
            }
        }
    }
}

}  // namespace cvs
#endif

