// scanline_kernels.cuh -- sm_100a kernels around lane_pipeline.cuh.
//
//   k_fields<R,VHS,CD,OUTFULL>   the fused scanline kernel: one launch processes a batch of
//                                 fields; a warp owns 31 consecutive rows of one field (+1 halo
//                                 lane), a lane streams its row through every stage of
//                                 composite_layer() (ffmpeg_ntsc.cpp:1570-1921) in registers.
//                                 HBM traffic = one BGRA read + one BGRA write per pixel.
//   k_headswitch<R>               pre-pass for the few rows per field that the VHS head switch
//                                 rotates (ffmpeg_ntsc.cpp:1683-1700): writes their composite
//                                 signal, already rotated, to a scratch row.
//
// Shared memory per CTA (NT threads): the per-lane glibc-rand rings for the luma and chroma
// noise streams ([2][32][NT] words, slot-major so a warp access is conflict-free), the per-lane
// chroma tail stash ([2][16][NT] R), one 64-word generator window per warp and the per-lane delay
// ring of the in-kernel head switch ([64][NT] R).
#ifndef CVS_SCANLINE_KERNELS_CUH
#define CVS_SCANLINE_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "lane_pipeline.cuh"

namespace cvs {

#ifndef CVS_NT
#define CVS_NT 128
#endif
constexpr int kNT = CVS_NT;              // threads per CTA
constexpr int kWarpsPerCta = kNT / 32;
#ifndef CVS_MIN_WARPS
#define CVS_MIN_WARPS 12            // warps per SM the register allocator must leave room for (168 registers)
#endif
#define CVS_MIN_CTAS (CVS_MIN_WARPS / (CVS_NT / 32))
#ifndef CVS_FAST_UNROLL
#define CVS_FAST_UNROLL 0       // steps per iteration of the lean interior loop: 0 = per kernel (see there), 1, 2
#endif
#ifndef CVS_LEAN_LOOP
#define CVS_LEAN_LOOP 1         // a lean interior loop for warps without head-switch / dropout rows
#endif
#ifndef CVS_EVEN_LOOP
#define CVS_EVEN_LOOP 0         // an even-phase interior loop of their own for the warps that are not lean (else: the generic one)
#endif
constexpr int kRowsPerWarp = 31;         // lane 0 is the halo row of the vertical chroma blend

struct FieldDesc {
    const uint8_t *src;                  // picture base (device)
    uint8_t *dst;
    const uint32_t *rowinfo;             // nl packed row records (rowinfo_pack)
    const uint32_t *seek;                // nl * 62 jump-polynomial words for this field parity
    int32_t *hs_scratch;                 // hs_count rows of w int32 (rotated composite rows)
    const int32_t *hs_shift;             // hs_count shifts
    unsigned long long fieldno;
    int32_t field, nl, hs_first, hs_count;
    int32_t row_start, pad_;             // rows of the preceding fields of the batch (packed mapping)
    uint32_t window[64];                 // q[base-31 .. base+29] (61 used)
};

struct HsItem {
    int32_t field_idx, slot;
};

template <typename R>
struct LaunchArgs {
    KConst<R> K;
    const FieldDesc *fields;
    int32_t nfields, warps_per_field, total_warps;
    // packed mapping: the rows of all fields of the batch form one sequence that is cut into warps of 31
    // rows, so that only the last warp of the launch has idle lanes (a field of 540 rows is 17.4 warps, not 18)
    int32_t packed, total_rows, max_nl;
    int32_t src_stride, dst_stride, opposite;
    int32_t vec_src, vec_dst;            // rows are 16-byte aligned: use 128-bit loads / stores
    int32_t bob;                         // also write every processed row y >= 1 into row y-1 (ffmpeg_ntsc.cpp:2232-2257)
    int32_t warm_px;                     // noise warm-up length in pixels (kWarmPx)
    int32_t *status;                     // sticky error word (CVS_ERR_NOISE_SYNC)
};

// Shared memory layout of k_fields (bytes): [rng rings: 1 (luma) or 2 (luma+chroma) x 32 x NT words]
// [generator windows: 64 words per warp] and, for the VHS kernels only, [chroma tail stash 2 x 16 x NT R]
// [head-switch delay ring 64 x NT R].  The composite-only kernels need a third of the VHS footprint.
template <typename R, bool VHS>
struct SmemLayout {
    static constexpr size_t rings = (size_t)2 * kRngSlots * kNT * sizeof(uint32_t);
    static constexpr size_t wins = (size_t)kWarpsPerCta * 128 * sizeof(uint32_t);   // two fields can meet in a warp
    static constexpr size_t tails = VHS ? (size_t)2 * kTailSlots * kNT * sizeof(R) : 0;
    static constexpr size_t hsring = VHS ? (size_t)kHsSlots * kNT * sizeof(R) : 0;
    static constexpr size_t off_wins = rings, off_tails = rings + wins, off_hsring = rings + wins + tails;
    static constexpr size_t total = rings + wins + tails + hsring;
};

__device__ __forceinline__ const uint32_t *row_ptr(const uint8_t *base, int stride, int y) {
    return (const uint32_t *)(base + (size_t)stride * (size_t)y);
}

// 31 raw words preceding the lane's noise segment: hist[k] = sum_i poly[i] * window[k+i]
__device__ __forceinline__ void rebase_dev(const uint32_t *win_smem, const uint32_t *__restrict__ poly_g,
                                           uint32_t hist[31]) {
    uint32_t poly[31];
#pragma unroll
    for (int i = 0; i < 31; i++) poly[i] = __ldg(poly_g + i);
#pragma unroll 1
    for (int k = 0; k < 31; k++) {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 31; i++) acc += poly[i] * win_smem[k + i];
        hist[k] = acc;
    }
}

// The warm-up's second chance (lane_pipeline.cuh, rewarm_bracket), out of line: it runs for one row in 2^58.
static __device__ __noinline__ void rewarm_dev(const uint32_t *hist, int streams, int extra_px, bool at_field_start, uint32_t mod,
                                        uint32_t magic, uint32_t shift, int v, int *bracket) {
    rewarm_bracket(hist, streams, extra_px, at_field_start, mod, magic, shift, v, bracket);
}

// Both streams at once: one pass over the window feeds two accumulators (half the shared-memory reads of two passes).
__device__ __forceinline__ void rebase2_dev(const uint32_t *win_smem, const uint32_t *__restrict__ poly_g,
                                            uint32_t histL[31], uint32_t histC[31]) {
    uint32_t pl[31], pc[31];
#pragma unroll
    for (int i = 0; i < 31; i++) { pl[i] = __ldg(poly_g + i); pc[i] = __ldg(poly_g + 31 + i); }
#pragma unroll 1
    for (int k = 0; k < 31; k++) {
        uint32_t al = 0, ac = 0;
#pragma unroll
        for (int i = 0; i < 31; i++) {
            const uint32_t wv = win_smem[k + i];
            al += pl[i] * wv;
            ac += pc[i] * wv;
        }
        histL[k] = al;
        histC[k] = ac;
    }
}

// Each lane reads its own row, 32 bytes (one sector) per step; four consecutive steps share one
// 128-byte line.  The L2::128B hint makes the first touch bring the whole line into L2, so DRAM sees
// every line exactly once (the first ncu capture, taken with evict-first loads, showed 1.7x the
// algorithmic read traffic).
__device__ __forceinline__ uint4 ld_row16(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void load_block_dev(const uint32_t *srow, int k, int w, bool vec, uint32_t px[kT]) {
    const int x0 = k * kT;
    if (vec && x0 + kT <= w) {
        const uint4 *p = reinterpret_cast<const uint4 *>(srow + x0);
#pragma unroll
        for (int v = 0; v < kT / 4; v++) {
            const uint4 a = ld_row16(p + v);
            px[4 * v] = a.x; px[4 * v + 1] = a.y; px[4 * v + 2] = a.z; px[4 * v + 3] = a.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kT; j++) px[j] = (x0 + j < w) ? __ldg(srow + x0 + j) : 0u;
    }
}

// interior blocks: the whole block is inside the line
__device__ __forceinline__ void load_block_fast(const uint32_t *srow, int k, bool vec, uint32_t px[kT]) {
    const int x0 = k * kT;
    if (vec) {
        const uint4 *p = reinterpret_cast<const uint4 *>(srow + x0);
#pragma unroll
        for (int v = 0; v < kT / 4; v++) {
            const uint4 a = ld_row16(p + v);
            px[4 * v] = a.x; px[4 * v + 1] = a.y; px[4 * v + 2] = a.z; px[4 * v + 3] = a.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kT; j++) px[j] = __ldg(srow + x0 + j);
    }
}

template <typename R, bool VHS, int CD, bool OUTFULL, bool NF>
struct Stepper {
    typedef Lane<R, VHS, CD, OUTFULL> L;
    typedef Pipeline<R, VHS, CD, OUTFULL> P;

    // LEAN (interior steps only): the warp has no pre-pass head-switch row and no dropout row, and the pictures are
    // 16-byte aligned, so none of that is compiled into the loop.  The instructions this removes were branched over,
    // not executed; what it buys is instruction-cache footprint (profiles/ab_variants_r2.txt): the interior loop
    // and the line-end body compete for the same cache, and a SECOND interior loop live on the SM costs more than
    // the dead code did -- so the rows delayed by the default head switch stay in this loop behind one branch.
    template <int MODE, bool LEAN = false>
    static __device__ __forceinline__ void step(const KConst<R> &K, const RowConst<R> &rc, L &ln, int s,
                                                const uint32_t px[kT], const uint32_t *srow, bool vec_src,
                                                const int32_t *hsrow, bool warp_hs,
                                                R *hsring, bool warp_inl, bool valid, uint32_t *drow, uint32_t *drow_bob, bool vec_dst) {
        constexpr bool EDGE = MODE >= 1;
        static_assert(!(LEAN && EDGE), "the lean step is an interior step");
        if (LEAN) { warp_hs = false; vec_dst = true; }     // (the in-kernel head-switch delay stays: a warp-uniform branch)
        R C[kT], Yb[kT];
        V2<R> IQb[kT];
        BlendXchg<R> xo;
        uint32_t pxprev[kT];
        if (EDGE) {
            // the tail path needs the raw chroma of the previous block: re-read it (L2 hit) instead of
            // carrying 8 registers through every interior step
            if (s >= 1) load_block_dev(srow, s - 1, K.w, vec_src, pxprev);
            else {
#pragma unroll
                for (int j = 0; j < kT; j++) pxprev[j] = 0;
            }
        }
        P::template stage_a<MODE, NF>(K, rc, ln, s, px, pxprev, hsrow, C);
        if (!EDGE && warp_hs) headswitch_substitute<R>(rc, hsrow, s - 1, C);
        if (warp_inl && (!EDGE || s >= 1)) headswitch_delay_block<R>(hsring, kNT, s - 1, K.w, rc.hs_delay, C);
        P::template stage_b<MODE, NF>(K, rc, ln, s, C, Yb, IQb, xo);
        uint32_t out[kT];
        bool have;
        int kf;
        if (VHS) {
            BlendXchg<R> above;
#pragma unroll
            for (int j = 0; j < kT; j++) {
                above.uv[j].x = __shfl_up_sync(0xffffffffu, xo.uv[j].x, 1);
                above.uv[j].y = __shfl_up_sync(0xffffffffu, xo.uv[j].y, 1);
            }
            R Yf[kT];
            V2<R> IQf[kT];
            P::template stage_c<MODE>(K, rc, ln, s, Yb, xo, above, Yf, IQf, kf);
            have = P::template stage_f<MODE, LEAN>(K, rc, ln, kf, Yf, IQf, out);
        } else {
            kf = s - 1 - kLB;
            have = P::template stage_f<MODE, LEAN>(K, rc, ln, kf, Yb, IQb, out);
        }
        if (LEAN) have = true;                             // interior: s >= LAG, so block kf - 1 >= 0 is complete
        if (have && valid) {
            const int x0 = (kf - 1) * kT;
            if (vec_dst && (!EDGE || x0 + kT <= K.w)) {
                uint4 *d = reinterpret_cast<uint4 *>(drow + x0);
#pragma unroll
                for (int v = 0; v < kT / 4; v++) d[v] = make_uint4(out[4 * v], out[4 * v + 1], out[4 * v + 2], out[4 * v + 3]);
                if (drow_bob) {                            // line doubling: the same pixels one row up
                    uint4 *d2 = reinterpret_cast<uint4 *>(drow_bob + x0);
#pragma unroll
                    for (int v = 0; v < kT / 4; v++) d2[v] = make_uint4(out[4 * v], out[4 * v + 1], out[4 * v + 2], out[4 * v + 3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < kT; j++)
                    if (x0 + j < K.w) {
                        drow[x0 + j] = out[j];
                        if (drow_bob) drow_bob[x0 + j] = out[j];
                    }
            }
        }
    }
};

// NF: fast noise mode (cvs_set_noise_mode): per-pixel noise from per-row counter generators; fp32 only
template <typename R, bool VHS, int CD, bool OUTFULL, bool NF = false>
__global__ void __launch_bounds__(kNT, (sizeof(R) == 4 ? CVS_MIN_CTAS : 1)) k_fields(const __grid_constant__ LaunchArgs<R> a) {
    typedef Lane<R, VHS, CD, OUTFULL> L;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *rings = reinterpret_cast<uint32_t *>(smem_raw);
    typedef SmemLayout<R, VHS> SL;
    uint32_t *wins = reinterpret_cast<uint32_t *>(smem_raw + SL::off_wins);
    R *tails = reinterpret_cast<R *>(smem_raw + SL::off_tails);          // VHS only
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gw = blockIdx.x * kWarpsPerCta + warp;
    if (gw >= a.total_warps) return;
    const KConst<R> &K = a.K;
    const int w = K.w, h = K.h;
    // which field row this lane computes.  Lane 0 is the halo row: the row above lane 1's.
    int fi, row;
    bool valid;
    if (a.packed) {
        int g = kRowsPerWarp * gw + lane - 1;                 // position in the batch's sequence of rows
        valid = (lane >= 1) && g < a.total_rows;
        g = g < 0 ? 0 : (g > a.total_rows - 1 ? a.total_rows - 1 : g);
        fi = g / a.max_nl;                                    // never beyond the row's field: row_start <= fi * max_nl
        while (g >= a.fields[fi].row_start + a.fields[fi].nl) fi++;
        row = g - a.fields[fi].row_start;
        // a halo row in the previous field is not needed (lane 1 is then row 0, which is not blended): park it
        const int fi1 = __shfl_sync(0xffffffffu, fi, 1);
        if (lane == 0 && fi != fi1) { fi = fi1; row = 0; }
    } else {
        fi = gw / a.warps_per_field;
        const int wp = gw - fi * a.warps_per_field, nlw = a.fields[fi].nl;
        if (kRowsPerWarp * wp >= nlw) return;                 // (odd heights: the short parity has fewer rows)
        row = kRowsPerWarp * wp + lane - 1;
        valid = (lane >= 1) && row < nlw;
        row = row < 0 ? 0 : (row > nlw - 1 ? nlw - 1 : row);
    }
    const FieldDesc &fd = a.fields[fi];
    const int nl = fd.nl;
    (void)nl;

    // generator windows of the (at most two) fields of this warp: [0,64) the field of lane 1, [64,128) that of lane 31
    uint32_t *win = wins + warp * 128;
    {
        const int fiA = __shfl_sync(0xffffffffu, fi, 1), fiB = __shfl_sync(0xffffffffu, fi, 31);
        win[lane] = a.fields[fiA].window[lane];
        win[lane + 32] = a.fields[fiA].window[lane + 32];
        win[lane + 64] = a.fields[fiB].window[lane];
        win[lane + 96] = a.fields[fiB].window[lane + 32];
        __syncwarp();
        if (fi != fiA) win += 64;
    }

    L ln;
    ln.reset(K);
    RowConst<R> rc;
    row_setup<R>(K, (unsigned)fd.field, fd.fieldno, row, __ldg(fd.rowinfo + row), rc);
    int sy = fd.field + 2 * row + a.opposite;            // source row, ffmpeg_ntsc.cpp:1599
    sy = sy > h - 1 ? h - 1 : sy;
    const uint32_t *srow = row_ptr(fd.src, a.src_stride, sy);
    uint32_t *drow = const_cast<uint32_t *>(row_ptr(fd.dst, a.dst_stride, fd.field + 2 * row));
    // bob (ffmpeg_ntsc.cpp:2232-2257): field 1 copies row y onto y-1, field 0 copies row y+1 onto y for
    // y = 1,3,.. < h-1: either way every processed row except row 0 is duplicated one row up
    uint32_t *drow_bob = (a.bob && (fd.field + 2 * row) >= 1)
                             ? const_cast<uint32_t *>(row_ptr(fd.dst, a.dst_stride, fd.field + 2 * row - 1)) : nullptr;
    const int32_t *hsrow = (rc.rflags & RF_HEADSW) ? fd.hs_scratch + (size_t)(row - fd.hs_first) * (size_t)w : nullptr;
    ln.tailU = tails + tid;
    ln.tailV = tails + (size_t)kTailSlots * kNT + tid;
    ln.tail_stride = kNT;
    ln.nY = ln.nU = ln.nV = 0;
    constexpr bool nfast = NF;
    if (nfast) {
        ln.rngL.l1 = lcg_seed(fd.fieldno, (unsigned)fd.field, row, 0u);
        ln.rngC.l1 = lcg_seed(fd.fieldno, (unsigned)fd.field, row, 1u);
    } else {
        const long long full = (long long)row * w;          // pixels of this field's noise segments before the row
        const int nd = (int)(full < a.warm_px ? full : a.warm_px);
        const bool from_start = full <= a.warm_px;
        // when the bracketing runs of a warm-up do not merge, the lane walks the generator further back and tries again
        const long long avail = full - nd;
        const int extra = (int)(avail < kRewarmPx ? avail : kRewarmPx);
        bool ok = true;
        uint32_t hist[31], histC[31];
        int br[4];
        const bool both = K.vnoise != 0 && K.cnoise != 0;
        if (both) rebase2_dev(win, fd.seek + (size_t)row * 62, hist, histC);
        if (K.vnoise != 0) {
            const uint32_t m = (uint32_t)(2 * K.vnoise + 1);
            if (!both) rebase_dev(win, fd.seek + (size_t)row * 62, hist);
            ln.rngL.init(rings, tid, kNT, hist, kRngBase - (uint32_t)nd);
            if (!warm_luma(m, K.vmagic, K.vshift, K.vnoise, ln.rngL, nd, from_start, ln.nY)) {
                rewarm_dev(hist, 1, extra, extra == avail, m, K.vmagic, K.vshift, K.vnoise, br);
                ln.rngL.init(rings, tid, kNT, hist, kRngBase - (uint32_t)nd);
                ok &= warm_luma(m, K.vmagic, K.vshift, K.vnoise, ln.rngL, nd, from_start, ln.nY, br);
            }
        }
        if (K.cnoise != 0) {
            const uint32_t m = (uint32_t)(2 * K.cnoise + 1);
            if (!both) rebase_dev(win, fd.seek + (size_t)row * 62 + 31, histC);
            ln.rngC.init(rings + (size_t)kRngSlots * kNT, tid, kNT, histC, kRngBase - 2u * (uint32_t)nd);
            if (!warm_chroma(m, K.cmagic, K.cshift, K.cnoise, ln.rngC, nd, from_start, ln.nU, ln.nV)) {
                rewarm_dev(histC, 2, extra, extra == avail, m, K.cmagic, K.cshift, K.cnoise, br);
                ln.rngC.init(rings + (size_t)kRngSlots * kNT, tid, kNT, histC, kRngBase - 2u * (uint32_t)nd);
                ok &= warm_chroma(m, K.cmagic, K.cshift, K.cnoise, ln.rngC, nd, from_start, ln.nU, ln.nV, br);
            }
        }
        if (!ok) atomicOr(a.status, 1);                     // (both attempts failed: p < 2^-2000)
    }

    const int nsteps = line_steps<VHS, CD>(w);
    int s_lo, s_hi;
    interior_steps<VHS, CD>(w, s_lo, s_hi);
    const bool general = (K.flags & F_GENERAL) != 0;
    if (general) s_hi = s_lo;
    const bool vec_src = a.vec_src != 0, vec_dst = a.vec_dst != 0;
    rc.odd_any = __any_sync(0xffffffffu, (rc.xi & 1) != 0);
    const bool warp_hs = __any_sync(0xffffffffu, hsrow != nullptr);
    const bool warp_inl = VHS && __any_sync(0xffffffffu, rc.hs_delay > 0);   // (the host only plans it for VHS kernels)
    R *hsring = reinterpret_cast<R *>(smem_raw + SL::off_hsring) + tid;   // VHS only (head switching needs -vhs)
    // the lean interior loop: aligned pictures and no row of the warp that needs a head-switch or dropout special case
    const bool lean = vec_src && vec_dst && !warp_hs && !__any_sync(0xffffffffu, (rc.rflags & RF_DROPOUT) != 0);
    (void)lean;

    // Step loop.  The interior (fast) steps get a loop of their own: if the three code variants shared
    // one loop the compiler would have to bring ~85 carried registers back to a common allocation at
    // every iteration (the r1e ncu capture showed exactly those moves at the loop head, 10 per pixel).
    // The edge/general loop runs twice, for the steps before and after the interior.
    typedef Stepper<R, VHS, CD, OUTFULL, NF> St;
    uint32_t px[kT];
    load_block_dev(srow, 0, w, vec_src, px);
    int s = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        const int s_end = (pass == 0) ? s_lo : nsteps;
#pragma unroll 1
        for (; s < s_end; s++) {
            uint32_t pxn[kT];
            load_block_dev(srow, s + 1, w, vec_src, pxn);            // prefetch the next block
            if (!general)
                St::template step<MODE_EDGE>(K, rc, ln, s, px, srow, vec_src, hsrow, warp_hs, hsring, warp_inl, valid, drow, drow_bob, vec_dst);
            else
                St::template step<MODE_GENERAL>(K, rc, ln, s, px, srow, vec_src, hsrow, warp_hs, hsring, warp_inl, valid, drow, drow_bob, vec_dst);
#pragma unroll
            for (int j = 0; j < kT; j++) px[j] = pxn[j];
        }
        if (pass == 0) {
            // Interior steps: three loops, picked per warp.  (One step per iteration: two steps per iteration
            // execute 9 % fewer instructions but measured slower twice, profiles/ab_variants_r1.txt and _r2.txt.)
            if (!rc.odd_any && lean && CVS_LEAN_LOOP) {
                // every row has an even line phase and none needs a pre-pass or dropout special case
                // Two steps per iteration where the doubled body still sits well in the instruction cache: the carried
                // blocks rename instead of moving (-7.5 % instructions).  Measured (profiles/ab_variants_r2.txt):
                // composite-only +13 % (9 KB body), VHS with fast noise +3.6 % (17 KB), VHS exact -2 % (19.5 KB).
                // A leftover odd step is taken by the loop below.
                if (CVS_FAST_UNROLL == 2 || (CVS_FAST_UNROLL == 0 && (!VHS || NF)))
#pragma unroll 1
                for (; s + 1 < s_hi; s += 2) {
                    uint32_t pxn[kT];
                    load_block_fast(srow, s + 1, true, pxn);
                    St::template step<MODE_FAST_EVEN, true>(K, rc, ln, s, px, srow, true, hsrow, false, hsring, warp_inl, valid, drow, drow_bob, true);
                    load_block_fast(srow, s + 2, true, px);
                    St::template step<MODE_FAST_EVEN, true>(K, rc, ln, s + 1, pxn, srow, true, hsrow, false, hsring, warp_inl, valid, drow, drow_bob, true);
                }
#pragma unroll 1
                for (; s < s_hi; s++) {
                    uint32_t pxn[kT];
                    load_block_fast(srow, s + 1, true, pxn);         // in range: interior_steps() keeps kT(s+2) <= w
                    St::template step<MODE_FAST_EVEN, true>(K, rc, ln, s, px, srow, true, hsrow, false, hsring, warp_inl, valid, drow, drow_bob, true);
#pragma unroll
                    for (int j = 0; j < kT; j++) px[j] = pxn[j];
                }
            } else if (!rc.odd_any && CVS_EVEN_LOOP) {
                // even line phase, but some row is rotated by the head switch, lost its chroma or is unaligned
#pragma unroll 1
                for (; s < s_hi; s++) {
                    uint32_t pxn[kT];
                    load_block_fast(srow, s + 1, vec_src, pxn);
                    St::template step<MODE_FAST_EVEN, false>(K, rc, ln, s, px, srow, vec_src, hsrow, warp_hs, hsring, warp_inl, valid, drow, drow_bob, vec_dst);
#pragma unroll
                    for (int j = 0; j < kT; j++) px[j] = pxn[j];
                }
            } else {
#pragma unroll 1
                for (; s < s_hi; s++) {
                    uint32_t pxn[kT];
                    load_block_fast(srow, s + 1, vec_src, pxn);
                    St::template step<MODE_FAST>(K, rc, ln, s, px, srow, vec_src, hsrow, warp_hs, hsring, warp_inl, valid, drow, drow_bob, vec_dst);
#pragma unroll
                    for (int j = 0; j < kT; j++) px[j] = pxn[j];
                }
            }
        }
    }
}

constexpr int kHsNT = 64;

template <typename R>
__global__ void __launch_bounds__(kHsNT) k_headswitch(const __grid_constant__ LaunchArgs<R> a,
                                                      const HsItem *__restrict__ items, int nitems) {
    __shared__ uint32_t ring[kRngSlots * kHsNT];
    const int it = blockIdx.x * kHsNT + threadIdx.x;
    if (it >= nitems) return;
    const HsItem item = items[it];
    const FieldDesc &fd = a.fields[item.field_idx];
    const KConst<R> &K = a.K;
    const int w = K.w, h = K.h;
    const int row = fd.hs_first + item.slot;
    Lane<R, false, 9, false> ln;
    ln.reset(K);
    RowConst<R> rc;
    row_setup<R>(K, (unsigned)fd.field, fd.fieldno, row, __ldg(fd.rowinfo + row), rc);
    ln.nY = 0;
    ln.rngL.l1 = lcg_seed(fd.fieldno, (unsigned)fd.field, row, 0u);       // fast noise mode (else init() below)
    if (K.vnoise != 0 && !(K.flags & F_NOISE_FAST)) {
        const long long full = (long long)row * w;
        const int nd = (int)(full < a.warm_px ? full : a.warm_px);
        const uint32_t m = (uint32_t)(2 * K.vnoise + 1);
        uint32_t hist[31];
        rng_rebase(fd.window, fd.seek + (size_t)row * 62, hist);
        ln.rngL.init(ring, (int)threadIdx.x, kHsNT, hist, kRngBase - (uint32_t)nd);
        if (!warm_luma(m, K.vmagic, K.vshift, K.vnoise, ln.rngL, nd, full <= a.warm_px, ln.nY)) {
            const long long avail = full - nd;
            const int extra = (int)(avail < kRewarmPx ? avail : kRewarmPx);
            int br[4];
            rewarm_dev(hist, 1, extra, extra == avail, m, K.vmagic, K.vshift, K.vnoise, br);
            ln.rngL.init(ring, (int)threadIdx.x, kHsNT, hist, kRngBase - (uint32_t)nd);
            if (!warm_luma(m, K.vmagic, K.vshift, K.vnoise, ln.rngL, nd, full <= a.warm_px, ln.nY, br)) atomicOr(a.status, 1);
        }
    }
    int sy = fd.field + 2 * row + a.opposite;
    sy = sy > h - 1 ? h - 1 : sy;
    if (K.flags & F_NOISE_FAST)
        headswitch_row<R, true>(K, rc, ln, row_ptr(fd.src, a.src_stride, sy), fd.hs_scratch + (size_t)item.slot * (size_t)w,
                                __ldg(fd.hs_shift + item.slot));
    else
        headswitch_row<R, false>(K, rc, ln, row_ptr(fd.src, a.src_stride, sy), fd.hs_scratch + (size_t)item.slot * (size_t)w,
                                 __ldg(fd.hs_shift + item.slot));
}

// host-callable launchers (one translation unit per instantiation, see kern_*.cu)
template <typename R, bool VHS, int CD, bool OUTFULL, bool NF = false>
cudaError_t launch_fields(const LaunchArgs<R> &a, cudaStream_t st);
// resident CTAs per SM of an instantiation (for wave-aligned batch sizes)
template <typename R, bool VHS, int CD, bool OUTFULL, bool NF = false>
cudaError_t occupancy_fields(int *ctas_per_sm);
template <typename R>
cudaError_t launch_headswitch(const LaunchArgs<R> &a, const HsItem *items, int nitems, cudaStream_t st);

#define CVS_DEFINE_LAUNCH_FIELDS_NF(R, VHS, CD, OUTFULL, NF)                                                          \
    template <>                                                                                                \
    cudaError_t launch_fields<R, VHS, CD, OUTFULL, NF>(const LaunchArgs<R> &a, cudaStream_t st) {                  \
        const size_t smem = SmemLayout<R, VHS>::total;                                                         \
        cudaError_t e = cudaFuncSetAttribute(k_fields<R, VHS, CD, OUTFULL, NF>,                                    \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
        if (e != cudaSuccess) return e;                                                                        \
        const int ctas = (a.total_warps + kWarpsPerCta - 1) / kWarpsPerCta;                                    \
        k_fields<R, VHS, CD, OUTFULL, NF><<<ctas, kNT, smem, st>>>(a);                                             \
        return cudaGetLastError();                                                                             \
    }                                                                                                          \
    template <>                                                                                                \
    cudaError_t occupancy_fields<R, VHS, CD, OUTFULL, NF>(int *ctas_per_sm) {                                      \
        const size_t smem = SmemLayout<R, VHS>::total;                                                         \
        cudaError_t e = cudaFuncSetAttribute(k_fields<R, VHS, CD, OUTFULL, NF>,                                    \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
        if (e != cudaSuccess) return e;                                                                        \
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_fields<R, VHS, CD, OUTFULL, NF>, kNT, smem); \
    }

#define CVS_DEFINE_LAUNCH_FIELDS(R, VHS, CD, OUTFULL) CVS_DEFINE_LAUNCH_FIELDS_NF(R, VHS, CD, OUTFULL, false)

}  // namespace cvs
#endif
