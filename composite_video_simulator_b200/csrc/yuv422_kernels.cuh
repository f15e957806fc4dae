// yuv422_kernels.cuh -- sm_100a kernels around yuv422_pipeline.cuh (the reference's 4:2:2 path,
// ffmpeg_to_composite.cpp:629-952 and :1001-1129).
//
//   k_yuv422_fast        one launch processes a batch of fields IN PLACE.  One CTA = one group of 31 consecutive
//                        rows (+ a halo lane) = four warps, one per role; a lane streams its row through the
//                        stages of its role, the roles talk through the group's byte rings in shared memory and
//                        meet at a barrier after every step.  HBM traffic = one read + one write of the field's
//                        rows (2 bytes per pixel each way).  Takes rows of whole blocks (w % 8 == 0) with the
//                        common switches on aligned planes.
//   k_yuv422             the general kernel (same mapping, the general variant of every stage): everything else.
//   k_yuv422_halo        copies the one row per group that a halo lane re-reads (the last row of the group
//                        above) before the in-place pass overwrites it.
//   k_yuv422_headswitch  pre-pass for head-switch rotations that are not a short delay.
//   k_render_field       render_field(): vertical 8.8 resampling of a source picture onto field rows.
#ifndef CVS_YUV422_KERNELS_CUH
#define CVS_YUV422_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "yuv422_pipeline.cuh"

namespace cvs422 {

#ifndef CVS422_MIN_CTAS
#define CVS422_MIN_CTAS 5
#endif
constexpr int kSmSlots = 256;            // >= SMs of the device (148 on B200)
// One CTA = one group of 31 consecutive rows (+ the halo lane) = kRoles warps: warp r runs role r (yuv422_pipeline.cuh,
// "roles") for all rows of the group.  Why (measured on B200, profiles/ab_variants_r2.txt section 6, profiles/probes_r2.txt):
// a lane that runs every stage of its row needs ~210 registers (35 doubles of filter state), i.e. 2 warps per
// scheduler, each an in-order stream of dependent 8.65-cycle FP64 operations and 4-cycle integer chains, and a
// 37 KB loop that does not fit the SM's 32 KB instruction cache (issue slots 52 % busy whatever else was tried).
// Split by role a warp needs ~95 registers (5 groups = 20 warps per SM) and the four loops together are 25 KB.
constexpr int kNT = 32 * kRoles;         // threads per CTA
constexpr int kRowsPerWarp = 31;         // rows per group; lane 0 is the halo row
constexpr int kStrideY = kRingY + 4;     // per-lane ring strides: +1 word so that equal offsets of the 32 lanes
constexpr int kStrideC = kRingC + 4;     // fall into 32 different banks
constexpr int kStrideA = kRingA + 4;

struct FieldDesc422 {
    uint8_t *y, *u, *v;                  // picture planes (device), processed in place
    const uint32_t *rowinfo;             // nl packed row records
    const uint32_t *seek;                // nl * 62 jump-polynomial words of this geometry / parity
    uint8_t *hs_scratch;                 // hs_count rotated composite rows of w bytes
    const int32_t *hs_shift;
    uint8_t *halo;                       // warps_per_field halo records (halo_pitch bytes each)
    unsigned long long fieldno;
    int32_t field, nl, hs_first, hs_count;
    int32_t row_start, pad_;             // rows of the preceding fields of the batch (packed mapping)
    uint32_t window[64];
};

struct HsItem422 {
    int32_t field_idx, slot;
};

struct Launch422 {
    K422 K;
    DivPair dv;
    const FieldDesc422 *fields;
    int32_t nfields, warps_per_field, total_warps;
    // packed mapping (as in scanline_kernels.cuh): the rows of all fields of the batch form one sequence cut into
    // warps of 31; halo records are then indexed by the global warp (halo_base), not by (field, warp of the field)
    int32_t packed, total_rows, max_nl;
    uint8_t *halo_base;
    int32_t ly, lu, lv;                  // linesizes
    long long by, bu, bv;                // plane sizes in bytes (linesize * h): reads past them return 0
    int32_t halo_pitch, halo_u, halo_v;  // halo record: [Y: w + 2 bytes][U at halo_u][V at halo_v]
    int32_t vec;                         // planes, linesizes and picture strides allow 8 / 4 byte accesses
    int32_t *status;
    int32_t *sm_slots;                   // kSmSlots arrival counters, one per SM (role_of)
    int32_t rotate;
};

// Dynamic shared memory of k_yuv422: the group's rings, then one generator ring per noise stream that is ON
struct Smem422 {
    static constexpr size_t rng1 = (size_t)kRngSlots * 32 * sizeof(uint32_t);             // one stream
    static constexpr size_t wins = (size_t)2 * 128 * sizeof(uint32_t);        // per stream; two fields can meet in a group
    static constexpr size_t ry = (size_t)32 * kStrideY, rya = (size_t)32 * kStrideA;
    static constexpr size_t rc = (size_t)32 * kStrideC;
    static constexpr size_t rcomb = (size_t)32 * 3 * kMaxRecombine * sizeof(int32_t);
    static constexpr size_t off_wins = 0, off_ry = off_wins + wins, off_rya = off_ry + ry, off_ru = off_rya + rya,
                            off_rv = off_ru + rc, off_rcomb = off_rv + rc, off_rng = off_rcomb + rcomb;
    static constexpr size_t total_max = off_rng + 2 * rng1 + (size_t)(2 * kPhaseMapMax + 1) * 512;
    static CVS_HD size_t off_rng_chroma(const K422 &K) { return off_rng + (K.vnoise != 0 ? rng1 : 0); }
    static CVS_HD size_t off_pmap(const K422 &K) { return off_rng_chroma(K) + (K.cnoise != 0 ? rng1 : 0); }
    static CVS_HD size_t total(const K422 &K) { return off_pmap(K) + phase_maps_bytes(K); }   // (only the fast kernel uses the maps)
};

// what a lane reads its row from
struct LaneSrc {
    const uint8_t *y, *u, *v;
    int y_avail;                         // readable luma bytes from y (<= w + 2; the rest reads as 0)
};

// ---- render_field (:1001-1129) ------------------------------------------------------------------------
struct RenderArgs {
    uint8_t *dst[3];
    const uint8_t *src[3];
    int32_t dst_ls[3], src_ls[3], row_bytes[3];
    int32_t dst_h, src_h, is420, interlaced, tff, second, field;
};

// source rows and 8-bit fraction for output row y: luma (sy, sy2, syf) and chroma (csy, csy2, csyf)
struct RowMap {
    unsigned sy, sy2, syf, csy, csy2, csyf;
};
CVS_HD RowMap render_row_map(unsigned y, int dst_h, int src_h, int is420, int interlaced, int tff, int second) {
    RowMap m;
    const unsigned chroma_h = is420 ? (unsigned)src_h >> 1 : (unsigned)src_h;
    unsigned sy = (y * 0x100u * (unsigned)src_h) / (unsigned)dst_h;
    unsigned syf = sy & 0xFFu;
    sy >>= 8;
    unsigned csy = sy, csyf = syf;
    if (is420) { if (!(csy & 1u)) csyf = 0; csy >>= 1; }
    if (interlaced) {
        unsigned which = tff ? 0u : 1u;
        if (second) which ^= 1u;
        if (which == 0) {
            if (sy & 1u) { sy++; syf = 0; }
            if (csy & 1u) { csy++; csyf = 0; }
        } else {
            if (!(sy & 1u)) { sy++; syf = 0; }
            if (!(csy & 1u)) { csy++; csyf = 0; }
        }
        if (sy >= (unsigned)(src_h - 2)) { sy = (unsigned)(src_h - 2); syf = 0; }
        m.sy2 = sy + 2;
        if (csy >= chroma_h - 2) { csy = chroma_h - 2; csyf = 0; }
        m.csy2 = csy + 1;
    } else {
        if (sy >= (unsigned)(src_h - 1)) { sy = (unsigned)(src_h - 1); syf = 0; }
        m.sy2 = sy + 1;
        if (csy >= chroma_h - 1) { csy = chroma_h - 1; csyf = 0; }
        m.csy2 = csy + 1;
    }
    m.sy = sy; m.syf = syf; m.csy = csy; m.csyf = csyf;
    return m;
}

#ifdef CVS_YUV422_DEFINE_KERNELS   // device code: compiled once, in yuv422_kernels.cu

__device__ __forceinline__ void rebase422(const uint32_t *win_smem, const uint32_t *poly_g, uint32_t hist[31]) {
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

__device__ __forceinline__ uint32_t gather4(const uint8_t *p, int n_valid) {
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < n_valid) w |= (uint32_t)p[k] << (8 * k);
    return w;
}

template <bool EDGE>
__device__ __forceinline__ void load_block(const LaneSrc &src, const K422 &K, int s, bool vec, StepIO &io) {
    const int x0 = s * kB, c0 = s * kBC;
    if (EDGE) {
        io.y0 = gather4(src.y + x0, src.y_avail - x0);
        io.y1 = gather4(src.y + x0 + 4, src.y_avail - x0 - 4);
        io.u = gather4(src.u + c0, K.cw - c0);
        io.v = gather4(src.v + c0, K.cw - c0);
    } else if (vec) {
        const uint2 yy = *reinterpret_cast<const uint2 *>(src.y + x0);
        io.y0 = yy.x; io.y1 = yy.y;
        io.u = *reinterpret_cast<const uint32_t *>(src.u + c0);
        io.v = *reinterpret_cast<const uint32_t *>(src.v + c0);
    } else {
        io.y0 = gather4(src.y + x0, 4);
        io.y1 = gather4(src.y + x0 + 4, 4);
        io.u = gather4(src.u + c0, 4);
        io.v = gather4(src.v + c0, 4);
    }
}

template <bool EDGE>
__device__ __forceinline__ void store_block(uint8_t *y, uint8_t *u, uint8_t *v, const K422 &K, int bs, bool vec, const StepIO &o) {
    const int x0 = bs * kB, c0 = bs * kBC;
    if (vec && (!EDGE || x0 + kB <= K.w)) {
        *reinterpret_cast<uint2 *>(y + x0) = make_uint2(o.y0, o.y1);
        *reinterpret_cast<uint32_t *>(u + c0) = o.u;
        *reinterpret_cast<uint32_t *>(v + c0) = o.v;
    } else {
#pragma unroll
        for (int j = 0; j < kB; j++)
            if (x0 + j < K.w) y[x0 + j] = (uint8_t)(((j < 4 ? o.y0 : o.y1) >> (8 * (j & 3))) & 0xFFu);
#pragma unroll
        for (int k = 0; k < kBC; k++)
            if (c0 + k < K.cw) {
                u[c0 + k] = (uint8_t)((o.u >> (8 * k)) & 0xFFu);
                v[c0 + k] = (uint8_t)((o.v >> (8 * k)) & 0xFFu);
            }
    }
}

__device__ __forceinline__ void load_block_vec(const LaneSrc &src, int s, StepIO &io) {
    const uint2 yy = *reinterpret_cast<const uint2 *>(src.y + s * kB);
    io.y0 = yy.x; io.y1 = yy.y;
    io.u = *reinterpret_cast<const uint32_t *>(src.u + s * kBC);
    io.v = *reinterpret_cast<const uint32_t *>(src.v + s * kBC);
}

#ifndef CVS422_NO_BARRIER
#define CVS422_NO_BARRIER 0              // 1: TIMING EXPERIMENT ONLY (wrong pictures): how much the per-step barrier costs
#endif
__device__ __forceinline__ void group_barrier() {
#if !CVS422_NO_BARRIER
    asm volatile("bar.sync 0;" ::: "memory");
#endif
}

// Which role a warp of this group runs: role = warp, unless the launch asks for a rotation (CVS422_ROTATE=1, an experiment
// that is kept because its result is the opposite of what the arithmetic suggests).  Warp k of a CTA sits on scheduler k
// of its SM and the roles are not equally long (role 2 issues 276 FP64 instructions per step, the others 123; the FP64
// pipe takes one warp instruction per two cycles and scheduler), so with role = warp the role-2 warps of all five
// resident groups share ONE scheduler, whose FP64 pipe is ~76 % busy while the SM's average is 48 %.  Mixing the roles
// over the schedulers should even that out.  The block index cannot do it -- blocks are dealt to the 148 SMs round robin
// and 148 is a multiple of 4, so (warp + blockIdx) % 4 gives every group of an SM the same assignment, which is why the
// first experiment with it showed no difference -- so a group draws the next number of ITS SM.  Measured (B200, 1080p
// -vhs -vhs-speed sp, profiles/ab_variants_r2.txt section 7): 83.2 k fields/s rotated against 90.2 k with role = warp.
// One role per scheduler keeps that scheduler's instruction stream to ONE loop of ~6 KB (each scheduler fetches through
// its own small L0 instruction cache in front of the SM's 32 KB); four loops per scheduler cost more than the idle FP64
// cycles they win back.  Re-dealing a stage instead (the luma sharpen from role 2 to role 0, CVS422_LSHARP_ROLE) was
// measured too and is slower as well (83.1 k): the FP64 count of the longest role is not what a step waits for either;
// DESIGN.md section 9.
__device__ __forceinline__ int role_of(const Launch422 &a, int tid) {
    __shared__ int s_rot;
    if (!a.rotate) return tid >> 5;
    if (tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_rot = atomicAdd(a.sm_slots + (smid % kSmSlots), 1);
    }
    __syncthreads();
    return ((tid >> 5) + s_rot) % kRoles;
}

// The general kernel: any width, any switch, pre-pass rows; every step is the general variant of its role
// (yuv422_pipeline.cuh, "roles").  Rows the fast kernel below can take never come here (launch_yuv422).
__global__ void __launch_bounds__(kNT, CVS422_MIN_CTAS) k_yuv422(const __grid_constant__ Launch422 a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int gw = blockIdx.x;            // the group of rows
    const int role = role_of(a, tid);
    const K422 &K = a.K;
    const int w = K.w;
    // which field row this lane computes (lane 0 = the halo row: the row above lane 1's)
    int fi, row;
    bool valid;
    const uint8_t *halo_rec;             // this group's halo record
    if (a.packed) {
        int g = kRowsPerWarp * gw + lane - 1;
        valid = (lane >= 1) && g < a.total_rows;
        g = g < 0 ? 0 : (g > a.total_rows - 1 ? a.total_rows - 1 : g);
        fi = g / a.max_nl;                                    // never beyond the row's field
        while (g >= a.fields[fi].row_start + a.fields[fi].nl) fi++;
        row = g - a.fields[fi].row_start;
        const int fi1 = __shfl_sync(0xffffffffu, fi, 1);
        if (lane == 0 && fi != fi1) { fi = fi1; row = 0; }    // lane 1 is row 0 of its field: no halo needed
        halo_rec = a.halo_base + (size_t)gw * (size_t)a.halo_pitch;
    } else {
        fi = gw / a.warps_per_field;
        const int wp = gw - fi * a.warps_per_field, nlw = a.fields[fi].nl;
        if (kRowsPerWarp * wp >= nlw) return;                 // (the whole group)
        row = kRowsPerWarp * wp + lane - 1;
        valid = (lane >= 1) && row < nlw;
        row = row < 0 ? 0 : (row > nlw - 1 ? nlw - 1 : row);
        halo_rec = a.fields[fi].halo + (size_t)wp * (size_t)a.halo_pitch;
    }
    const FieldDesc422 &fd = a.fields[fi];
    // lane 0 re-reads the row above lane 1's row from the copy k_yuv422_halo took, unless lane 1 is row 0
    const bool halo_copy = __shfl_sync(0xffffffffu, row, 1) >= 1;
    const long long y = (long long)fd.field + 2 * row;

    Lane422 ln;
    ln.reset();
    ln.ry = smem + Smem422::off_ry + (size_t)lane * kStrideY;
    ln.rya = smem + Smem422::off_rya + (size_t)lane * kStrideA;
    ln.ru = smem + Smem422::off_ru + (size_t)lane * kStrideC;
    ln.rv = smem + Smem422::off_rv + (size_t)lane * kStrideC;
    ln.rcomb = reinterpret_cast<int32_t *>(smem + Smem422::off_rcomb) + (size_t)lane * 3 * kMaxRecombine;

    const Lags L = lags_of(K);
    const Geo G = geo_of(K);
    const int nsteps = line_steps(K);
    Row422 rc;
    row_setup(K, (unsigned)fd.field, fd.fieldno, row, __ldg(fd.rowinfo + row), rc);

    // the generator of a noise stream lives in the role that draws from it: luma in role 0, chroma in role 1
    if ((role == 0 && K.vnoise != 0) || (role == 1 && K.cnoise != 0)) {
        // generator windows of the (at most two) fields of this group: [0,64) the field of lane 1, [64,128) that of lane 31
        uint32_t *win = reinterpret_cast<uint32_t *>(smem + Smem422::off_wins) + role * 128;
        const int fiA = __shfl_sync(0xffffffffu, fi, 1), fiB = __shfl_sync(0xffffffffu, fi, 31);
        win[lane] = a.fields[fiA].window[lane];
        win[lane + 32] = a.fields[fiA].window[lane + 32];
        win[lane + 64] = a.fields[fiB].window[lane];
        win[lane + 96] = a.fields[fiB].window[lane + 32];
        __syncwarp();
        if (fi != fiA) win += 64;
        uint32_t hist[31];
        bool ok;
        if (role == 0) {
            const long long pre = (long long)row * w;
            const int nd = (int)(pre < kWarm ? pre : kWarm);
            rebase422(win, fd.seek + (size_t)row * 62, hist);
            ln.rngL.init(reinterpret_cast<uint32_t *>(smem + Smem422::off_rng) + lane, 32, hist, kRngBase - (uint32_t)nd);
            ok = cvs::warm_luma((uint32_t)(2 * K.vnoise + 1), K.vmagic, K.vshift, K.vnoise, ln.rngL, nd, pre <= kWarm, ln.nY);
        } else {
            const long long pre = (long long)row * K.cw;
            const int nd = (int)(pre < kWarm ? pre : kWarm);
            rebase422(win, fd.seek + (size_t)row * 62 + 31, hist);
            ln.rngC.init(reinterpret_cast<uint32_t *>(smem + Smem422::off_rng_chroma(K)) + lane, 32, hist, kRngBase - 2u * (uint32_t)nd);
            ok = cvs::warm_chroma((uint32_t)(2 * K.cnoise + 1), K.cmagic, K.cshift, K.cnoise, ln.rngC, nd, pre <= kWarm, ln.nU, ln.nV);
        }
        if (!ok) atomicOr(a.status, 1);
    }

    if (role == 0) {
        // rows: lanes 1..31 read the picture; the halo lane of groups 1.. reads the copy k_yuv422_halo took
        LaneSrc src;
        if (lane == 0 && halo_copy) {
            const uint8_t *hr = halo_rec;
            src.y = hr; src.u = hr + a.halo_u; src.v = hr + a.halo_v;
            src.y_avail = w + 2;
        } else {
            src.y = fd.y + y * a.ly; src.u = fd.u + y * a.lu; src.v = fd.v + y * a.lv;
            const long long left = a.by - y * a.ly;            // bytes of the plane from the row start
            src.y_avail = (int)(left < (long long)(w + 2) ? left : (long long)(w + 2));
        }
        const uint8_t *hsrow = (rc.rflags & RG_HEADSW_PRE) ? fd.hs_scratch + (size_t)(row - fd.hs_first) * (size_t)w : nullptr;
        const bool warp_hs = __any_sync(0xffffffffu, rc.hs_delay > 0);
        const bool vec = a.vec != 0;      // (halo records are 16-byte aligned, so the halo lane qualifies too)
        StepIO cur, nxt;
        load_block<true>(src, K, 0, vec, cur);
#pragma unroll 1
        for (int s = 0; s < nsteps; s++) {
            load_block<true>(src, K, s + 1, vec, nxt);
            role0_step(K, L, G, rc, ln, s, cur, warp_hs, hsrow);
            cur = nxt;
            group_barrier();
        }
    } else if (role == 1) {
#pragma unroll 1
        for (int s = 0; s < nsteps; s++) {
            role1_step(K, L, G, a.dv, rc, ln, s);
            group_barrier();
        }
    } else if (role == 2) {
#pragma unroll 1
        for (int s = 0; s < nsteps; s++) {
            uint32_t pu, pv;
            role2_front(K, L, G, ln, s, pu, pv);
            const uint32_t au = __shfl_up_sync(0xffffffffu, pu, 1), av = __shfl_up_sync(0xffffffffu, pv, 1);
            role2_back(K, L, G, rc, ln, s, pu, pv, au, av);
            group_barrier();
        }
    } else {
        for (int i = 0; i < 3 * kMaxRecombine; i++) ln.rcomb[i] = 16;
        uint8_t *dy = fd.y + y * a.ly, *du = fd.u + y * a.lu, *dvp = fd.v + y * a.lv;
        const bool vec = a.vec != 0;
#pragma unroll 1
        for (int s = 0; s < nsteps; s++) {
            StepIO o;
            int bs;
            if (role3_step(K, L, G, a.dv, rc, ln, s, o, bs) && valid) store_block<true>(dy, du, dvp, K, bs, vec, o);
            group_barrier();
        }
    }
}

// The fast kernel: rows of whole blocks with the common switches (fast_row_ok), aligned planes, no pre-pass rows.
// Same mapping as k_yuv422 (one CTA = one group of 31 rows + halo lane = four role warps), but a single code path per
// role -- no general variant, so the loops of the four roles (~25 KB) are all an SM ever executes.
__global__ void __launch_bounds__(kNT, CVS422_MIN_CTAS) k_yuv422_fast(const __grid_constant__ Launch422 a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int gw = blockIdx.x;
    const int role = role_of(a, tid);         // (the four loops together fit the SM's instruction cache: the mix costs nothing there)
    const K422 &K = a.K;
    const int w = K.w;
    int fi, row;
    bool valid;
    const uint8_t *halo_rec;
    if (a.packed) {
        int g = kRowsPerWarp * gw + lane - 1;
        valid = (lane >= 1) && g < a.total_rows;
        g = g < 0 ? 0 : (g > a.total_rows - 1 ? a.total_rows - 1 : g);
        fi = g / a.max_nl;
        while (g >= a.fields[fi].row_start + a.fields[fi].nl) fi++;
        row = g - a.fields[fi].row_start;
        const int fi1 = __shfl_sync(0xffffffffu, fi, 1);
        if (lane == 0 && fi != fi1) { fi = fi1; row = 0; }
        halo_rec = a.halo_base + (size_t)gw * (size_t)a.halo_pitch;
    } else {
        fi = gw / a.warps_per_field;
        const int wp = gw - fi * a.warps_per_field, nlw = a.fields[fi].nl;
        if (kRowsPerWarp * wp >= nlw) return;
        row = kRowsPerWarp * wp + lane - 1;
        valid = (lane >= 1) && row < nlw;
        row = row < 0 ? 0 : (row > nlw - 1 ? nlw - 1 : row);
        halo_rec = a.fields[fi].halo + (size_t)wp * (size_t)a.halo_pitch;
    }
    const FieldDesc422 &fd = a.fields[fi];
    const bool halo_copy = __shfl_sync(0xffffffffu, row, 1) >= 1;
    const long long y = (long long)fd.field + 2 * row;

    Lane422 ln;
    ln.reset();
    ln.ry = smem + Smem422::off_ry + (size_t)lane * kStrideY;
    ln.rya = smem + Smem422::off_rya + (size_t)lane * kStrideA;
    ln.ru = smem + Smem422::off_ru + (size_t)lane * kStrideC;
    ln.rv = smem + Smem422::off_rv + (size_t)lane * kStrideC;
    ln.rcomb = nullptr;
    Fast422::prime(K, ln);

    const Lags L = lags_of(K);
    const int nb = w / kB, nsteps = line_steps(K);
    Row422 rc;
    row_setup(K, (unsigned)fd.field, fd.fieldno, row, __ldg(fd.rowinfo + row), rc);
    if (K.flags & G_PHASE_MAP) {          // the rotation tables into shared memory (first used after the first barrier)
        uint8_t *maps = smem + Smem422::off_pmap(K);
        const uint4 *from = reinterpret_cast<const uint4 *>(phase_maps(K));
        for (int i = tid; i < (int)(phase_maps_bytes(K) / 16); i += kNT) reinterpret_cast<uint4 *>(maps)[i] = __ldg(from + i);
        rc.pmap = row_phase_map(K, maps, __ldg(fd.rowinfo + row));
    }

    if ((role == 0 && K.vnoise != 0) || (role == 1 && K.cnoise != 0)) {
        uint32_t *win = reinterpret_cast<uint32_t *>(smem + Smem422::off_wins) + role * 128;
        const int fiA = __shfl_sync(0xffffffffu, fi, 1), fiB = __shfl_sync(0xffffffffu, fi, 31);
        win[lane] = a.fields[fiA].window[lane];
        win[lane + 32] = a.fields[fiA].window[lane + 32];
        win[lane + 64] = a.fields[fiB].window[lane];
        win[lane + 96] = a.fields[fiB].window[lane + 32];
        __syncwarp();
        if (fi != fiA) win += 64;
        uint32_t hist[31];
        bool ok;
        if (role == 0) {
            const long long pre = (long long)row * w;
            const int nd = (int)(pre < kWarm ? pre : kWarm);
            rebase422(win, fd.seek + (size_t)row * 62, hist);
            ln.rngL.init(reinterpret_cast<uint32_t *>(smem + Smem422::off_rng) + lane, 32, hist, kRngBase - (uint32_t)nd);
            ok = cvs::warm_luma((uint32_t)(2 * K.vnoise + 1), K.vmagic, K.vshift, K.vnoise, ln.rngL, nd, pre <= kWarm, ln.nY);
        } else {
            const long long pre = (long long)row * K.cw;
            const int nd = (int)(pre < kWarm ? pre : kWarm);
            rebase422(win, fd.seek + (size_t)row * 62 + 31, hist);
            ln.rngC.init(reinterpret_cast<uint32_t *>(smem + Smem422::off_rng_chroma(K)) + lane, 32, hist, kRngBase - 2u * (uint32_t)nd);
            ok = cvs::warm_chroma((uint32_t)(2 * K.cnoise + 1), K.cmagic, K.cshift, K.cnoise, ln.rngC, nd, pre <= kWarm, ln.nU, ln.nV);
        }
        if (!ok) atomicOr(a.status, 1);
    }

    if (role == 0) {
        LaneSrc src;
        if (lane == 0 && halo_copy) {
            src.y = halo_rec; src.u = halo_rec + a.halo_u; src.v = halo_rec + a.halo_v;
            src.y_avail = w + 2;
        } else {
            src.y = fd.y + y * a.ly; src.u = fd.u + y * a.lu; src.v = fd.v + y * a.lv;
            const long long left = a.by - y * a.ly;
            src.y_avail = (int)(left < (long long)(w + 2) ? left : (long long)(w + 2));
        }
        const bool warp_hs = __any_sync(0xffffffffu, rc.hs_delay > 0);
        StepIO cur, nxt;
        load_block_vec(src, 0, cur);
#pragma unroll 1
        for (int s = 0; s < nsteps; s++) {
            nxt.y0 = nxt.y1 = nxt.u = nxt.v = 0;
            if (s + 1 < nb) load_block_vec(src, s + 1, nxt);
            else if (s + 1 == nb) nxt.y0 = gather4(src.y + w, src.y_avail - w);       // the two bytes past the row
            frow0_step(K, L, nb, rc, ln, s, cur, warp_hs);
            cur = nxt;
            group_barrier();
        }
    } else if (role == 1) {
#pragma unroll 1
        for (int s = 0; s < nsteps; s++) {
            frow1_step(K, L, nb, rc, ln, s);
            group_barrier();
        }
    } else if (role == 2) {
#pragma unroll 1
        for (int s = 0; s < nsteps; s++) {
            uint32_t pu, pv;
            frow2_front(K, L, nb, ln, s, pu, pv);
            const uint32_t au = __shfl_up_sync(0xffffffffu, pu, 1), av = __shfl_up_sync(0xffffffffu, pv, 1);
            frow2_back(K, L, nb, rc, ln, s, pu, pv, au, av);
            group_barrier();
        }
    } else {
        uint8_t *dy = fd.y + y * a.ly, *du = fd.u + y * a.lu, *dvp = fd.v + y * a.lv;
#pragma unroll 1
        for (int s = 0; s < nsteps; s++) {
            StepIO o;
            int bs;
            if (frow3_step(K, L, nb, rc, ln, s, o, bs) && valid) {
                *reinterpret_cast<uint2 *>(dy + bs * kB) = make_uint2(o.y0, o.y1);
                *reinterpret_cast<uint32_t *>(du + bs * kBC) = o.u;
                *reinterpret_cast<uint32_t *>(dvp + bs * kBC) = o.v;
            }
            group_barrier();
        }
    }
}

// one CTA per warp of the main pass: copy the row above the warp's first row (unless that is row 0 of a field)
__global__ void __launch_bounds__(128) k_yuv422_halo(const __grid_constant__ Launch422 a) {
    int fi, row;
    uint8_t *hr;
    if (a.packed) {
        const int gw = blockIdx.x, g1 = kRowsPerWarp * gw;     // lane 1's position in the batch's sequence of rows
        if (g1 >= a.total_rows) return;
        fi = g1 / a.max_nl;
        while (g1 >= a.fields[fi].row_start + a.fields[fi].nl) fi++;
        row = g1 - a.fields[fi].row_start - 1;
        if (row < 0) return;
        hr = a.halo_base + (size_t)gw * (size_t)a.halo_pitch;
    } else {
        fi = blockIdx.x / a.warps_per_field;
        const int wp = blockIdx.x - fi * a.warps_per_field;
        if (wp == 0 || kRowsPerWarp * wp >= a.fields[fi].nl) return;
        row = kRowsPerWarp * wp - 1;
        hr = a.fields[fi].halo + (size_t)wp * (size_t)a.halo_pitch;
    }
    const FieldDesc422 &fd = a.fields[fi];
    const long long y = (long long)fd.field + 2 * row;
    const int w = a.K.w, cw = a.K.cw;
    const long long left = a.by - y * a.ly;
    for (int x = threadIdx.x; x < w + 2; x += blockDim.x) hr[x] = (x < left) ? fd.y[y * a.ly + x] : (uint8_t)0;
    for (int c = threadIdx.x; c < cw; c += blockDim.x) {
        hr[a.halo_u + c] = fd.u[y * a.lu + c];
        hr[a.halo_v + c] = fd.v[y * a.lv + c];
    }
}

constexpr int kHsNT = 32;

__global__ void __launch_bounds__(kHsNT) k_yuv422_headswitch(const __grid_constant__ Launch422 a,
                                                            const HsItem422 *__restrict__ items, int nitems) {
    __shared__ uint32_t ring[kRngSlots * kHsNT];
    __shared__ __align__(4) uint8_t rb[kHsNT * (kStrideY + kStrideA + 2 * kStrideC)];
    const int it = blockIdx.x * kHsNT + threadIdx.x;
    if (it >= nitems) return;
    const HsItem422 item = items[it];
    const FieldDesc422 &fd = a.fields[item.field_idx];
    const K422 &K = a.K;
    const int w = K.w;
    const int row = fd.hs_first + item.slot;
    Lane422 ln;
    ln.reset();
    uint8_t *base = rb + (size_t)threadIdx.x * (kStrideY + kStrideA + 2 * kStrideC);
    ln.ry = base; ln.rya = base + kStrideY; ln.ru = base + kStrideY + kStrideA; ln.rv = base + kStrideY + kStrideA + kStrideC;
    ln.rcomb = nullptr;
    Row422 rc;
    row_setup(K, (unsigned)fd.field, fd.fieldno, row, __ldg(fd.rowinfo + row), rc);
    if (K.vnoise != 0) {
        const long long pre = (long long)row * w;
        const int nd = (int)(pre < kWarm ? pre : kWarm);
        uint32_t hist[31];
        cvs::rng_rebase(fd.window, fd.seek + (size_t)row * 62, hist);
        ln.rngL.init(ring + threadIdx.x, kHsNT, hist, kRngBase - (uint32_t)nd);
        if (!cvs::warm_luma((uint32_t)(2 * K.vnoise + 1), K.vmagic, K.vshift, K.vnoise, ln.rngL, nd, pre <= kWarm, ln.nY))
            atomicOr(a.status, 1);
    }
    const long long y = (long long)fd.field + 2 * row;
    headswitch_row(K, rc, ln, fd.y + y * a.ly, fd.u + y * a.lu, fd.v + y * a.lv,
                   fd.hs_scratch + (size_t)item.slot * (size_t)w, __ldg(fd.hs_shift + item.slot));
}

// grid: (ceil(max row_bytes / 256), field rows, 3 planes); one byte per thread, coalesced along the row
__global__ void __launch_bounds__(256) k_render_field(const __grid_constant__ RenderArgs a) {
    const int p = blockIdx.z;
    const unsigned y = (unsigned)a.field + 2u * blockIdx.y;
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (y >= (unsigned)a.dst_h || x >= a.row_bytes[p]) return;
    const RowMap m = render_row_map(y, a.dst_h, a.src_h, a.is420, a.interlaced, a.tff, a.second);
    const bool chroma_map = a.is420 && p > 0;            // 4:2:2 sources use the luma mapping for all planes
    const unsigned r1 = chroma_map ? m.csy : m.sy, r2 = chroma_map ? m.csy2 : m.sy2, fr = chroma_map ? m.csyf : m.syf;
    const int s1 = a.src[p][(size_t)a.src_ls[p] * r1 + x];
    int v = s1;
    if (fr != 0) {
        const int s2 = a.src[p][(size_t)a.src_ls[p] * r2 + x];
        v = s1 + (int)(uint8_t)(((s2 - s1) * (int)fr) >> 8);
    }
    a.dst[p][(size_t)a.dst_ls[p] * y + x] = (uint8_t)v;
}

#endif  // CVS_YUV422_DEFINE_KERNELS

cudaError_t launch_yuv422(const Launch422 &a, const HsItem422 *d_items, int nitems, cudaStream_t st);
cudaError_t occupancy_yuv422(const K422 &K, int *ctas_per_sm);
cudaError_t launch_render_field(const RenderArgs &a, cudaStream_t st);

}  // namespace cvs422
#endif
