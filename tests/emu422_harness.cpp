// emu422_harness.cpp -- TEST INFRASTRUCTURE: runs composite_video_simulator_b200/csrc/yuv422_pipeline.cuh
// on the CPU exactly as the kernel k_yuv422 runs it on the GPU: warps of 32 lanes in lock-step (lane 0
// is the halo row of the vertical chroma blend), per-lane byte rings, the jump-ahead planner, the noise
// warm-up.  Lets the CPU test-suite check the whole GPU algorithm against the oracle without a GPU.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../composite_video_simulator_b200/csrc/yuv422_plan.h"

using namespace cvs422;

namespace {

struct LaneMem {
    alignas(4) uint8_t ry[kRingY], ru[kRingC], rv[kRingC], rya[kRingA];
    uint32_t ringL[cvs::kRngSlots], ringC[cvs::kRngSlots];
    int32_t rcomb[3 * kMaxRecombine];
};

// block b of a plane row as the kernel's loader delivers it: bytes inside the plane, 0 outside
uint32_t load_word(const uint8_t *plane, long long plane_bytes, long long off, int n_valid) {
    uint32_t w = 0;
    for (int k = 0; k < 4; k++) {
        const long long o = off + k;
        if (k < n_valid && o >= 0 && o < plane_bytes) w |= (uint32_t)plane[o] << (8 * k);
    }
    return w;
}

}  // namespace

extern "C" {

// force_general: bit 0: 0 = the fast kernel where it applies, 1 = the general kernel everywhere;
// bit 1: the roles of a step in forward instead of reverse order (they must not depend on each other)
int emu422_process(const cvs422_params *p, unsigned long long rng_pos,
                   uint8_t *yp, int ly, uint8_t *up, int lu, uint8_t *vp, int lv,
                   int w, int h, unsigned field, unsigned long long fieldno, int force_general,
                   unsigned long long *draws_out) {
    K422 K;
    DivPair dv;
    std::vector<double> lut;
    int rc0 = make_k422(*p, w, h, K, dv, lut);
    if (rc0 != CVS_OK) return rc0;
    cvs::GeomPlan g;
    build_geom_plan(*p, w, h, field, g);
    cvs::RandCursor cur;
    cur.seed(1);
    cur.seek(rng_pos);
    cvs::FieldSide fs;
    build_field_side_at(*p, g, cur, fs);
    if (draws_out) *draws_out = g.ndraws;

    // the kernel reads the halo rows from a copy taken before the launch; here: copy everything
    std::vector<uint8_t> sy(yp, yp + (size_t)ly * h), su(up, up + (size_t)lu * h), sv(vp, vp + (size_t)lv * h);
    const long long by = (long long)ly * h, bu = (long long)lu * h, bv = (long long)lv * h;

    // head-switch pre-pass: one lane per rotated row, as k_yuv422_headswitch does
    std::vector<uint8_t> scratch((size_t)(fs.hs_count > 0 ? fs.hs_count : 1) * (size_t)w);
    for (int i = 0; i < fs.hs_count; i++) {
        const int row = fs.hs_first + i;
        LaneMem m;
        std::memset(&m, 0, sizeof(m));
        Lane422 L;
        L.reset();
        L.ry = m.ry; L.ru = m.ru; L.rv = m.rv; L.rya = m.rya; L.rcomb = m.rcomb;
        Row422 rcw;
        row_setup(K, field, fieldno, row, fs.rowinfo[(size_t)row], rcw);
        if (K.vnoise != 0) {
            uint32_t hist[31];
            const long long pre = (long long)row * w;
            const int nd = warm_samples(pre);
            cvs::rng_rebase(fs.window, &g.seek[(size_t)row * 62], hist);
            L.rngL.init(m.ringL, 1, hist, kRngBase - (uint32_t)nd);
            if (!cvs::warm_luma((uint32_t)(2 * K.vnoise + 1), K.vmagic, K.vshift, K.vnoise, L.rngL, nd, pre <= kWarm, L.nY))
                return CVS_ERR_NOISE_SYNC;
        }
        const long long y = (long long)field + 2 * row;
        headswitch_row(K, rcw, L, yp + y * ly, up + y * lu, vp + y * lv, scratch.data() + (size_t)i * w, fs.hs_shift[(size_t)i]);
    }

    const int nl = g.nl, nsteps = line_steps(K);
    const Lags LG = lags_of(K);
    const Geo GE = geo_of(K);
    int status = CVS_OK;
    for (int wp = 0; wp * 31 < nl; wp++) {
        std::vector<LaneMem> mem(32);
        Lane422 ln[32];
        Row422 rc[32];
        bool valid[32];
        int rows[32];
        bool warp_hs = false;
        const uint8_t *hsrow[32];
        for (int lane = 0; lane < 32; lane++) {
            int row = 31 * wp + lane - 1;
            valid[lane] = lane >= 1 && row < nl;
            row = row < 0 ? 0 : (row > nl - 1 ? nl - 1 : row);
            rows[lane] = row;
            Lane422 &L = ln[lane];
            L.reset();
            L.ry = mem[lane].ry; L.ru = mem[lane].ru; L.rv = mem[lane].rv; L.rya = mem[lane].rya;
            L.rcomb = mem[lane].rcomb;
            std::memset(&mem[lane], 0, sizeof(LaneMem));
            row_setup(K, field, fieldno, row, fs.rowinfo[(size_t)row], rc[lane]);
            warp_hs |= rc[lane].hs_delay > 0;
            hsrow[lane] = (rc[lane].rflags & RG_HEADSW_PRE) ? scratch.data() + (size_t)(row - fs.hs_first) * w : nullptr;
            uint32_t hist[31];
            bool ok = true;
            if (K.vnoise != 0) {
                const long long pre = (long long)row * w;
                const int nd = warm_samples(pre);
                cvs::rng_rebase(fs.window, &g.seek[(size_t)row * 62], hist);
                L.rngL.init(mem[lane].ringL, 1, hist, kRngBase - (uint32_t)nd);
                ok &= cvs::warm_luma((uint32_t)(2 * K.vnoise + 1), K.vmagic, K.vshift, K.vnoise, L.rngL, nd, pre <= kWarm, L.nY);
            }
            if (K.cnoise != 0) {
                const long long pre = (long long)row * K.cw;
                const int nd = warm_samples(pre);
                cvs::rng_rebase(fs.window, &g.seek[(size_t)row * 62 + 31], hist);
                L.rngC.init(mem[lane].ringC, 1, hist, kRngBase - 2u * (uint32_t)nd);
                ok &= cvs::warm_chroma((uint32_t)(2 * K.cnoise + 1), K.cmagic, K.cshift, K.cnoise, L.rngC, nd, pre <= kWarm, L.nU, L.nV);
            }
            if (!ok) status = CVS_ERR_NOISE_SYNC;
            for (int i = 0; i < K.recombine; i++) { L.rcomb[3 * i] = 16; L.rcomb[3 * i + 1] = 16; L.rcomb[3 * i + 2] = 16; }
        }
        // the fast kernel (k_yuv422_fast): one code path per role, no general steps
        if (!(force_general & 1) && fast_row_ok(K) && fs.hs_count == 0) {
            const int nb = GE.nb;
            for (int lane = 0; lane < 32; lane++) {
                Fast422::prime(K, ln[lane]);
                if (K.flags & G_PHASE_MAP) rc[lane].pmap = row_phase_map(K, phase_maps(K), fs.rowinfo[(size_t)rows[lane]]);
            }
            for (int s = 0; s < nsteps; s++) {
                for (int ri = 0; ri < kRoles; ri++) {
                    const int role = (force_general & 2) ? ri : kRoles - 1 - ri;
                    if (role == 0) {
                        for (int lane = 0; lane < 32; lane++) {
                            const long long y = (long long)field + 2 * rows[lane];
                            StepIO in;
                            const int x0 = s * kB, c0 = s * kBC;
                            in.y0 = load_word(sy.data(), by, y * ly + x0, w + 2 - x0);
                            in.y1 = load_word(sy.data(), by, y * ly + x0 + 4, w + 2 - x0 - 4);
                            in.u = load_word(su.data(), bu, y * lu + c0, K.cw - c0);
                            in.v = load_word(sv.data(), bv, y * lv + c0, K.cw - c0);
                            frow0_step(K, LG, nb, rc[lane], ln[lane], s, in, warp_hs);
                        }
                    } else if (role == 1) {
                        for (int lane = 0; lane < 32; lane++) frow1_step(K, LG, nb, rc[lane], ln[lane], s);
                    } else if (role == 2) {
                        uint32_t pu[32], pv[32];
                        for (int lane = 0; lane < 32; lane++) frow2_front(K, LG, nb, ln[lane], s, pu[lane], pv[lane]);
                        for (int lane = 0; lane < 32; lane++)
                            frow2_back(K, LG, nb, rc[lane], ln[lane], s, pu[lane], pv[lane], lane ? pu[lane - 1] : 0, lane ? pv[lane - 1] : 0);
                    } else {
                        for (int lane = 0; lane < 32; lane++) {
                            StepIO out;
                            int bs;
                            if (frow3_step(K, LG, nb, rc[lane], ln[lane], s, out, bs) && valid[lane]) {
                                const long long y = (long long)field + 2 * rows[lane];
                                for (int j = 0; j < kB; j++) yp[y * ly + bs * kB + j] = (uint8_t)(((j < 4 ? out.y0 : out.y1) >> (8 * (j & 3))) & 0xFF);
                                for (int k = 0; k < kBC; k++) {
                                    up[y * lu + bs * kBC + k] = (uint8_t)((out.u >> (8 * k)) & 0xFF);
                                    vp[y * lv + bs * kBC + k] = (uint8_t)((out.v >> (8 * k)) & 0xFF);
                                }
                            }
                        }
                    }
                }
            }
            continue;
        }
        // The general kernel (k_yuv422).  The roles of a row run concurrently on the GPU and meet at a barrier after
        // every step, so within a step no role may depend on another one: run them in REVERSE order here.
        for (int s = 0; s < nsteps; s++) {
            for (int ri = 0; ri < kRoles; ri++) {
                const int role = (force_general & 2) ? ri : kRoles - 1 - ri;
                if (role == 0) {
                    for (int lane = 0; lane < 32; lane++) {
                        const long long y = (long long)field + 2 * rows[lane];
                        StepIO in;
                        const int x0 = s * kB, c0 = s * kBC;
                        // luma: bytes up to w+1 count (the reference's read past the row), chroma: cw samples
                        in.y0 = load_word(sy.data(), by, y * ly + x0, w + 2 - x0);
                        in.y1 = load_word(sy.data(), by, y * ly + x0 + 4, w + 2 - x0 - 4);
                        in.u = load_word(su.data(), bu, y * lu + c0, K.cw - c0);
                        in.v = load_word(sv.data(), bv, y * lv + c0, K.cw - c0);
                        role0_step(K, LG, GE, rc[lane], ln[lane], s, in, warp_hs, hsrow[lane]);
                    }
                } else if (role == 1) {
                    for (int lane = 0; lane < 32; lane++) role1_step(K, LG, GE, dv, rc[lane], ln[lane], s);
                } else if (role == 2) {
                    uint32_t pu[32], pv[32];
                    for (int lane = 0; lane < 32; lane++) role2_front(K, LG, GE, ln[lane], s, pu[lane], pv[lane]);
                    for (int lane = 0; lane < 32; lane++)
                        role2_back(K, LG, GE, rc[lane], ln[lane], s, pu[lane], pv[lane], lane ? pu[lane - 1] : 0, lane ? pv[lane - 1] : 0);
                } else {
                    for (int lane = 0; lane < 32; lane++) {
                        StepIO out;
                        int bs;
                        if (role3_step(K, LG, GE, dv, rc[lane], ln[lane], s, out, bs) && valid[lane]) {
                            const long long y = (long long)field + 2 * rows[lane];
                            for (int j = 0; j < kB; j++) {
                                const int x = bs * kB + j;
                                if (x < w) yp[y * ly + x] = (uint8_t)(((j < 4 ? out.y0 : out.y1) >> (8 * (j & 3))) & 0xFF);
                            }
                            for (int k = 0; k < kBC; k++) {
                                const int c = bs * kBC + k;
                                if (c < K.cw) {
                                    up[y * lu + c] = (uint8_t)((out.u >> (8 * k)) & 0xFF);
                                    vp[y * lv + c] = (uint8_t)((out.v >> (8 * k)) & 0xFF);
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    return status;
}

}  // extern "C"
