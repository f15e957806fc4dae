// cvs_api.cu -- the C ABI of include/cvs_ntsc.h on top of the scanline kernels.
//
// Host side of the drop-in boundary: owns the CUDA stream, the device-side tables and the
// position in the libc rand() stream (the reference's only hidden state across
// composite_layer() calls), plans each batch (field_plan.cpp) and
// launches k_headswitch + k_fields.  Nothing here computes pixels on the CPU: if CUDA is not
// usable every compute entry point fails with CVS_ERR_CUDA.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <thread>
#include <vector>

#include "../../include/cvs_ntsc.h"
#include "field_plan.h"
#include "glibc_rand.h"
#include "scanline_kernels.cuh"
#include "scale_convert.cuh"
#include "yuv_convert.cuh"
#include "sws_filter.h"

using namespace cvs;

namespace {

constexpr int kStagingSlots = 3;
constexpr size_t kMaxEventPairs = 4096;
constexpr int kHostChunkDefault = 32;       // fields per pipeline stage of the host-pointer entry points
                                            // (measured: 16 -> 8.6k, 32 -> 9.5k, 64 -> 9.6k fields/s at 1080p)

// device copy of the tap tables of one scaling geometry (scale_convert.cuh)
struct ScalePlan {
    int dw = 0, dh = 0, sw = 0, sh = 0, format = 0;
    int32_t *d_first = nullptr;      // fx | fy | cfx | cfy
    int16_t *d_w = nullptr;          // wx | wy | cwx | cwy
    size_t off_first[4] = {0, 0, 0, 0}, off_w[4] = {0, 0, 0, 0};
    int taps[4] = {0, 0, 0, 0};
    // the libswscale-exact path (k_sws_yuv_to_bgra): four banks in one allocation, pos then weights per bank
    bool sws = false, direct = false, full = false;
    bool rgb = false, rgb_half = false;          // BGRA source at another size through the library's route (k_sws_bgra_to_bgra)
    int32_t *d_sws = nullptr;
    size_t sws_pos[4] = {0, 0, 0, 0}, sws_w[4] = {0, 0, 0, 0};
    int sws_taps[4] = {0, 0, 0, 0};
};

struct DevPlan {
    int w = 0, h = 0;
    unsigned field = 0;
    GeomPlan g;
    uint32_t *d_seek = nullptr;
};

struct Staging {
    FieldDesc *h_fields = nullptr;     // pinned
    uint32_t *h_rowinfo = nullptr;
    int32_t *h_hsshift = nullptr;
    HsItem *h_items = nullptr;
    // device copies of the same tables: one set per slot, uploaded on the table stream while the previous
    // batch is still computing (with a single set the upload had to queue behind the previous kernel and
    // left the GPU idle for ~0.1 ms per batch)
    FieldDesc *d_fields = nullptr;
    uint32_t *d_rowinfo = nullptr;
    int32_t *d_hsshift = nullptr;
    HsItem *d_items = nullptr;
    cudaEvent_t consumed = nullptr;    // the H2D copies that read this slot have finished (= tables ready)
    cudaEvent_t kernel_done = nullptr; // the kernels that read this slot's device tables have finished
    bool in_flight = false;
};

template <typename T>
cudaError_t dev_alloc(T **p, size_t n) { return cudaMalloc((void **)p, n * sizeof(T)); }
template <typename T>
cudaError_t pin_alloc(T **p, size_t n) { return cudaMallocHost((void **)p, n * sizeof(T)); }

}  // namespace

struct cvs_ctx {
    cvs_params p;
    int device = 0;
    int max_w = 0, max_h = 0, max_batch = 0, nl_max = 0, hs_max = 0;
    int precision = 0;                         // 0 = float (production), 1 = double (reference arithmetic)
    int noise_fast = 0;                        // CVS_NOISE_FAST requested (fp32 only, small amplitudes only)
    int host_chunk = kHostChunkDefault;
    int bob = 0;                               // fused line doubling (cvs_set_bob)
    int plan_threads = 4;                      // host threads that build the per-row side tables of a batch
    int warm_px = kWarmPx;                     // noise warm-up length (CVS_WARM_PX: tests force the second-chance path)
    int packed_rows = 1;                       // cut the batch's rows into warps across field boundaries (CVS_PACKED_ROWS=0: per field)
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaStream_t s_in = nullptr, s_out = nullptr;      // upload / download streams of the host-pointer path
    std::vector<cudaEvent_t> ev_in, ev_k;              // per chunk: upload done / kernels done
    // CUDA-event pairs around every k_fields launch since the last cvs_kernel_time_reset()
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    size_t ev_used = 0;
    RandCursor cur;
    std::vector<std::unique_ptr<DevPlan>> plans;
    std::vector<std::unique_ptr<ScalePlan>> scale_plans;
    Staging slots[kStagingSlots];
    int next_slot = 0;
    cudaStream_t s_tab = nullptr;              // uploads the per-batch side tables
    int32_t *d_scratch = nullptr;              // head-switch pre-pass rows (written and read on `stream`)
    int32_t *d_status = nullptr;
    int32_t *h_status = nullptr;               // pinned
    float *d_lut_f = nullptr;
    double *d_lut_d = nullptr;
    size_t lut_cap_f = 0, lut_cap_d = 0;       // one capacity per table: they are separate allocations
    bool lut_dirty = true;
    // device pictures for the host-pointer entry points
    // two sets, alternating per call, so an asynchronous call can upload while the previous one drains
    uint8_t *d_src[2] = {nullptr, nullptr}, *d_dst[2] = {nullptr, nullptr};
    size_t d_pic_cap[2] = {0, 0};
    cudaEvent_t host_call_done[2] = {nullptr, nullptr};   // recorded on s_out when a call's last download is queued
    int host_call_parity = 0;
    unsigned long long launches = 0;
    // cvs_field_loop_host: device buffers of the chain and the ring row that field-0 pictures inherit
    uint8_t *fl_src = nullptr, *fl_scaled = nullptr, *fl_out = nullptr, *fl_yuv = nullptr, *fl_last_row = nullptr;
    size_t fl_src_cap = 0, fl_scaled_cap = 0, fl_out_cap = 0, fl_yuv_cap = 0, fl_last_row_cap = 0;
    // cvs_bgra_to_yuv_device: the filter banks of a geometry on the device (sws_filter.h), built on first use
    struct YuvBanks {
        int w = 0, h = 0, v420 = 0, vtaps = 0, htaps = 0;
        bool tile_ok = false;                      // every tile of k_bgra_to_yuv420_tiled fits its shared rows
        int32_t *d = nullptr;                      // vpos | vcoef | hpos | hcoef
        size_t off_vcoef = 0, off_hpos = 0, off_hcoef = 0;
    };
    std::vector<YuvBanks> yuv_banks;
    int sws_rows_kernel = 1;                       // CVS_SWS_ROWS=0: always the one-row kernel (A/B, fallback)
};

namespace {

#define CVS_CUDA(x)                                  \
    do {                                             \
        cudaError_t e_ = (x);                        \
        if (e_ != cudaSuccess) return CVS_ERR_CUDA;  \
    } while (0)

int head_switch_rows_bound(int w) {
    // rows rotated by one head switch: the shift starts at |ishif| <= twidth/2 and decays by 7/8
    // (truncating) per row until it reaches zero, ffmpeg_ntsc.cpp:1704-1707
    int shif = (w + w / 10) / 2 + 1, n = 0;
    while (shif != 0) { shif = (shif * 7) / 8; n++; }
    return n + 1;
}

void free_all(cvs_ctx *c) {
    if (!c) return;
    for (auto &b : c->yuv_banks) if (b.d) cudaFree(b.d);
    c->yuv_banks.clear();
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto &pl : c->plans) if (pl && pl->d_seek) cudaFree(pl->d_seek);
    c->plans.clear();
    for (auto &sp : c->scale_plans) if (sp) { cudaFree(sp->d_first); cudaFree(sp->d_w); cudaFree(sp->d_sws); }
    c->scale_plans.clear();
    for (auto &s : c->slots) {
        if (s.h_fields) cudaFreeHost(s.h_fields);
        if (s.h_rowinfo) cudaFreeHost(s.h_rowinfo);
        if (s.h_hsshift) cudaFreeHost(s.h_hsshift);
        if (s.h_items) cudaFreeHost(s.h_items);
        if (s.consumed) cudaEventDestroy(s.consumed);
        if (s.kernel_done) cudaEventDestroy(s.kernel_done);
        cudaFree(s.d_fields); cudaFree(s.d_rowinfo); cudaFree(s.d_hsshift); cudaFree(s.d_items);
        s = Staging();
    }
    if (c->s_tab) cudaStreamDestroy(c->s_tab);
    cudaFree(c->d_scratch); cudaFree(c->d_status); cudaFree(c->d_lut_f); cudaFree(c->d_lut_d);
    for (int i = 0; i < 2; i++) { cudaFree(c->d_src[i]); cudaFree(c->d_dst[i]); if (c->host_call_done[i]) cudaEventDestroy(c->host_call_done[i]); }
    cudaFree(c->fl_src); cudaFree(c->fl_scaled); cudaFree(c->fl_out); cudaFree(c->fl_yuv); cudaFree(c->fl_last_row);
    if (c->h_status) cudaFreeHost(c->h_status);
    for (auto &pr : c->ev_pool) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto e : c->ev_in) cudaEventDestroy(e);
    for (auto e : c->ev_k) cudaEventDestroy(e);
    c->ev_in.clear(); c->ev_k.clear();
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    c->ev_pool.clear();
    if (c->stream && c->own_stream) cudaStreamDestroy(c->stream);
}

int get_plan(cvs_ctx *c, int w, int h, unsigned field, DevPlan **out) {
    for (auto &pl : c->plans)
        if (pl->w == w && pl->h == h && pl->field == field) { *out = pl.get(); return CVS_OK; }
    std::unique_ptr<DevPlan> pl(new (std::nothrow) DevPlan());
    if (!pl) return CVS_ERR_NOMEM;
    pl->w = w; pl->h = h; pl->field = field;
    build_geom_plan(c->p, w, h, field, pl->g, c->warm_px);
    const size_t words = pl->g.seek.size() > 0 ? pl->g.seek.size() : 1;
    CVS_CUDA(dev_alloc(&pl->d_seek, words));
    if (!pl->g.seek.empty())
        CVS_CUDA(cudaMemcpyAsync(pl->d_seek, pl->g.seek.data(), pl->g.seek.size() * sizeof(uint32_t),
                                 cudaMemcpyHostToDevice, c->stream));
    CVS_CUDA(cudaStreamSynchronize(c->stream));     // g.seek is pageable: finish before anyone can drop it
    *out = pl.get();
    c->plans.push_back(std::move(pl));
    return CVS_OK;
}

template <typename R>
cudaError_t launch_variant(const Variant &v, const LaunchArgs<R> &a, cudaStream_t st);

// the fast-noise instantiations (kern_n_*.cu)
static cudaError_t launch_variant_fast_noise(const Variant &v, const LaunchArgs<float> &a, cudaStream_t st) {
    if (!v.vhs) return v.outfull ? launch_fields<float, false, 9, true, true>(a, st) : launch_fields<float, false, 9, false, true>(a, st);
    if (v.cd == 9) return v.outfull ? launch_fields<float, true, 9, true, true>(a, st) : launch_fields<float, true, 9, false, true>(a, st);
    if (v.cd == 12) return v.outfull ? launch_fields<float, true, 12, true, true>(a, st) : launch_fields<float, true, 12, false, true>(a, st);
    return v.outfull ? launch_fields<float, true, 14, true, true>(a, st) : launch_fields<float, true, 14, false, true>(a, st);
}

template <>
cudaError_t launch_variant<float>(const Variant &v, const LaunchArgs<float> &a, cudaStream_t st) {
    if (a.K.flags & F_NOISE_FAST) return launch_variant_fast_noise(v, a, st);
    if (!v.vhs) return v.outfull ? launch_fields<float, false, 9, true>(a, st) : launch_fields<float, false, 9, false>(a, st);
    if (v.cd == 9) return v.outfull ? launch_fields<float, true, 9, true>(a, st) : launch_fields<float, true, 9, false>(a, st);
    if (v.cd == 12) return v.outfull ? launch_fields<float, true, 12, true>(a, st) : launch_fields<float, true, 12, false>(a, st);
    return v.outfull ? launch_fields<float, true, 14, true>(a, st) : launch_fields<float, true, 14, false>(a, st);
}
template <>
cudaError_t launch_variant<double>(const Variant &v, const LaunchArgs<double> &a, cudaStream_t st) {
    if (!v.vhs) return v.outfull ? launch_fields<double, false, 9, true>(a, st) : launch_fields<double, false, 9, false>(a, st);
    if (v.cd == 9) return v.outfull ? launch_fields<double, true, 9, true>(a, st) : launch_fields<double, true, 9, false>(a, st);
    if (v.cd == 12) return v.outfull ? launch_fields<double, true, 12, true>(a, st) : launch_fields<double, true, 12, false>(a, st);
    return v.outfull ? launch_fields<double, true, 14, true>(a, st) : launch_fields<double, true, 14, false>(a, st);
}

template <typename R>
cudaError_t occupancy_variant(const Variant &v, int *n) {
    if (!v.vhs) return v.outfull ? occupancy_fields<R, false, 9, true>(n) : occupancy_fields<R, false, 9, false>(n);
    if (v.cd == 9) return v.outfull ? occupancy_fields<R, true, 9, true>(n) : occupancy_fields<R, true, 9, false>(n);
    if (v.cd == 12) return v.outfull ? occupancy_fields<R, true, 12, true>(n) : occupancy_fields<R, true, 12, false>(n);
    return v.outfull ? occupancy_fields<R, true, 14, true>(n) : occupancy_fields<R, true, 14, false>(n);
}

template <typename R> R *&lut_ptr(cvs_ctx *c);
template <> float *&lut_ptr<float>(cvs_ctx *c) { return c->d_lut_f; }
template <> double *&lut_ptr<double>(cvs_ctx *c) { return c->d_lut_d; }
template <typename R> size_t &lut_cap(cvs_ctx *c);
template <> size_t &lut_cap<float>(cvs_ctx *c) { return c->lut_cap_f; }
template <> size_t &lut_cap<double>(cvs_ctx *c) { return c->lut_cap_d; }

template <typename R>
int launch_batch(cvs_ctx *c, const Staging &sl, const Variant &v, int w, int h, int nfields, int max_nl, int total_rows,
                 bool packed, int nitems, int src_stride, int dst_stride, int opposite, bool vec_src, bool vec_dst) {
    LaunchArgs<R> a;
    std::vector<R> lut;
    make_kconst<R>(c->p, w, h, v.outfull, a.K, lut);
    if (lut.size() > lut_cap<R>(c) || !lut_ptr<R>(c)) {
        if (lut_ptr<R>(c)) { CVS_CUDA(cudaStreamSynchronize(c->stream)); cudaFree(lut_ptr<R>(c)); lut_ptr<R>(c) = nullptr; lut_cap<R>(c) = 0; }
        CVS_CUDA(dev_alloc(&lut_ptr<R>(c), lut.size() + 2));
        lut_cap<R>(c) = lut.size() + 2;
        c->lut_dirty = true;
    }
    if (c->lut_dirty) {
        CVS_CUDA(cudaMemcpyAsync(lut_ptr<R>(c), lut.data(), lut.size() * sizeof(R), cudaMemcpyHostToDevice, c->stream));
        CVS_CUDA(cudaStreamSynchronize(c->stream));   // `lut` is a pageable temporary
        c->lut_dirty = false;
    }
    a.K.phase_lut = lut_ptr<R>(c);
    // fast per-pixel noise: production arithmetic only, and only while the noise is sub-LSB (include/cvs_ntsc.h)
    if (c->noise_fast && sizeof(R) == 4 && a.K.vnoise + 2 * a.K.cnoise <= 96) a.K.flags |= F_NOISE_FAST;
    a.fields = sl.d_fields;
    a.nfields = nfields;
    a.warps_per_field = (max_nl + kRowsPerWarp - 1) / kRowsPerWarp;
    a.total_warps = a.warps_per_field * nfields;
    a.max_nl = max_nl;
    a.total_rows = total_rows;
    a.packed = packed ? 1 : 0;
    if (packed) a.total_warps = (total_rows + kRowsPerWarp - 1) / kRowsPerWarp;
    a.src_stride = src_stride;
    a.dst_stride = dst_stride;
    a.opposite = opposite;
    a.vec_src = vec_src;
    a.vec_dst = vec_dst;
    a.bob = c->bob;
    a.warm_px = c->warm_px;
    a.status = c->d_status;
    if (nitems > 0) {
        CVS_CUDA(launch_headswitch<R>(a, sl.d_items, nitems, c->stream));
        c->launches++;
    }
    if (c->ev_used == c->ev_pool.size() && c->ev_pool.size() < kMaxEventPairs) {
        cudaEvent_t e0, e1;
        CVS_CUDA(cudaEventCreate(&e0));
        CVS_CUDA(cudaEventCreate(&e1));
        c->ev_pool.push_back(std::make_pair(e0, e1));
    }
    const bool timed = c->ev_used < c->ev_pool.size();
    if (timed) CVS_CUDA(cudaEventRecord(c->ev_pool[c->ev_used].first, c->stream));
    CVS_CUDA(launch_variant<R>(v, a, c->stream));
    c->launches++;
    if (timed) { CVS_CUDA(cudaEventRecord(c->ev_pool[c->ev_used].second, c->stream)); c->ev_used++; }
    return CVS_OK;
}

// Plan + launch n fields whose pictures are device-resident.  explicit_field < 0 => the
// reference loop's schedule field = ((fieldno) & 1) ^ 1  (ffmpeg_ntsc.cpp:2229).
// src_index (optional): source picture of field k is src + src_index[k] * src_pic_stride instead of src + k * ...
int run_device(cvs_ctx *c, uint8_t *dst, size_t dst_pic_stride, int dst_stride, const uint8_t *src,
               size_t src_pic_stride, int src_stride, int w, int h, int interlaced, int tff, int n,
               unsigned long long first_fieldno, int explicit_field, const int32_t *src_index = nullptr) {
    if (!c || !dst || !src) return CVS_ERR_INVALID_ARG;
    if (w <= 0 || h <= 0 || n < 0) return CVS_ERR_INVALID_ARG;
    if (dst_stride < 4 * w || src_stride < 4 * w) return CVS_ERR_INVALID_ARG;          // :1580-1581
    if ((dst_stride & 3) || (src_stride & 3) || ((uintptr_t)dst & 3) || ((uintptr_t)src & 3)) return CVS_ERR_INVALID_ARG;
    if (w > c->max_w || h > c->max_h || n > c->max_batch) return CVS_ERR_CAPACITY;
    if (n == 0) return CVS_OK;
    if (cudaSetDevice(c->device) != cudaSuccess) return CVS_ERR_CUDA;

    const Variant v = pick_variant(c->p);
    Staging &sl = c->slots[c->next_slot];
    c->next_slot = (c->next_slot + 1) % kStagingSlots;
    if (sl.in_flight) { CVS_CUDA(cudaEventSynchronize(sl.consumed)); sl.in_flight = false; }

    const int opposite = interlaced ? (tff ? 1 : 0) : 0;                               // :1585-1588
    // Planning.  Pass 1 (serial, cheap): per field, its plan and the rand() position of its first draw
    // relative to the batch (a prefix sum: the draw count is a function of the geometry).  Pass 2 (a few
    // host threads): every thread jumps a copy of the cursor to its first field (one x^d, d < 2^40) and then
    // walks its fields (one 31x31 jump each) building the per-row side tables, which is where the time goes
    // (5.8 us per 1080p field on one thread).  Pass 3 (serial): the head-switch pre-pass work list.
    int nitems = 0, max_nl = 0, min_nl = 1 << 30, total_rows = 0;
    struct Job { DevPlan *pl; unsigned long long rel; int hs_count; };   // rel: draws of the batch before this field
    unsigned long long total_draws = 0;
    std::vector<Job> jobs((size_t)n);
    for (int k = 0; k < n; k++) {
        const unsigned long long fieldno = first_fieldno + (unsigned long long)k;
        const unsigned field = explicit_field >= 0 ? (unsigned)explicit_field : (unsigned)((fieldno & 1) ^ 1);
        FieldDesc &fd = sl.h_fields[k];
        std::memset(&fd, 0, sizeof(fd));
        fd.src = src + (size_t)(src_index ? src_index[k] : k) * src_pic_stride;
        fd.dst = dst + (size_t)k * dst_pic_stride;
        fd.fieldno = fieldno;
        fd.field = (int32_t)field;
        jobs[(size_t)k].pl = nullptr;
        jobs[(size_t)k].rel = total_draws;
        jobs[(size_t)k].hs_count = 0;
        fd.row_start = total_rows;
        if ((int)field >= h) { fd.nl = 0; continue; }      // no rows of this parity: nothing drawn, nothing written
        DevPlan *pl = nullptr;
        int rc = get_plan(c, w, h, field, &pl);
        if (rc != CVS_OK) return rc;
        jobs[(size_t)k].pl = pl;
        total_draws += pl->g.ndraws;
        fd.nl = pl->g.nl;
        fd.row_start = total_rows;
        total_rows += fd.nl;
        if (fd.nl > max_nl) max_nl = fd.nl;
        if (fd.nl < min_nl) min_nl = fd.nl;
        fd.seek = pl->d_seek;
        fd.rowinfo = sl.d_rowinfo + (size_t)k * c->nl_max;
        fd.hs_scratch = c->d_scratch + (size_t)k * c->hs_max * (size_t)c->max_w;
        fd.hs_shift = sl.d_hsshift + (size_t)k * c->hs_max;
    }
    std::atomic<bool> capacity_error(false);
    const RandCursor batch_start = c->cur;
    RandCursor batch_end = c->cur;
    auto plan_range = [&](int k0, int k1) {
        FieldSide fs;
        RandCursor cur = batch_start;
        if (jobs[(size_t)k0].rel) cur.advance(jobs[(size_t)k0].rel);
        for (int k = k0; k < k1; k++) {
            Job &jb = jobs[(size_t)k];
            if (!jb.pl) continue;
            FieldDesc &fd = sl.h_fields[k];
            build_field_side_at(c->p, jb.pl->g, cur, fs);
            cur.jump(jb.pl->g.jumpN, jb.pl->g.ndraws);
            std::memcpy(sl.h_rowinfo + (size_t)k * c->nl_max, fs.rowinfo.data(), fs.rowinfo.size() * sizeof(uint32_t));
            std::memcpy(fd.window, fs.window, sizeof(fs.window));
            fd.hs_first = fs.hs_first;
            fd.hs_count = fs.hs_count;
            if (fs.hs_count > c->hs_max) { capacity_error = true; fd.hs_count = 0; continue; }
            for (int i = 0; i < fs.hs_count; i++) sl.h_hsshift[(size_t)k * c->hs_max + i] = fs.hs_shift[(size_t)i];
            jb.hs_count = fs.hs_count;
        }
        if (k1 == n) batch_end = cur;                  // (exactly one range ends the batch)
    };
    const int nthreads = (n >= 64) ? c->plan_threads : 1;
    if (nthreads <= 1) {
        plan_range(0, n);
    } else {
        // a thread that cannot be started (std::system_error must not cross the C ABI) leaves its range, and
        // every later one, to this thread
        std::vector<std::thread> pool;
        const int per = (n + nthreads - 1) / nthreads;
        int done_to = per < n ? per : n;               // ranges [done_to, n) still need a planner
        try {
            pool.reserve((size_t)nthreads);
            for (int t = 1; t < nthreads && t * per < n; t++) {
                const int k1 = (t + 1) * per < n ? (t + 1) * per : n;
                pool.emplace_back(plan_range, t * per, k1);
                done_to = k1;
            }
        } catch (...) {
        }
        plan_range(0, per < n ? per : n);
        for (auto &th : pool) th.join();
        const int started_to = done_to;
        for (int k0 = started_to; k0 < n; k0 += per) plan_range(k0, k0 + per < n ? k0 + per : n);
    }
    // From here on the call either completes or leaves the rand() position where it was (the header's
    // contract: an invalid or failed call does not consume draws).
    struct CursorGuard {
        cvs_ctx *c; const RandCursor &start; bool armed;
        ~CursorGuard() { if (armed) c->cur = start; }
    } guard{c, batch_start, true};
    c->cur = batch_end;                                // the stream position after the batch
    if (capacity_error.load()) return CVS_ERR_CAPACITY;
    for (int k = 0; k < n; k++)
        for (int i = 0; i < jobs[(size_t)k].hs_count; i++) {
            sl.h_items[nitems].field_idx = k;
            sl.h_items[nitems].slot = i;
            nitems++;
        }
    if (max_nl == 0) { guard.armed = false; return CVS_OK; }

    // tables go up on their own stream, behind the kernels that last read this slot's device copies,
    // and the compute stream picks them up through an event: the upload overlaps the previous batch
    CVS_CUDA(cudaStreamWaitEvent(c->s_tab, sl.kernel_done, 0));
    CVS_CUDA(cudaMemcpyAsync(sl.d_fields, sl.h_fields, (size_t)n * sizeof(FieldDesc), cudaMemcpyHostToDevice, c->s_tab));
    CVS_CUDA(cudaMemcpyAsync(sl.d_rowinfo, sl.h_rowinfo, (size_t)n * c->nl_max * sizeof(uint32_t), cudaMemcpyHostToDevice, c->s_tab));
    if (nitems > 0) {
        CVS_CUDA(cudaMemcpyAsync(sl.d_hsshift, sl.h_hsshift, (size_t)n * c->hs_max * sizeof(int32_t), cudaMemcpyHostToDevice, c->s_tab));
        CVS_CUDA(cudaMemcpyAsync(sl.d_items, sl.h_items, (size_t)nitems * sizeof(HsItem), cudaMemcpyHostToDevice, c->s_tab));
    }
    CVS_CUDA(cudaEventRecord(sl.consumed, c->s_tab));
    sl.in_flight = true;
    CVS_CUDA(cudaStreamWaitEvent(c->stream, sl.consumed, 0));

    // packed mapping (scanline_kernels.cuh): a warp of 31 consecutive rows then meets at most two fields
    const bool packed = c->packed_rows && n > 1 && min_nl > kRowsPerWarp;
    const bool vec_src = ((uintptr_t)src % 16 == 0) && (src_stride % 16 == 0) && (src_pic_stride % 16 == 0);
    const bool vec_dst = ((uintptr_t)dst % 16 == 0) && (dst_stride % 16 == 0) && (dst_pic_stride % 16 == 0);
    const int rc = c->precision
        ? launch_batch<double>(c, sl, v, w, h, n, max_nl, total_rows, packed, nitems, src_stride, dst_stride, opposite, vec_src, vec_dst)
        : launch_batch<float>(c, sl, v, w, h, n, max_nl, total_rows, packed, nitems, src_stride, dst_stride, opposite, vec_src, vec_dst);
    if (rc != CVS_OK) return rc;
    CVS_CUDA(cudaEventRecord(sl.kernel_done, c->stream));
    guard.armed = false;
    return CVS_OK;
}

int check_status(cvs_ctx *c) {
    CVS_CUDA(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CVS_CUDA(cudaStreamSynchronize(c->stream));
    if (*c->h_status != 0) {
        CVS_CUDA(cudaMemsetAsync(c->d_status, 0, sizeof(int32_t), c->stream));
        return CVS_ERR_NOISE_SYNC;
    }
    return CVS_OK;
}

// Host-pointer form.  The batch is cut into chunks of kHostChunk fields and pipelined over three
// streams: chunk i+1 is uploaded (copy engine 1) while chunk i is computed and chunk i-1 is
// downloaded (copy engine 2), so PCIe runs full duplex and the kernels hide behind it.  Only the rows
// of the requested field cross the bus: source rows min(field + 2r + opposite, h-1) (:1599) in, dst
// rows field + 2r out; every other dst row stays untouched on the host, as in the reference (:1910).
int run_host(cvs_ctx *c, uint8_t *dst, size_t dst_pic_stride, int dst_stride, const uint8_t *src,
             size_t src_pic_stride, int src_stride, int w, int h, int interlaced, int tff, int n,
             unsigned long long first_fieldno, int explicit_field, bool async) {
    if (!c || !dst || !src) return CVS_ERR_INVALID_ARG;
    if (w <= 0 || h <= 0 || n < 0) return CVS_ERR_INVALID_ARG;
    if (dst_stride < 4 * w || src_stride < 4 * w) return CVS_ERR_INVALID_ARG;
    if (w > c->max_w || h > c->max_h || n > c->max_batch) return CVS_ERR_CAPACITY;
    if (n == 0) return CVS_OK;
    if (cudaSetDevice(c->device) != cudaSuccess) return CVS_ERR_CUDA;
    // device pictures: rows padded to 16 bytes, full height (only the field rows are transferred)
    const int dstride = ((4 * w + 15) / 16) * 16;
    const size_t dpic = (size_t)dstride * (size_t)h;
    const int set = c->host_call_parity;
    c->host_call_parity ^= 1;
    if (!c->host_call_done[set]) CVS_CUDA(cudaEventCreateWithFlags(&c->host_call_done[set], cudaEventDisableTiming));
    if (dpic * (size_t)n > c->d_pic_cap[set]) {
        CVS_CUDA(cudaStreamSynchronize(c->s_in));
        CVS_CUDA(cudaStreamSynchronize(c->stream));
        CVS_CUDA(cudaStreamSynchronize(c->s_out));
        cudaFree(c->d_src[set]); cudaFree(c->d_dst[set]);
        c->d_src[set] = c->d_dst[set] = nullptr;
        c->d_pic_cap[set] = 0;
        CVS_CUDA(cudaMalloc((void **)&c->d_src[set], dpic * (size_t)n));
        CVS_CUDA(cudaMalloc((void **)&c->d_dst[set], dpic * (size_t)n));
        c->d_pic_cap[set] = dpic * (size_t)n;
    }
    uint8_t *const d_src = c->d_src[set], *const d_dst = c->d_dst[set];
    // this buffer set was last used two calls ago: its downloads must have drained before we overwrite it
    CVS_CUDA(cudaStreamWaitEvent(c->s_in, c->host_call_done[set], 0));
    // chunk boundaries: a synchronous call ramps up (8, 8, 16, then full chunks) so the first kernel
    // starts early and the pipeline fills fast; an asynchronous call is already overlapped with its
    // predecessor and uses full chunks throughout
    std::vector<int> bounds;
    bounds.push_back(0);
    {
        int ramp[3] = {8, 8, 16}, ri = 0;
        while (bounds.back() < n) {
            int sz = c->host_chunk;
            if (!async && ri < 3 && ramp[ri] < sz) sz = ramp[ri++];
            bounds.push_back(bounds.back() + sz < n ? bounds.back() + sz : n);
        }
    }
    const int nchunks = (int)bounds.size() - 1;
    while ((int)c->ev_in.size() < nchunks) {
        cudaEvent_t a, b;
        CVS_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CVS_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        c->ev_in.push_back(a);
        c->ev_k.push_back(b);
    }
    const int opposite = interlaced ? (tff ? 1 : 0) : 0;
    const RandCursor call_start = c->cur;
    auto field_of = [&](int k) {
        const unsigned long long fieldno = first_fieldno + (unsigned long long)k;
        return explicit_field >= 0 ? explicit_field : (int)((fieldno & 1) ^ 1);
    };
    for (int ci = 0; ci < nchunks; ci++) {
        const int k0 = bounds[(size_t)ci], k1 = bounds[(size_t)ci + 1];
        for (int k = k0; k < k1; k++) {                       // upload, stream s_in
            const int field = field_of(k);
            if (field >= h) continue;
            const int nl = (h - field + 1) / 2;
            const int y0 = field + opposite;
            int nreg = nl;                                    // stride-2 run, plus the clamped last row
            if (y0 + 2 * (nl - 1) > h - 1) nreg = nl - 1;
            const uint8_t *hs = src + (size_t)k * src_pic_stride;
            uint8_t *ds = d_src + (size_t)k * dpic;
            if (nreg > 0)
                CVS_CUDA(cudaMemcpy2DAsync(ds + (size_t)y0 * dstride, (size_t)2 * dstride, hs + (size_t)y0 * src_stride,
                                           (size_t)2 * src_stride, (size_t)4 * w, (size_t)nreg, cudaMemcpyHostToDevice, c->s_in));
            if (nreg < nl)
                CVS_CUDA(cudaMemcpyAsync(ds + (size_t)(h - 1) * dstride, hs + (size_t)(h - 1) * src_stride, (size_t)4 * w,
                                         cudaMemcpyHostToDevice, c->s_in));
        }
        CVS_CUDA(cudaEventRecord(c->ev_in[ci], c->s_in));
        CVS_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in[ci], 0));
        int rc = run_device(c, d_dst + (size_t)k0 * dpic, dpic, dstride, d_src + (size_t)k0 * dpic, dpic, dstride, w, h,
                            interlaced, tff, k1 - k0, first_fieldno + (unsigned long long)k0, explicit_field);
        if (rc != CVS_OK) {        // a failed call consumes no draws, whichever chunk failed
            cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->s_out);
            c->cur = call_start;
            return rc;
        }
        CVS_CUDA(cudaEventRecord(c->ev_k[ci], c->stream));
        CVS_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_k[ci], 0));
        for (int k = k0; k < k1; k++) {                       // download, stream s_out
            const int field = field_of(k);
            if (field >= h) continue;
            const int nl = (h - field + 1) / 2;
            if (!c->bob) {
                CVS_CUDA(cudaMemcpy2DAsync(dst + (size_t)k * dst_pic_stride + (size_t)field * dst_stride, (size_t)2 * dst_stride,
                                           d_dst + (size_t)k * dpic + (size_t)field * dstride, (size_t)2 * dstride,
                                           (size_t)4 * w, (size_t)nl, cudaMemcpyDeviceToHost, c->s_out));
            } else {
                // line-doubled picture: rows 0 .. field + 2(nl-1) are all written (for field 0 and even h the
                // last row is not, and keeps what the host buffer held: ffmpeg_ntsc.cpp:2247)
                const int rows = field + 2 * (nl - 1) + 1;
                CVS_CUDA(cudaMemcpy2DAsync(dst + (size_t)k * dst_pic_stride, (size_t)dst_stride,
                                           d_dst + (size_t)k * dpic, (size_t)dstride,
                                           (size_t)4 * w, (size_t)rows, cudaMemcpyDeviceToHost, c->s_out));
            }
        }
    }
    CVS_CUDA(cudaEventRecord(c->host_call_done[set], c->s_out));
    if (async) return CVS_OK;
    CVS_CUDA(cudaStreamSynchronize(c->s_out));
    return check_status(c);
}

}  // namespace

extern "C" {

int cvs_create(cvs_ctx **out, const cvs_params *p, int device, int max_w, int max_h, int max_batch) {
    if (!out || !p || max_w <= 0 || max_h <= 0 || max_batch <= 0) return CVS_ERR_INVALID_ARG;
    if (p->struct_size != (int32_t)sizeof(cvs_params)) return CVS_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return CVS_ERR_CUDA;   // no CPU fallback
    if (device < 0 || device >= ndev) return CVS_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return CVS_ERR_CUDA;
    cvs_ctx *c = new (std::nothrow) cvs_ctx();
    if (!c) return CVS_ERR_NOMEM;
    c->p = *p;
    c->device = device;
    c->max_w = max_w; c->max_h = max_h; c->max_batch = max_batch;
    c->nl_max = (max_h + 1) / 2;
    if (const char *e = std::getenv("CVS_HOST_CHUNK")) {      // tuning knob for experiments
        const int v = std::atoi(e);
        if (v >= 1 && v <= 1024) c->host_chunk = v;
    }
    {
        unsigned hw = std::thread::hardware_concurrency();
        if (hw > 0 && (int)hw < c->plan_threads) c->plan_threads = (int)hw;
        if (const char *e = std::getenv("CVS_PLAN_THREADS")) {  // tuning knob
            const int v = std::atoi(e);
            if (v >= 1 && v <= 64) c->plan_threads = v;
        }
    }
    if (const char *e = std::getenv("CVS_PACKED_ROWS")) c->packed_rows = std::atoi(e) != 0;   // experiments / tests
    if (const char *e = std::getenv("CVS_SWS_ROWS")) c->sws_rows_kernel = std::atoi(e) != 0;
    if (const char *e = std::getenv("CVS_WARM_PX")) {           // tests: a short warm-up makes the first attempt fail
        const int v = std::atoi(e);
        if (v >= 0 && v <= kWarmPx) c->warm_px = v;
    }
    c->hs_max = head_switch_rows_bound(max_w);
    if (c->hs_max > c->nl_max) c->hs_max = c->nl_max;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
    for (auto &s : c->slots) {
        if (e == cudaSuccess) e = pin_alloc(&s.h_fields, (size_t)max_batch);
        if (e == cudaSuccess) e = pin_alloc(&s.h_rowinfo, (size_t)max_batch * c->nl_max);
        if (e == cudaSuccess) e = pin_alloc(&s.h_hsshift, (size_t)max_batch * c->hs_max);
        if (e == cudaSuccess) e = pin_alloc(&s.h_items, (size_t)max_batch * c->hs_max);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.kernel_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = dev_alloc(&s.d_fields, (size_t)max_batch);
        if (e == cudaSuccess) e = dev_alloc(&s.d_rowinfo, (size_t)max_batch * c->nl_max);
        if (e == cudaSuccess) e = dev_alloc(&s.d_hsshift, (size_t)max_batch * c->hs_max);
        if (e == cudaSuccess) e = dev_alloc(&s.d_items, (size_t)max_batch * c->hs_max);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_tab, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = dev_alloc(&c->d_scratch, (size_t)max_batch * c->hs_max * (size_t)max_w);
    if (e == cudaSuccess) e = dev_alloc(&c->d_status, 1);
    if (e == cudaSuccess) e = pin_alloc(&c->h_status, 1);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_status, 0, sizeof(int32_t), c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        free_all(c);
        delete c;
        return e == cudaErrorMemoryAllocation ? CVS_ERR_NOMEM : CVS_ERR_CUDA;
    }
    *out = c;
    return CVS_OK;
}

void cvs_destroy(cvs_ctx *ctx) {
    if (!ctx) return;
    free_all(ctx);
    delete ctx;
}

int cvs_set_params(cvs_ctx *ctx, const cvs_params *p) {
    if (!ctx || !p || p->struct_size != (int32_t)sizeof(cvs_params)) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    CVS_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->p = *p;
    for (auto &pl : ctx->plans) if (pl && pl->d_seek) cudaFree(pl->d_seek);
    ctx->plans.clear();                      // draw layout depends on the enabled stages
    ctx->lut_dirty = true;
    return CVS_OK;
}

int cvs_preferred_batch(cvs_ctx *ctx, int w, int h, int max_batch) {
    // A field is ceil(nl/31) warp-tasks of equal length (one scanline per lane), so a launch runs in
    // waves of (SMs x resident CTAs x warps per CTA) tasks; a batch that fills whole waves wastes none.
    if (!ctx || w <= 0 || h <= 0 || max_batch <= 0) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    int sms = 0, ctas = 0;
    CVS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    const Variant v = pick_variant(ctx->p);
    if (ctx->precision) CVS_CUDA(occupancy_variant<double>(v, &ctas));
    else CVS_CUDA(occupancy_variant<float>(v, &ctas));
    if (max_batch > ctx->max_batch) max_batch = ctx->max_batch;
    const long long slots = (long long)sms * ctas * kWarpsPerCta;
    if (slots <= 0) return max_batch;
    if (ctx->packed_rows && h / 2 > kRowsPerWarp) {
        // packed mapping: b consecutive fields are ceil(rows(b) / 31) tasks, rows(b) = nl(0) + nl(1) + ... with
        // the parities alternating (odd heights: (h+1)/2 and h/2 rows)
        auto tasks = [&](long long b) { return (((b + 1) / 2) * ((h + 1) / 2) + (b / 2) * (h / 2) + kRowsPerWarp - 1) / kRowsPerWarp; };
        const long long waves = tasks(max_batch) / slots;
        if (waves < 1) return max_batch;
        long long b = max_batch;
        while (b > 1 && tasks(b) > waves * slots) b--;
        return (int)b;
    }
    const long long tasks_per_field = ((h + 1) / 2 + kRowsPerWarp - 1) / kRowsPerWarp;
    // largest batch <= max_batch whose task count is at most a whole number of waves
    const long long waves = ((long long)max_batch * tasks_per_field) / slots;
    if (waves < 1) return max_batch;
    const long long b = (waves * slots) / tasks_per_field;
    return (int)(b < 1 ? 1 : b);
}

int cvs_set_bob(cvs_ctx *ctx, int enable) {
    if (!ctx) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    CVS_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->bob = enable ? 1 : 0;
    return CVS_OK;
}

int cvs_set_noise_mode(cvs_ctx *ctx, int mode) {
    if (!ctx || (mode != CVS_NOISE_EXACT && mode != CVS_NOISE_FAST)) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    CVS_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->noise_fast = mode == CVS_NOISE_FAST;
    return CVS_OK;
}

int cvs_set_precision(cvs_ctx *ctx, int use_double) {
    if (!ctx) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    CVS_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->precision = use_double ? 1 : 0;
    ctx->lut_dirty = true;
    return CVS_OK;
}

int cvs_composite_layer(cvs_ctx *ctx, uint8_t *dst, int dst_stride, const uint8_t *src, int src_stride,
                        int w, int h, int src_interlaced, int src_top_field_first,
                        unsigned field, unsigned long long fieldno) {
    if (field > 1) {
        // the reference accepts any `field` (rows y = field, field+2, ...); only 0/1 occur (:2229)
        return CVS_ERR_INVALID_ARG;
    }
    return run_host(ctx, dst, 0, dst_stride, src, 0, src_stride, w, h, src_interlaced, src_top_field_first, 1,
                    fieldno, (int)field, false);
}

int cvs_composite_fields_device(cvs_ctx *ctx, void *dst, size_t dst_pic_stride, int dst_stride, const void *src,
                                size_t src_pic_stride, int src_stride, int w, int h, int src_interlaced,
                                int src_top_field_first, int n, unsigned long long first_fieldno) {
    return run_device(ctx, (uint8_t *)dst, dst_pic_stride, dst_stride, (const uint8_t *)src, src_pic_stride, src_stride,
                      w, h, src_interlaced, src_top_field_first, n, first_fieldno, -1);
}

int cvs_composite_fields_host(cvs_ctx *ctx, void *dst, size_t dst_pic_stride, int dst_stride, const void *src,
                              size_t src_pic_stride, int src_stride, int w, int h, int src_interlaced,
                              int src_top_field_first, int n, unsigned long long first_fieldno) {
    return run_host(ctx, (uint8_t *)dst, dst_pic_stride, dst_stride, (const uint8_t *)src, src_pic_stride, src_stride,
                    w, h, src_interlaced, src_top_field_first, n, first_fieldno, -1, false);
}

int cvs_composite_fields_host_async(cvs_ctx *ctx, void *dst, size_t dst_pic_stride, int dst_stride, const void *src,
                                    size_t src_pic_stride, int src_stride, int w, int h, int src_interlaced,
                                    int src_top_field_first, int n, unsigned long long first_fieldno) {
    return run_host(ctx, (uint8_t *)dst, dst_pic_stride, dst_stride, (const uint8_t *)src, src_pic_stride, src_stride,
                    w, h, src_interlaced, src_top_field_first, n, first_fieldno, -1, true);
}

int cvs_synchronize(cvs_ctx *ctx) {
    if (!ctx) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    CVS_CUDA(cudaStreamSynchronize(ctx->s_in));
    const int rc = check_status(ctx);          // synchronises the compute stream
    CVS_CUDA(cudaStreamSynchronize(ctx->s_out));
    return rc;
}

int cvs_alloc_host(cvs_ctx *ctx, void **out, size_t bytes) {
    if (!ctx || !out || bytes == 0) return CVS_ERR_INVALID_ARG;
    *out = nullptr;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    const cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { *out = nullptr; return e == cudaErrorMemoryAllocation ? CVS_ERR_NOMEM : CVS_ERR_CUDA; }
    return CVS_OK;
}

int cvs_free_host(cvs_ctx *ctx, void *p) {
    if (!ctx) return CVS_ERR_INVALID_ARG;
    if (!p) return CVS_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    CVS_CUDA(cudaFreeHost(p));
    return CVS_OK;
}

int cvs_rng_seek(cvs_ctx *ctx, unsigned long long draws_consumed) {
    if (!ctx) return CVS_ERR_INVALID_ARG;
    ctx->cur.seek(draws_consumed);
    return CVS_OK;
}

int cvs_rng_tell(const cvs_ctx *ctx, unsigned long long *draws_consumed) {
    if (!ctx || !draws_consumed) return CVS_ERR_INVALID_ARG;
    *draws_consumed = ctx->cur.pos();
    return CVS_OK;
}

unsigned long long cvs_kernel_launches(const cvs_ctx *ctx) { return ctx ? ctx->launches : 0; }

int cvs_set_stream(cvs_ctx *ctx, void *cuda_stream) {
    if (!ctx) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    CVS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    for (auto &s : ctx->slots) s.in_flight = false;
    ctx->ev_used = 0;
    return CVS_OK;
}

int cvs_bgra_to_yuv_device(cvs_ctx *ctx, void *y, int ly, long long y_pic_stride, void *u, int lu, long long u_pic_stride,
                           void *v, int lv, long long v_pic_stride, const void *bgra, int stride,
                           long long bgra_pic_stride, int w, int h, int n, int format) {
    if (!ctx || !y || !u || !v || !bgra || w <= 0 || h <= 0 || n < 0) return CVS_ERR_INVALID_ARG;
    if (format != CVS_YUV420P && format != CVS_YUV422P) return CVS_ERR_INVALID_ARG;
    const int cw = (w + 1) / 2;
    if (stride < 4 * w || (stride & 3) || ((uintptr_t)bgra & 3) || ly < w || lu < cw || lv < cw) return CVS_ERR_INVALID_ARG;
    if (n == 0) return CVS_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    const bool v420 = format == CVS_YUV420P;
    const cvs_ctx::YuvBanks *bk = nullptr;
    for (const auto &b : ctx->yuv_banks) if (b.w == w && b.h == h && b.v420 == (int)v420) { bk = &b; break; }
    if (!bk) {
        // the library's filter banks for this geometry: chroma rows (4:2:0: 2:1 bilinear; 4:2:2: the unit filter) and,
        // for odd widths only, chroma columns
        const int crows_ = v420 ? (h + 1) / 2 : h;
        const FilterBank vb = bilinear_bank(h, crows_, 1 << 12);
        FilterBank hb;
        if (w & 1) hb = bilinear_bank(w, cw, 1 << 14);
        auto inside = [](const FilterBank &f, int srcn) {
            for (int32_t q : f.pos) if (q < 0 || q + f.taps > srcn) return false;
            return f.taps >= 1 && f.taps <= kYuvMaxTaps;
        };
        if (!inside(vb, h) || ((w & 1) && !inside(hb, w))) return CVS_ERR_UNSUPPORTED;
        cvs_ctx::YuvBanks nb;
        nb.w = w; nb.h = h; nb.v420 = v420; nb.vtaps = vb.taps; nb.htaps = hb.taps;
        nb.tile_ok = v420;
        for (int c0 = 0; v420 && c0 < crows_; c0 += kTileC) {
            const int c1 = std::min(c0 + kTileC, crows_) - 1;
            const int lo = std::min(vb.pos[(size_t)c0], 2 * c0), hi = std::max(vb.pos[(size_t)c1] + vb.taps - 1, std::min(2 * c1 + 1, h - 1));
            if (hi - lo + 1 > kTileRows) nb.tile_ok = false;
            for (int c = c0; c <= c1; c++) if (vb.pos[(size_t)c] < lo) nb.tile_ok = false;
        }
        std::vector<int32_t> blob(vb.pos);
        nb.off_vcoef = blob.size(); blob.insert(blob.end(), vb.coef.begin(), vb.coef.end());
        nb.off_hpos = blob.size(); blob.insert(blob.end(), hb.pos.begin(), hb.pos.end());
        nb.off_hcoef = blob.size(); blob.insert(blob.end(), hb.coef.begin(), hb.coef.end());
        CVS_CUDA(cudaMalloc((void **)&nb.d, blob.size() * sizeof(int32_t)));
        // pageable source: the copy has left the host buffer when the call returns
        if (cudaMemcpyAsync(nb.d, blob.data(), blob.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
            cudaFree(nb.d);
            return CVS_ERR_CUDA;
        }
        if (ctx->yuv_banks.size() >= 8) {                  // a handful of geometries per context at most
            CVS_CUDA(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->yuv_banks.front().d);
            ctx->yuv_banks.erase(ctx->yuv_banks.begin());
        }
        ctx->yuv_banks.push_back(nb);
        bk = &ctx->yuv_banks.back();
    }
    YuvArgs a;
    a.bgra = (const uint8_t *)bgra; a.y = (uint8_t *)y; a.u = (uint8_t *)u; a.v = (uint8_t *)v;
    a.sp_bgra = bgra_pic_stride; a.sp_y = y_pic_stride; a.sp_u = u_pic_stride; a.sp_v = v_pic_stride;
    a.stride = stride; a.ly = ly; a.lu = lu; a.lv = lv;
    a.w = w; a.h = h; a.n = n; a.v420 = v420;
    a.vpos = bk->d; a.vcoef = bk->d + bk->off_vcoef; a.hpos = bk->d + bk->off_hpos; a.hcoef = bk->d + bk->off_hcoef;
    a.vtaps = bk->vtaps; a.htaps = bk->htaps;
    a.c = yuv_coef_bt601();
    const int crows = v420 ? (h + 1) / 2 : h;
    if (crows > 65535 || n > 65535) return CVS_ERR_CAPACITY;
    if (w & 1) {
        const dim3 block(cw < 256 ? ((cw + 31) / 32) * 32 : 256);
        const dim3 grid((cw + block.x - 1) / block.x, crows, n);
        k_bgra_to_yuv_oddw<<<grid, block, 0, ctx->stream>>>(a);
    } else {
        const int groups = (w + 7) / 8;
        // the streaming kernels want whole 8-pixel groups and aligned rows: what encoders' frames have
        static const bool no_clip = yuv_bounds_ok(yuv_coef_bt601());     // what lets the streaming kernels drop min / max
        const bool aligned = no_clip && (w % 8) == 0 && ((uintptr_t)bgra % 16) == 0 && (stride % 16) == 0 && (bgra_pic_stride % 16) == 0 &&
                             ((uintptr_t)y % 8) == 0 && (ly % 8) == 0 && (y_pic_stride % 8) == 0 &&
                             ((uintptr_t)u % 4) == 0 && (lu % 4) == 0 && (u_pic_stride % 4) == 0 &&
                             ((uintptr_t)v % 4) == 0 && (lv % 4) == 0 && (v_pic_stride % 4) == 0;
        if (aligned && !v420) {
            const dim3 block(groups < 256 ? ((groups + 31) / 32) * 32 : 256);
            const dim3 grid((groups + block.x - 1) / block.x, h, n);
            k_bgra_to_yuv422_fast<<<grid, block, 0, ctx->stream>>>(a);
        } else if (aligned && v420 && bk->tile_ok) {
            const dim3 grid((groups + 31) / 32, (crows + kTileC - 1) / kTileC, n);
            k_bgra_to_yuv420_tiled<<<grid, 256, 0, ctx->stream>>>(a);
        } else {
            const dim3 block(groups < 256 ? ((groups + 31) / 32) * 32 : 256);
            const dim3 grid((groups + block.x - 1) / block.x, crows, n);
            k_bgra_to_yuv<<<grid, block, 0, ctx->stream>>>(a);
        }
    }
    CVS_CUDA(cudaGetLastError());
    ctx->launches++;
    return CVS_OK;
}

// The filter bank of one axis as the conversions use it (for tests and for hosts that want to see the taps).
int cvs_sws_bilinear_bank(int srcn, int dstn, int one, int32_t *pos, int32_t *coef, int max_taps) {
    if (srcn <= 0 || dstn <= 0 || one <= 0 || !pos || !coef) return CVS_ERR_INVALID_ARG;
    const FilterBank f = bilinear_bank(srcn, dstn, one);
    if (f.taps > max_taps) return CVS_ERR_CAPACITY;
    std::copy(f.pos.begin(), f.pos.end(), pos);
    std::copy(f.coef.begin(), f.coef.end(), coef);
    return f.taps;
}

int cvs_scale_to_bgra_device(cvs_ctx *ctx, void *dst, int dst_stride, long long dst_pic_stride, int dw, int dh,
                             const void *const src[3], const int src_linesize[3], const long long src_pic_stride[3],
                             int sw, int sh, int format, int n) {
    if (!ctx || !dst || !src || !src_linesize || !src_pic_stride || !src[0]) return CVS_ERR_INVALID_ARG;
    if (dw <= 0 || dh <= 0 || sw <= 0 || sh <= 0 || n < 0) return CVS_ERR_INVALID_ARG;
    if (format < CVS_PIX_BGRA || format > CVS_PIX_NV12) return CVS_ERR_INVALID_ARG;
    if (dst_stride < 4 * dw || (dst_stride & 3) || ((uintptr_t)dst & 3) || (dst_pic_stride & 3)) return CVS_ERR_INVALID_ARG;
    const int cw = (sw + 1) / 2, ch = format == CVS_PIX_YUV422P ? sh : (sh + 1) / 2;
    if (format == CVS_PIX_BGRA) {
        if (src_linesize[0] < 4 * sw) return CVS_ERR_INVALID_ARG;
    } else {
        if (src_linesize[0] < sw || !src[1]) return CVS_ERR_INVALID_ARG;
        if (format == CVS_PIX_NV12 ? src_linesize[1] < 2 * cw : (src_linesize[1] < cw || !src[2] || src_linesize[2] < cw))
            return CVS_ERR_INVALID_ARG;
    }
    if (sw > 16 * dw || sh > 16 * dh || dh > 65535 || n > 65535) return CVS_ERR_CAPACITY;
    if (n == 0) return CVS_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    ScalePlan *sp = nullptr;
    for (auto &q : ctx->scale_plans)
        if (q->dw == dw && q->dh == dh && q->sw == sw && q->sh == sh && q->format == format) sp = q.get();
    if (!sp) {
        std::unique_ptr<ScalePlan> q(new (std::nothrow) ScalePlan());
        if (!q) return CVS_ERR_NOMEM;
        q->dw = dw; q->dh = dh; q->sw = sw; q->sh = sh; q->format = format;
        // planar YUV: libswscale's own banks (sws_filter.cpp); BGRA sources: the repository's resampler
        q->sws = format != CVS_PIX_BGRA;
        q->full = q->sws && (dw & 1) != 0;         // odd width: the library's full-chroma writers, one chroma sample per pixel
        q->direct = q->sws && !q->full && format == CVS_PIX_YUV420P && sw == dw && sh == dh && (dh & 1) == 0;
        // BGRA at another size: the library's RGB -> YUV(A) -> RGB route, except the one geometry it reads past the row for
        q->rgb = format == CVS_PIX_BGRA && (sw != dw || sh != dh) && !((sw & 1) && dw <= (sw >> 1));
        q->rgb_half = q->rgb && dw <= (sw >> 1);
        if (q->rgb) {
            const int ccw = q->rgb_half ? sw / 2 : sw;
            const FilterBank banks[3] = {bilinear_bank(sw, dw, 1 << 14), bilinear_bank(ccw, dw, 1 << 14), bilinear_bank(sh, dh, 1 << 12)};
            const int srcn[3] = {sw, ccw, sh};
            std::vector<int32_t> blob;
            for (int i = 0; i < 3; i++) {
                for (int32_t v : banks[i].pos) if (v < 0 || v + banks[i].taps > srcn[i]) return CVS_ERR_UNSUPPORTED;
                q->sws_taps[i] = banks[i].taps;
                q->sws_pos[i] = blob.size(); blob.insert(blob.end(), banks[i].pos.begin(), banks[i].pos.end());
                q->sws_w[i] = blob.size(); blob.insert(blob.end(), banks[i].coef.begin(), banks[i].coef.end());
            }
            CVS_CUDA(dev_alloc(&q->d_sws, blob.size()));
            CVS_CUDA(cudaMemcpyAsync(q->d_sws, blob.data(), blob.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
            CVS_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        if (q->sws && !q->direct) {
            const FilterBank banks[4] = {bilinear_bank(sw, dw, 1 << 14), bilinear_bank(cw, q->full ? dw : (dw + 1) / 2, 1 << 14),
                                         bilinear_bank(sh, dh, 1 << 12), bilinear_bank(ch, dh, 1 << 12)};
            const int srcn[4] = {sw, cw, sh, ch};
            std::vector<int32_t> blob;
            for (int i = 0; i < 4; i++) {
                for (int32_t v : banks[i].pos) if (v < 0 || v + banks[i].taps > srcn[i]) return CVS_ERR_UNSUPPORTED;
                q->sws_taps[i] = banks[i].taps;
                q->sws_pos[i] = blob.size(); blob.insert(blob.end(), banks[i].pos.begin(), banks[i].pos.end());
                q->sws_w[i] = blob.size(); blob.insert(blob.end(), banks[i].coef.begin(), banks[i].coef.end());
            }
            CVS_CUDA(dev_alloc(&q->d_sws, blob.size()));
            CVS_CUDA(cudaMemcpyAsync(q->d_sws, blob.data(), blob.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
            CVS_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        if (!q->sws && !q->rgb) {
        ScaleAxis ax[4];
        const int suby = format == CVS_PIX_YUV422P ? 1 : 2;
        scale_build_axis(dw, sw, 1, 0, ax[0]);
        scale_build_axis(dh, sh, 1, 0, ax[1]);
        scale_build_axis(dw, sw, 2, 0, ax[2]);                       // chroma: co-sited with the even luma columns
        scale_build_axis(dh, sh, suby, suby == 2 ? 1 : 0, ax[3]);    // 4:2:0 chroma rows sit between two luma rows
        std::vector<int32_t> first;
        std::vector<int16_t> wts;
        for (int i = 0; i < 4; i++) {
            q->off_first[i] = first.size();
            q->off_w[i] = wts.size();
            q->taps[i] = ax[i].taps;
            first.insert(first.end(), ax[i].first.begin(), ax[i].first.end());
            wts.insert(wts.end(), ax[i].w.begin(), ax[i].w.end());
        }
        CVS_CUDA(dev_alloc(&q->d_first, first.size()));
        CVS_CUDA(dev_alloc(&q->d_w, wts.size()));
        CVS_CUDA(cudaMemcpyAsync(q->d_first, first.data(), first.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CVS_CUDA(cudaMemcpyAsync(q->d_w, wts.data(), wts.size() * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
        CVS_CUDA(cudaStreamSynchronize(ctx->stream));                // the tables are pageable temporaries
        }
        sp = q.get();
        ctx->scale_plans.push_back(std::move(q));
    }
    if (sp->rgb) {
        SwsRgbArgs r;
        r.dst = (uint8_t *)dst; r.dst_pic_stride = dst_pic_stride; r.dst_stride = dst_stride; r.dw = dw; r.dh = dh;
        r.src = (const uint8_t *)src[0]; r.sp = src_pic_stride[0]; r.ls = src_linesize[0]; r.half = sp->rgb_half;
        const int32_t *t = sp->d_sws;
        r.hl_pos = t + sp->sws_pos[0]; r.hl_w = t + sp->sws_w[0]; r.hc_pos = t + sp->sws_pos[1]; r.hc_w = t + sp->sws_w[1];
        r.v_pos = t + sp->sws_pos[2]; r.v_w = t + sp->sws_w[2];
        r.hl_t = sp->sws_taps[0]; r.hc_t = sp->sws_taps[1]; r.v_t = sp->sws_taps[2];
        const YuvCoef c = yuv_coef_bt601();
        r.ry = c.ry; r.gy = c.gy; r.by = c.by; r.ru = c.ru; r.gu = c.gu; r.bu = c.bu; r.rv = c.rv; r.gv = c.gv; r.bv = c.bv;
        const dim3 block(dw < 256 ? ((dw + 31) / 32) * 32 : 256);
        const dim3 grid((dw + block.x - 1) / block.x, dh, n);
        k_sws_bgra_to_bgra<<<grid, block, 0, ctx->stream>>>(r);
        CVS_CUDA(cudaGetLastError());
        ctx->launches++;
        return CVS_OK;
    }
    if (sp->sws) {
        SwsScaleArgs b;
        b.dst = (uint8_t *)dst; b.dst_pic_stride = dst_pic_stride; b.dst_stride = dst_stride; b.dw = dw; b.dh = dh;
        b.y = (const uint8_t *)src[0]; b.u = (const uint8_t *)src[1];
        b.v = format == CVS_PIX_NV12 ? (const uint8_t *)src[1] + 1 : (const uint8_t *)src[2];
        b.sp_y = src_pic_stride[0]; b.sp_c = src_pic_stride[1];
        b.ly = src_linesize[0]; b.lc = src_linesize[1]; b.cstep = format == CVS_PIX_NV12 ? 2 : 1;
        if (format != CVS_PIX_NV12 && (src_linesize[2] != src_linesize[1] || src_pic_stride[2] != src_pic_stride[1]))
            return CVS_ERR_UNSUPPORTED;                              // U and V planes of one layout (every decoder's)
        b.direct = sp->direct;
        const int32_t *t = sp->d_sws;
        b.hl_pos = t + sp->sws_pos[0]; b.hl_w = t + sp->sws_w[0]; b.hc_pos = t + sp->sws_pos[1]; b.hc_w = t + sp->sws_w[1];
        b.vl_pos = t + sp->sws_pos[2]; b.vl_w = t + sp->sws_w[2]; b.vc_pos = t + sp->sws_pos[3]; b.vc_w = t + sp->sws_w[3];
        b.hl_t = sp->sws_taps[0]; b.hc_t = sp->sws_taps[1]; b.vl_t = sp->sws_taps[2]; b.vc_t = sp->sws_taps[3];
        b.n = n;
        if (sp->full) {
            const dim3 block(dw < 256 ? ((dw + 31) / 32) * 32 : 256);
            const dim3 grid((dw + block.x - 1) / block.x, dh, n);
            k_sws_yuv_to_bgra_full<<<grid, block, 0, ctx->stream>>>(b);
            CVS_CUDA(cudaGetLastError());
            ctx->launches++;
            return CVS_OK;
        }
        const int pairs = dw / 2;
        const dim3 block(pairs < 256 ? ((pairs + 31) / 32) * 32 : 256);
        // at most two taps per axis and plane (same size or enlarging): the kernel that filters every source row once
        const bool few_taps = !sp->direct && b.hl_t <= 2 && b.hc_t <= 2 && b.vl_t <= 2 && b.vc_t <= 2 && ctx->sws_rows_kernel;
        if (few_taps) {
            const dim3 grid((pairs + block.x - 1) / block.x, (dh + kSwsRows - 1) / kSwsRows, n);
            k_sws_yuv_to_bgra_rows<<<grid, block, 0, ctx->stream>>>(b);
        } else {
            const dim3 grid((pairs + block.x - 1) / block.x, dh, n);
            k_sws_yuv_to_bgra<<<grid, block, 0, ctx->stream>>>(b);
        }
        CVS_CUDA(cudaGetLastError());
        ctx->launches++;
        return CVS_OK;
    }
    ScaleArgs a;
    a.dst = (uint8_t *)dst; a.dst_pic_stride = dst_pic_stride; a.dst_stride = dst_stride; a.dw = dw; a.dh = dh;
    for (int i = 0; i < 3; i++) {
        const bool used = i == 0 || (format != CVS_PIX_BGRA && (i == 1 || format != CVS_PIX_NV12));
        a.src[i] = used ? (const uint8_t *)src[i] : nullptr;
        a.src_pic_stride[i] = used ? src_pic_stride[i] : 0;
        a.src_linesize[i] = used ? src_linesize[i] : 0;
    }
    a.sw = sw; a.sh = sh; a.cw = cw; a.ch = ch; a.format = format; a.n = n;
    a.fx = sp->d_first + sp->off_first[0]; a.fy = sp->d_first + sp->off_first[1];
    a.cfx = sp->d_first + sp->off_first[2]; a.cfy = sp->d_first + sp->off_first[3];
    a.wx = sp->d_w + sp->off_w[0]; a.wy = sp->d_w + sp->off_w[1]; a.cwx = sp->d_w + sp->off_w[2]; a.cwy = sp->d_w + sp->off_w[3];
    a.tx = sp->taps[0]; a.ty = sp->taps[1]; a.ctx_ = sp->taps[2]; a.cty = sp->taps[3];
    const dim3 block(dw < 256 ? ((dw + 31) / 32) * 32 : 256);
    const dim3 grid((dw + block.x - 1) / block.x, dh, n);
    k_scale_to_bgra<<<grid, block, 0, ctx->stream>>>(a);
    CVS_CUDA(cudaGetLastError());
    ctx->launches++;
    return CVS_OK;
}

namespace {

// row h-1 of the pictures whose field does not own it keeps what the frame ring held (ffmpeg_ntsc.cpp:2247): with a
// ring of one picture that is the previous output picture's last row (or the row saved by the previous call)
__global__ void __launch_bounds__(256) k_inherit_last_row(uint8_t *out, size_t pic, int stride, int w, int h, int n,
                                                          unsigned long long first_fieldno, const uint8_t *saved) {
    const int k = blockIdx.y;
    const unsigned field = (unsigned)(((first_fieldno + (unsigned long long)k) & 1ull) ^ 1ull);
    if (k >= n || (unsigned)((h - 1) & 1) == field) return;              // the field wrote the row itself
    const uint32_t *from = reinterpret_cast<const uint32_t *>(k == 0 ? saved : out + (size_t)(k - 1) * pic + (size_t)(h - 1) * stride);
    uint32_t *to = reinterpret_cast<uint32_t *>(out + (size_t)k * pic + (size_t)(h - 1) * stride);
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x) to[x] = from[x];
}

cudaError_t grow(uint8_t **p, size_t *cap, size_t need, bool zero = false) {
    if (need <= *cap && *p) return cudaSuccess;
    cudaFree(*p);
    *p = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc((void **)p, need);
    if (e != cudaSuccess) return e;
    *cap = need;
    return zero ? cudaMemset(*p, 0, need) : cudaSuccess;
}

}  // namespace

int cvs_field_loop_host(cvs_ctx *ctx, const cvs_field_loop *d, int n, unsigned long long first_fieldno) {
    if (!ctx || !d || d->struct_size != (int32_t)sizeof(cvs_field_loop) || n < 0) return CVS_ERR_INVALID_ARG;
    const int w = d->w, h = d->h, sw = d->src_w, sh = d->src_h, fmt = d->src_format;
    if (w <= 0 || h <= 0 || sw <= 0 || sh <= 0 || d->nsrc <= 0 || !d->src[0] || !d->y || !d->u || !d->v) return CVS_ERR_INVALID_ARG;
    if (fmt < CVS_PIX_BGRA || fmt > CVS_PIX_NV12) return CVS_ERR_INVALID_ARG;
    if (d->out_format != CVS_YUV420P && d->out_format != CVS_YUV422P) return CVS_ERR_INVALID_ARG;
    const int cw = (w + 1) / 2, chh = d->out_format == CVS_YUV420P ? (h + 1) / 2 : h;
    if (d->ly < w || d->lu < cw || d->lv < cw) return CVS_ERR_INVALID_ARG;
    if (w > ctx->max_w || h > ctx->max_h || n > ctx->max_batch) return CVS_ERR_CAPACITY;
    if (d->src_of_field)
        for (int k = 0; k < n; k++)
            if (d->src_of_field[k] < 0 || d->src_of_field[k] >= d->nsrc) return CVS_ERR_INVALID_ARG;
    if (n == 0) return CVS_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    cvs_ctx *c = ctx;

    // source planes: rows packed to their own width on the device
    const int scw = (sw + 1) / 2, sch = fmt == CVS_PIX_YUV422P ? sh : (sh + 1) / 2;
    const int nplanes = fmt == CVS_PIX_BGRA ? 1 : (fmt == CVS_PIX_NV12 ? 2 : 3);
    int rowb[3] = {0, 0, 0}, rows[3] = {0, 0, 0};
    rowb[0] = fmt == CVS_PIX_BGRA ? 4 * sw : sw; rows[0] = sh;
    if (nplanes >= 2) { rowb[1] = fmt == CVS_PIX_NV12 ? 2 * scw : scw; rows[1] = sch; }
    if (nplanes == 3) { rowb[2] = scw; rows[2] = sch; }
    size_t plane_off[3] = {0, 0, 0}, src_pic = 0;
    int dl[3] = {0, 0, 0};
    for (int i = 0; i < nplanes; i++) {
        if (!d->src[i] || d->src_linesize[i] < rowb[i]) return CVS_ERR_INVALID_ARG;
        dl[i] = (rowb[i] + 15) / 16 * 16;
        plane_off[i] = src_pic;
        src_pic += (size_t)dl[i] * (size_t)rows[i];
    }
    const int dstride = ((4 * w + 15) / 16) * 16;
    const size_t dpic = (size_t)dstride * (size_t)h;
    const int yl = (w + 15) / 16 * 16, cl = (cw + 15) / 16 * 16;
    const size_t ypl = (size_t)yl * h, cpl = (size_t)cl * chh, yuv_pic = ypl + 2 * cpl;
    CVS_CUDA(cudaStreamSynchronize(c->stream));
    CVS_CUDA(cudaStreamSynchronize(c->s_out));
    CVS_CUDA(grow(&c->fl_src, &c->fl_src_cap, src_pic * (size_t)d->nsrc));
    CVS_CUDA(grow(&c->fl_scaled, &c->fl_scaled_cap, dpic * (size_t)d->nsrc));
    CVS_CUDA(grow(&c->fl_out, &c->fl_out_cap, dpic * (size_t)n));
    CVS_CUDA(grow(&c->fl_yuv, &c->fl_yuv_cap, yuv_pic * (size_t)n));
    CVS_CUDA(grow(&c->fl_last_row, &c->fl_last_row_cap, (size_t)dstride, true));     // the zeroed ring (:2069-2092)

    // 1. decoder pictures up (s_in) and 2. scaled to BGRA at the output size (frame_copy_scale) happen chunk by chunk below,
    // just ahead of the fields that need them: the uploads of later chunks then overlap the kernels and the downloads
    // of earlier ones (PCIe is full duplex; uploading everything first cost the upload time on top).
    // 3. per chunk: composite_layer() + line doubling, the inherited ring row, the encoder's YUV; 4. pictures down
    std::vector<int32_t> index((size_t)n);
    for (int k = 0; k < n; k++) index[(size_t)k] = d->src_of_field ? d->src_of_field[k] : (int32_t)(((long long)k * d->nsrc) / n);
    const int saved_bob = c->bob;
    const RandCursor call_start = c->cur;
    c->bob = 1;
    int rc = CVS_OK;
    const int chunk = c->host_chunk;
    int ci = 0, up = 0;                   // sources [0, up) are on the device and scaled
    for (int k0 = 0; k0 < n && rc == CVS_OK; k0 += chunk, ci++) {
        const int m = n - k0 < chunk ? n - k0 : chunk;
        while ((int)c->ev_k.size() <= ci) {
            cudaEvent_t a, b;
            if (cudaEventCreateWithFlags(&a, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&b, cudaEventDisableTiming) != cudaSuccess) { rc = CVS_ERR_CUDA; break; }
            c->ev_in.push_back(a);
            c->ev_k.push_back(b);
        }
        if (rc != CVS_OK) break;
        int need = 0;
        for (int k = k0; k < k0 + m; k++) need = index[(size_t)k] + 1 > need ? index[(size_t)k] + 1 : need;
        if (need > up) {
            cudaError_t e = cudaSuccess;
            for (int q = up; q < need && e == cudaSuccess; q++)
                for (int i = 0; i < nplanes && e == cudaSuccess; i++)
                    e = cudaMemcpy2DAsync(c->fl_src + (size_t)q * src_pic + plane_off[i], (size_t)dl[i],
                                          (const uint8_t *)d->src[i] + (size_t)q * (size_t)d->src_pic_stride[i], (size_t)d->src_linesize[i],
                                          (size_t)rowb[i], (size_t)rows[i], cudaMemcpyHostToDevice, c->s_in);
            if (e == cudaSuccess) e = cudaEventRecord(c->ev_in[ci], c->s_in);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->ev_in[ci], 0);
            if (e != cudaSuccess) { rc = CVS_ERR_CUDA; break; }
            const uint8_t *base = c->fl_src + (size_t)up * src_pic;
            const void *sp[3] = {base + plane_off[0], nplanes >= 2 ? base + plane_off[1] : nullptr, nplanes == 3 ? base + plane_off[2] : nullptr};
            const long long sps[3] = {(long long)src_pic, (long long)src_pic, (long long)src_pic};
            rc = cvs_scale_to_bgra_device(c, c->fl_scaled + (size_t)up * dpic, dstride, (long long)dpic, w, h, sp, dl, sps, sw, sh, fmt, need - up);
            if (rc != CVS_OK) break;
            up = need;
        }
        uint8_t *out = c->fl_out + (size_t)k0 * dpic;
        rc = run_device(c, out, dpic, dstride, c->fl_scaled, dpic, dstride, w, h, 0, 0, m, first_fieldno + (unsigned long long)k0, -1,
                        index.data() + k0);
        if (rc != CVS_OK) break;
        {
            const uint8_t *saved = k0 == 0 ? c->fl_last_row : out - dpic + (size_t)(h - 1) * dstride;
            const dim3 grid((unsigned)((w + 255) / 256 < 8 ? (w + 255) / 256 : 8), (unsigned)m);
            k_inherit_last_row<<<grid, 256, 0, c->stream>>>(out, dpic, dstride, w, h, m, first_fieldno + (unsigned long long)k0, saved);
            if (cudaGetLastError() != cudaSuccess) { rc = CVS_ERR_CUDA; break; }
            c->launches++;
        }
        uint8_t *yv = c->fl_yuv + (size_t)k0 * yuv_pic;
        rc = cvs_bgra_to_yuv_device(c, yv, yl, (long long)yuv_pic, yv + ypl, cl, (long long)yuv_pic, yv + ypl + cpl, cl, (long long)yuv_pic,
                                    out, dstride, (long long)dpic, w, h, m, d->out_format);
        if (rc != CVS_OK) break;
        if (cudaEventRecord(c->ev_k[ci], c->stream) != cudaSuccess || cudaStreamWaitEvent(c->s_out, c->ev_k[ci], 0) != cudaSuccess) { rc = CVS_ERR_CUDA; break; }
        for (int k = k0; k < k0 + m && rc == CVS_OK; k++) {
            const uint8_t *pk = c->fl_yuv + (size_t)k * yuv_pic;
            cudaError_t e = cudaMemcpy2DAsync((uint8_t *)d->y + (size_t)k * (size_t)d->y_pic_stride, (size_t)d->ly, pk, (size_t)yl, (size_t)w, (size_t)h,
                                              cudaMemcpyDeviceToHost, c->s_out);
            if (e == cudaSuccess) e = cudaMemcpy2DAsync((uint8_t *)d->u + (size_t)k * (size_t)d->u_pic_stride, (size_t)d->lu, pk + ypl, (size_t)cl, (size_t)cw,
                                                        (size_t)chh, cudaMemcpyDeviceToHost, c->s_out);
            if (e == cudaSuccess) e = cudaMemcpy2DAsync((uint8_t *)d->v + (size_t)k * (size_t)d->v_pic_stride, (size_t)d->lv, pk + ypl + cpl, (size_t)cl, (size_t)cw,
                                                        (size_t)chh, cudaMemcpyDeviceToHost, c->s_out);
            if (e != cudaSuccess) rc = CVS_ERR_CUDA;
        }
    }
    if (rc == CVS_OK && cudaMemcpyAsync(c->fl_last_row, c->fl_out + (size_t)(n - 1) * dpic + (size_t)(h - 1) * dstride, (size_t)dstride,
                                        cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) rc = CVS_ERR_CUDA;
    c->bob = saved_bob;
    cudaStreamSynchronize(c->s_in);
    cudaStreamSynchronize(c->s_out);
    if (rc != CVS_OK) { cudaStreamSynchronize(c->stream); c->cur = call_start; return rc; }
    return check_status(c);
}

int cvs_kernel_time_reset(cvs_ctx *ctx) {
    if (!ctx) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    CVS_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->ev_used = 0;
    return CVS_OK;
}

int cvs_kernel_time_query(cvs_ctx *ctx, double *total_ms, int *launches) {
    if (!ctx || !total_ms || !launches) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    CVS_CUDA(cudaStreamSynchronize(ctx->stream));
    double sum = 0;
    for (size_t i = 0; i < ctx->ev_used; i++) {
        float ms = 0;
        CVS_CUDA(cudaEventElapsedTime(&ms, ctx->ev_pool[i].first, ctx->ev_pool[i].second));
        sum += ms;
    }
    *total_ms = sum;
    *launches = (int)ctx->ev_used;
    return CVS_OK;
}

}  // extern "C"
