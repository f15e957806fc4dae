// yuv422_api.cu -- the C ABI of include/cvs_yuv422.h on top of the 4:2:2 kernels.
//
// Host side of the second drop-in boundary (composite_video_process() / render_field() of
// ffmpeg_to_composite.cpp): owns the CUDA stream, the device-side tables and the position in the libc
// rand() stream, plans each batch (yuv422_plan.cpp) and launches k_yuv422_halo + k_yuv422_headswitch +
// k_yuv422.  Nothing here computes pixels on the CPU: if CUDA is not usable every compute entry point
// fails with CVS_ERR_CUDA.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <thread>
#include <memory>
#include <new>
#include <vector>

#include "../../include/cvs_yuv422.h"
#include "glibc_rand.h"
#include "yuv422_kernels.cuh"
#include "yuv422_plan.h"

using namespace cvs422;
using cvs::FieldSide;
using cvs::GeomPlan;
using cvs::RandCursor;

namespace {

constexpr int kSlots = 2;
constexpr size_t kMaxEventPairs = 4096;

#define CVS_CUDA(x)                                  \
    do {                                             \
        cudaError_t e_ = (x);                        \
        if (e_ != cudaSuccess) return CVS_ERR_CUDA;  \
    } while (0)

struct DevPlan422 {
    int w = 0, h = 0;
    unsigned field = 0;
    GeomPlan g;
    uint32_t *d_seek = nullptr;
};

struct Slot {
    FieldDesc422 *h_fields = nullptr, *d_fields = nullptr;   // h_*: pinned
    uint32_t *h_rowinfo = nullptr, *d_rowinfo = nullptr;
    int32_t *h_hsshift = nullptr, *d_hsshift = nullptr;
    HsItem422 *h_items = nullptr, *d_items = nullptr;
    cudaEvent_t uploaded = nullptr;        // the table copies out of this slot's pinned buffers have finished
    cudaEvent_t kernel_done = nullptr;     // the kernels reading this slot's device tables have finished
    bool in_flight = false;
};

int head_switch_rows_bound(int w) {
    int shif = (w + w / 10) / 2 + 1, n = 0;
    while (shif != 0) { shif = (shif * 7) / 8; n++; }
    return n + 1;
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

struct cvs422_ctx {
    cvs422_params p;
    int device = 0;
    int max_w = 0, max_h = 0, max_batch = 0, nl_max = 0, wpf_max = 0, hs_max = 0, halo_pitch_max = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaStream_t s_tab = nullptr;                 // uploads the per-batch side tables while the previous batch computes
    RandCursor cur;
    int packed_rows = 1;                          // CVS_PACKED_ROWS=0: per-field warps (experiments / tests)
    std::vector<std::unique_ptr<DevPlan422>> plans;
    Slot slots[kSlots];
    int next_slot = 0;
    uint8_t *d_scratch = nullptr, *d_halo = nullptr;
    int32_t *d_status = nullptr, *h_status = nullptr;
    int rotate_roles = 0;              // CVS422_ROTATE=1: the role assignment rotates from group to group on an SM (measured slower)
    double *d_lut = nullptr;
    int plan_threads = 4;                         // host threads that build the per-row side tables of a batch (CVS_PLAN_THREADS)
    size_t lut_cap = 0;
    std::vector<double> lut_host;                 // what d_lut currently holds
    uint8_t *d_planes = nullptr;                  // device pictures of the host-pointer entry points
    size_t d_planes_cap = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    size_t ev_used = 0;
    unsigned long long launches = 0;
};

namespace {

void free_all(cvs422_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto &pl : c->plans) if (pl && pl->d_seek) cudaFree(pl->d_seek);
    c->plans.clear();
    for (auto &s : c->slots) {
        if (s.h_fields) cudaFreeHost(s.h_fields);
        if (s.h_rowinfo) cudaFreeHost(s.h_rowinfo);
        if (s.h_hsshift) cudaFreeHost(s.h_hsshift);
        if (s.h_items) cudaFreeHost(s.h_items);
        cudaFree(s.d_fields); cudaFree(s.d_rowinfo); cudaFree(s.d_hsshift); cudaFree(s.d_items);
        if (s.uploaded) cudaEventDestroy(s.uploaded);
        if (s.kernel_done) cudaEventDestroy(s.kernel_done);
        s = Slot();
    }
    cudaFree(c->d_scratch); cudaFree(c->d_halo); cudaFree(c->d_status); cudaFree(c->d_lut); cudaFree(c->d_planes);
    if (c->h_status) cudaFreeHost(c->h_status);
    for (auto &pr : c->ev_pool) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    c->ev_pool.clear();
    if (c->s_tab) cudaStreamDestroy(c->s_tab);
    if (c->stream && c->own_stream) cudaStreamDestroy(c->stream);
}

int get_plan(cvs422_ctx *c, int w, int h, unsigned field, DevPlan422 **out) {
    for (auto &pl : c->plans)
        if (pl->w == w && pl->h == h && pl->field == field) { *out = pl.get(); return CVS_OK; }
    std::unique_ptr<DevPlan422> pl(new (std::nothrow) DevPlan422());
    if (!pl) return CVS_ERR_NOMEM;
    pl->w = w; pl->h = h; pl->field = field;
    build_geom_plan(c->p, w, h, field, pl->g);
    const size_t words = pl->g.seek.size() > 0 ? pl->g.seek.size() : 1;
    CVS_CUDA(cudaMalloc((void **)&pl->d_seek, words * sizeof(uint32_t)));
    if (!pl->g.seek.empty())
        CVS_CUDA(cudaMemcpyAsync(pl->d_seek, pl->g.seek.data(), pl->g.seek.size() * sizeof(uint32_t),
                                 cudaMemcpyHostToDevice, c->stream));
    CVS_CUDA(cudaStreamSynchronize(c->stream));
    *out = pl.get();
    c->plans.push_back(std::move(pl));
    return CVS_OK;
}

void drop_plans(cvs422_ctx *c) {
    cudaStreamSynchronize(c->stream);
    for (auto &pl : c->plans) if (pl && pl->d_seek) cudaFree(pl->d_seek);
    c->plans.clear();
}

// n consecutive fields on device planes, asynchronous on c->stream.  Picture k: planes at y + k*psy etc.
int run_device(cvs422_ctx *c, uint8_t *y, uint8_t *u, uint8_t *v, long long psy, long long psu, long long psv,
               int ly, int lu, int lv, int w, int h, int n, unsigned long long first_fieldno) {
    if (!y || !u || !v || w <= 0 || h <= 0 || n < 0) return CVS_ERR_INVALID_ARG;
    if (ly < w || lu < w / 2 || lv < w / 2) return CVS_ERR_INVALID_ARG;
    if (w > c->max_w || h > c->max_h || n > c->max_batch) return CVS_ERR_CAPACITY;
    if (n == 0) return CVS_OK;
    if (!c->p.enable_composite_emulation) {                   // the reference skips the call (:1789)
        return CVS_OK;
    }
    // a call that fails leaves the rand() position where it was (include/cvs_yuv422.h)
    struct CursorGuard {
        cvs422_ctx *c; RandCursor start; bool armed;
        ~CursorGuard() { if (armed) c->cur = start; }
    } guard{c, c->cur, true};
    Launch422 a;
    std::vector<double> lut;
    int rc = make_k422(c->p, w, h, a.K, a.dv, lut);
    if (rc != CVS_OK) return rc;
    // (bytes, not values: the byte tables behind the cos / sin pairs can look like NaNs, which never compare equal)
    const bool same = lut.size() == c->lut_host.size() && (lut.empty() || std::memcmp(lut.data(), c->lut_host.data(), lut.size() * sizeof(double)) == 0);
    if (!same || !c->d_lut) {
        if (lut.size() > c->lut_cap || !c->d_lut) {
            if (c->d_lut) { CVS_CUDA(cudaStreamSynchronize(c->stream)); cudaFree(c->d_lut); c->d_lut = nullptr; }
            c->lut_cap = lut.size() > 64 ? lut.size() : 64;
            CVS_CUDA(cudaMalloc((void **)&c->d_lut, c->lut_cap * sizeof(double)));
        }
        CVS_CUDA(cudaStreamSynchronize(c->stream));         // the previous table may still be in use
        c->lut_host = lut;
        CVS_CUDA(cudaMemcpyAsync(c->d_lut, c->lut_host.data(), lut.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CVS_CUDA(cudaStreamSynchronize(c->stream));
    }
    a.K.phase_lut = c->d_lut;

    DevPlan422 *plan[2] = {nullptr, nullptr};
    Slot &sl = c->slots[c->next_slot];
    c->next_slot = (c->next_slot + 1) % kSlots;
    if (sl.in_flight) {
        CVS_CUDA(cudaEventSynchronize(sl.uploaded));            // pinned buffers free again
        CVS_CUDA(cudaStreamWaitEvent(c->s_tab, sl.kernel_done, 0));   // device tables free again
        sl.in_flight = false;
    }

    const int nl_max = (h + 1) / 2;
    const int wpf = (nl_max + kRowsPerWarp - 1) / kRowsPerWarp;
    const int halo_y = round_up(w + 2, 16), halo_c = round_up(w / 2, 16);
    const int halo_pitch = halo_y + 2 * halo_c;
    int nitems = 0, total_rows = 0, max_nl = 0, min_nl = 1 << 30;
    // pass 1 (sequential, cheap): where every field starts in the rand() stream and in the batch's sequence of rows
    std::vector<RandCursor> at((size_t)n);
    for (int k = 0; k < n; k++) {
        const unsigned long long fieldno = first_fieldno + (unsigned long long)k;
        const unsigned field = (unsigned)((fieldno & 1ULL) ^ 1ULL);          // bottom field first (:1784)
        if (!plan[field]) { rc = get_plan(c, w, h, field, &plan[field]); if (rc != CVS_OK) return rc; }
        const GeomPlan &g = plan[field]->g;
        at[(size_t)k] = c->cur;
        c->cur.jump(g.jumpN, g.ndraws);
        FieldDesc422 &fd = sl.h_fields[k];
        fd.row_start = total_rows; fd.pad_ = 0;
        total_rows += g.nl;
        if (g.nl > max_nl) max_nl = g.nl;
        if (g.nl < min_nl) min_nl = g.nl;
    }
    // pass 2: the per-row side tables of the fields, on a few host threads (a field costs ~12 us; one thread kept a
    // 339-field launch waiting once the kernel took less than 4 ms)
    std::atomic<bool> capacity_error{false};
    auto plan_range = [&](int k0, int k1) {
        FieldSide fs;
        for (int k = k0; k < k1; k++) {
            const unsigned long long fieldno = first_fieldno + (unsigned long long)k;
            const unsigned field = (unsigned)((fieldno & 1ULL) ^ 1ULL);
            const GeomPlan &g = plan[field]->g;
            build_field_side_at(c->p, g, at[(size_t)k], fs);
            FieldDesc422 &fd = sl.h_fields[k];
            fd.y = y + (long long)k * psy; fd.u = u + (long long)k * psu; fd.v = v + (long long)k * psv;
            fd.rowinfo = sl.d_rowinfo + (size_t)k * (size_t)c->nl_max;
            fd.seek = plan[field]->d_seek;
            fd.hs_scratch = c->d_scratch + (size_t)k * (size_t)c->hs_max * (size_t)c->max_w;
            fd.hs_shift = sl.d_hsshift + (size_t)k * (size_t)c->hs_max;
            fd.halo = c->d_halo + (size_t)k * (size_t)c->wpf_max * (size_t)c->halo_pitch_max;
            fd.fieldno = fieldno;
            fd.field = (int32_t)field; fd.nl = g.nl; fd.hs_first = fs.hs_first; fd.hs_count = fs.hs_count;
            std::memcpy(fd.window, fs.window, sizeof(fs.window));
            if (g.nl > 0) std::memcpy(sl.h_rowinfo + (size_t)k * (size_t)c->nl_max, fs.rowinfo.data(), (size_t)g.nl * sizeof(uint32_t));
            if (fs.hs_count > c->hs_max) { capacity_error = true; fd.hs_count = 0; continue; }
            for (int i = 0; i < fs.hs_count; i++) sl.h_hsshift[(size_t)k * (size_t)c->hs_max + (size_t)i] = fs.hs_shift[(size_t)i];
        }
    };
    const int nthreads = (n >= 64) ? c->plan_threads : 1;
    if (nthreads <= 1) {
        plan_range(0, n);
    } else {
        // a thread that cannot be started (std::system_error must not cross the C ABI) leaves its range, and every
        // later one, to this thread
        std::vector<std::thread> pool;
        const int per = (n + nthreads - 1) / nthreads;
        int done_to = per < n ? per : n;
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
        for (int k0 = done_to; k0 < n; k0 += per) plan_range(k0, k0 + per < n ? k0 + per : n);
    }
    if (capacity_error.load()) return CVS_ERR_CAPACITY;
    for (int k = 0; k < n; k++)
        for (int i = 0; i < sl.h_fields[k].hs_count; i++) {
            sl.h_items[nitems].field_idx = k;
            sl.h_items[nitems].slot = i;
            nitems++;
        }
    CVS_CUDA(cudaMemcpyAsync(sl.d_fields, sl.h_fields, (size_t)n * sizeof(FieldDesc422), cudaMemcpyHostToDevice, c->s_tab));
    CVS_CUDA(cudaMemcpyAsync(sl.d_rowinfo, sl.h_rowinfo, (size_t)n * (size_t)c->nl_max * sizeof(uint32_t), cudaMemcpyHostToDevice, c->s_tab));
    if (nitems > 0) {
        CVS_CUDA(cudaMemcpyAsync(sl.d_hsshift, sl.h_hsshift, (size_t)n * (size_t)c->hs_max * sizeof(int32_t), cudaMemcpyHostToDevice, c->s_tab));
        CVS_CUDA(cudaMemcpyAsync(sl.d_items, sl.h_items, (size_t)nitems * sizeof(HsItem422), cudaMemcpyHostToDevice, c->s_tab));
    }
    CVS_CUDA(cudaEventRecord(sl.uploaded, c->s_tab));
    sl.in_flight = true;
    CVS_CUDA(cudaStreamWaitEvent(c->stream, sl.uploaded, 0));

    a.fields = sl.d_fields;
    a.nfields = n;
    a.warps_per_field = wpf;
    a.total_warps = n * wpf;
    // packed mapping: a warp of 31 consecutive rows of the batch then meets at most two fields
    a.packed = (c->packed_rows && n > 1 && min_nl > kRowsPerWarp) ? 1 : 0;
    a.total_rows = total_rows;
    a.max_nl = max_nl > 0 ? max_nl : 1;
    a.halo_base = c->d_halo;
    if (a.packed) a.total_warps = (total_rows + kRowsPerWarp - 1) / kRowsPerWarp;
    a.ly = ly; a.lu = lu; a.lv = lv;
    a.by = (long long)ly * h; a.bu = (long long)lu * h; a.bv = (long long)lv * h;
    a.halo_pitch = halo_pitch; a.halo_u = halo_y; a.halo_v = halo_y + halo_c;
    const auto al = [](const void *p, long long s, int ls, int m) {
        return ((uintptr_t)p % (uintptr_t)m) == 0 && (s % m) == 0 && (ls % m) == 0;
    };
    a.vec = al(y, psy, ly, 8) && al(u, psu, lu, 4) && al(v, psv, lv, 4);
    a.status = c->d_status;
    a.sm_slots = c->d_status + 1;
    a.rotate = c->rotate_roles;

    // the head-switch pre-pass and the halo copy read the pictures before the in-place pass writes them
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->ev_used < kMaxEventPairs) {
        if (c->ev_used == c->ev_pool.size()) {
            cudaEvent_t a0, a1;
            CVS_CUDA(cudaEventCreate(&a0));
            CVS_CUDA(cudaEventCreate(&a1));
            c->ev_pool.emplace_back(a0, a1);
        }
        e0 = c->ev_pool[c->ev_used].first;
        e1 = c->ev_pool[c->ev_used].second;
        c->ev_used++;
        CVS_CUDA(cudaEventRecord(e0, c->stream));
    }
    CVS_CUDA(launch_yuv422(a, sl.d_items, nitems, c->stream));
    if (e1) CVS_CUDA(cudaEventRecord(e1, c->stream));
    CVS_CUDA(cudaEventRecord(sl.kernel_done, c->stream));
    c->launches += 1 + ((a.packed ? a.total_warps > 1 : wpf > 1) ? 1 : 0) + (nitems > 0 ? 1 : 0);
    guard.armed = false;
    return CVS_OK;
}

int check_status(cvs422_ctx *c) {
    CVS_CUDA(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CVS_CUDA(cudaStreamSynchronize(c->stream));
    if (*c->h_status != 0) {
        CVS_CUDA(cudaMemsetAsync(c->d_status, 0, sizeof(int32_t), c->stream));
        return CVS_ERR_NOISE_SYNC;
    }
    return CVS_OK;
}

}  // namespace

extern "C" {

int cvs422_create(cvs422_ctx **out, const cvs422_params *p, int device, int max_w, int max_h, int max_batch) {
    if (!out || !p || max_w <= 0 || max_h <= 0 || max_batch <= 0) return CVS_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return CVS_ERR_CUDA;   // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return CVS_ERR_CUDA;
    cvs422_ctx *c = new (std::nothrow) cvs422_ctx();
    if (!c) return CVS_ERR_NOMEM;
    {
        unsigned hw = std::thread::hardware_concurrency();
        if (hw > 0 && (int)hw < c->plan_threads) c->plan_threads = (int)hw;
        if (const char *e = std::getenv("CVS_PLAN_THREADS")) {
            const int v = std::atoi(e);
            if (v >= 1 && v <= 64) c->plan_threads = v;
        }
    }
    c->p = *p;
    c->device = device;
    c->max_w = max_w; c->max_h = max_h; c->max_batch = max_batch;
    c->nl_max = (max_h + 1) / 2;
    c->wpf_max = (c->nl_max + kRowsPerWarp - 1) / kRowsPerWarp;
    c->hs_max = head_switch_rows_bound(max_w);
    c->halo_pitch_max = round_up(max_w + 2, 16) + 2 * round_up(max_w / 2, 16);
    c->cur.seed(1);
    if (const char *e = std::getenv("CVS_PACKED_ROWS")) c->packed_rows = std::atoi(e) != 0;
    if (const char *e = std::getenv("CVS422_ROTATE")) c->rotate_roles = std::atoi(e) != 0;
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&c->s_tab, cudaStreamNonBlocking) == cudaSuccess;
    for (auto &s : c->slots) {
        if (!ok) break;
        const size_t nb = (size_t)max_batch;
        ok = ok && cudaMallocHost((void **)&s.h_fields, nb * sizeof(FieldDesc422)) == cudaSuccess;
        ok = ok && cudaMallocHost((void **)&s.h_rowinfo, nb * (size_t)c->nl_max * sizeof(uint32_t)) == cudaSuccess;
        ok = ok && cudaMallocHost((void **)&s.h_hsshift, nb * (size_t)c->hs_max * sizeof(int32_t)) == cudaSuccess;
        ok = ok && cudaMallocHost((void **)&s.h_items, nb * (size_t)c->hs_max * sizeof(HsItem422)) == cudaSuccess;
        ok = ok && cudaMalloc((void **)&s.d_fields, nb * sizeof(FieldDesc422)) == cudaSuccess;
        ok = ok && cudaMalloc((void **)&s.d_rowinfo, nb * (size_t)c->nl_max * sizeof(uint32_t)) == cudaSuccess;
        ok = ok && cudaMalloc((void **)&s.d_hsshift, nb * (size_t)c->hs_max * sizeof(int32_t)) == cudaSuccess;
        ok = ok && cudaMalloc((void **)&s.d_items, nb * (size_t)c->hs_max * sizeof(HsItem422)) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&s.kernel_done, cudaEventDisableTiming) == cudaSuccess;
        if (ok) std::memset(s.h_hsshift, 0, nb * (size_t)c->hs_max * sizeof(int32_t));
    }
    ok = ok && cudaMalloc((void **)&c->d_scratch, (size_t)max_batch * (size_t)c->hs_max * (size_t)max_w) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&c->d_halo, (size_t)max_batch * (size_t)c->wpf_max * (size_t)c->halo_pitch_max) == cudaSuccess;
    // d_status[0]: the noise warm-up flag; d_status[1 ..]: per-SM arrival counters (role rotation, yuv422_kernels.cuh)
    ok = ok && cudaMalloc((void **)&c->d_status, (size_t)(1 + kSmSlots) * sizeof(int32_t)) == cudaSuccess;
    ok = ok && cudaMallocHost((void **)&c->h_status, sizeof(int32_t)) == cudaSuccess;
    ok = ok && cudaMemsetAsync(c->d_status, 0, (size_t)(1 + kSmSlots) * sizeof(int32_t), c->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess;
    if (!ok) { free_all(c); delete c; return CVS_ERR_CUDA; }
    *out = c;
    return CVS_OK;
}

void cvs422_destroy(cvs422_ctx *ctx) {
    if (!ctx) return;
    free_all(ctx);
    delete ctx;
}

int cvs422_set_params(cvs422_ctx *ctx, const cvs422_params *p) {
    if (!ctx || !p) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    drop_plans(ctx);                       // the draw layout depends on the parameters
    ctx->p = *p;
    return CVS_OK;
}

int cvs422_process_fields_device(cvs422_ctx *ctx, uint8_t *y, uint8_t *u, uint8_t *v,
                                 long long pic_stride_y, long long pic_stride_u, long long pic_stride_v,
                                 int linesize_y, int linesize_u, int linesize_v,
                                 int w, int h, int n, unsigned long long first_fieldno) {
    if (!ctx) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    return run_device(ctx, y, u, v, pic_stride_y, pic_stride_u, pic_stride_v, linesize_y, linesize_u, linesize_v,
                      w, h, n, first_fieldno);
}

int cvs422_process_fields_host(cvs422_ctx *ctx, uint8_t *y, uint8_t *u, uint8_t *v,
                               long long psy, long long psu, long long psv, int ly, int lu, int lv,
                               int w, int h, int n, unsigned long long first_fieldno) {
    if (!ctx || !y || !u || !v || w <= 0 || h <= 0 || n < 0) return CVS_ERR_INVALID_ARG;
    if (ly < w || lu < w / 2 || lv < w / 2 || (w & 1)) return CVS_ERR_INVALID_ARG;
    if (w > ctx->max_w || h > ctx->max_h || n > ctx->max_batch) return CVS_ERR_CAPACITY;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    if (n == 0 || !ctx->p.enable_composite_emulation) return CVS_OK;
    // device pictures mirror the host layout (same linesizes), one after the other, 256-byte aligned
    const size_t sy = (size_t)round_up(ly * h, 256), su = (size_t)round_up(lu * h, 256), sv = (size_t)round_up(lv * h, 256);
    const size_t pic = sy + su + sv, need = pic * (size_t)n;
    if (need > ctx->d_planes_cap) {
        CVS_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_planes);
        ctx->d_planes = nullptr;
        ctx->d_planes_cap = 0;
        CVS_CUDA(cudaMalloc((void **)&ctx->d_planes, need));
        ctx->d_planes_cap = need;
    }
    uint8_t *dy = ctx->d_planes, *du = dy + sy, *dv = du + su;
    const int cw = w / 2;
    const int ywid = ly < w + 2 ? ly : w + 2;               // the two bytes past a row travel with it
    const bool whole = ly < w + 2;                           // ... unless they are the start of the next row
    for (int k = 0; k < n; k++) {
        const unsigned field = (unsigned)(((first_fieldno + (unsigned long long)k) & 1ULL) ^ 1ULL);
        const int rows = (h > (int)field) ? (h - (int)field + 1) / 2 : 0;
        if (rows == 0) continue;
        if (whole) {
            CVS_CUDA(cudaMemcpyAsync(dy + (size_t)k * pic, y + (long long)k * psy, (size_t)ly * h, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            CVS_CUDA(cudaMemcpy2DAsync(dy + (size_t)k * pic + (size_t)field * ly, 2 * (size_t)ly,
                                       y + (long long)k * psy + (size_t)field * ly, 2 * (size_t)ly, (size_t)ywid, (size_t)rows,
                                       cudaMemcpyHostToDevice, ctx->stream));
        }
        CVS_CUDA(cudaMemcpy2DAsync(du + (size_t)k * pic + (size_t)field * lu, 2 * (size_t)lu,
                                   u + (long long)k * psu + (size_t)field * lu, 2 * (size_t)lu, (size_t)cw, (size_t)rows,
                                   cudaMemcpyHostToDevice, ctx->stream));
        CVS_CUDA(cudaMemcpy2DAsync(dv + (size_t)k * pic + (size_t)field * lv, 2 * (size_t)lv,
                                   v + (long long)k * psv + (size_t)field * lv, 2 * (size_t)lv, (size_t)cw, (size_t)rows,
                                   cudaMemcpyHostToDevice, ctx->stream));
    }
    int rc = run_device(ctx, dy, du, dv, (long long)pic, (long long)pic, (long long)pic, ly, lu, lv, w, h, n, first_fieldno);
    if (rc != CVS_OK) return rc;
    for (int k = 0; k < n; k++) {
        const unsigned field = (unsigned)(((first_fieldno + (unsigned long long)k) & 1ULL) ^ 1ULL);
        const int rows = (h > (int)field) ? (h - (int)field + 1) / 2 : 0;
        if (rows == 0) continue;
        CVS_CUDA(cudaMemcpy2DAsync(y + (long long)k * psy + (size_t)field * ly, 2 * (size_t)ly,
                                   dy + (size_t)k * pic + (size_t)field * ly, 2 * (size_t)ly, (size_t)w, (size_t)rows,
                                   cudaMemcpyDeviceToHost, ctx->stream));
        CVS_CUDA(cudaMemcpy2DAsync(u + (long long)k * psu + (size_t)field * lu, 2 * (size_t)lu,
                                   du + (size_t)k * pic + (size_t)field * lu, 2 * (size_t)lu, (size_t)cw, (size_t)rows,
                                   cudaMemcpyDeviceToHost, ctx->stream));
        CVS_CUDA(cudaMemcpy2DAsync(v + (long long)k * psv + (size_t)field * lv, 2 * (size_t)lv,
                                   dv + (size_t)k * pic + (size_t)field * lv, 2 * (size_t)lv, (size_t)cw, (size_t)rows,
                                   cudaMemcpyDeviceToHost, ctx->stream));
    }
    return check_status(ctx);
}

int cvs422_composite_video_process(cvs422_ctx *ctx, uint8_t *y, int ly, uint8_t *u, int lu, uint8_t *v, int lv,
                                   int w, int h, unsigned field, unsigned long long fieldno) {
    // the reference's call site always passes field == (fieldno & 1) ^ 1 (:1790); anything else is not
    // expressible through the batch form and is rejected rather than silently reinterpreted
    if (field != (unsigned)((fieldno & 1ULL) ^ 1ULL)) return CVS_ERR_INVALID_ARG;
    return cvs422_process_fields_host(ctx, y, u, v, 0, 0, 0, ly, lu, lv, w, h, 1, fieldno);
}

int cvs422_render_field_device(cvs422_ctx *ctx, uint8_t *const dst[3], const int dst_linesize[3], int dst_h,
                               const uint8_t *const src[3], const int src_linesize[3], int src_h,
                               const int row_bytes[3], int src_is_420,
                               int src_interlaced, int src_top_field_first, int second_field, unsigned field) {
    if (!ctx || !dst || !src || !dst_linesize || !src_linesize || !row_bytes) return CVS_ERR_INVALID_ARG;
    if (dst_h <= 0 || src_h < 2 || field > 1) return CVS_ERR_INVALID_ARG;
    if (src_is_420 && src_interlaced && (src_h >> 1) < 2) return CVS_ERR_INVALID_ARG;
    if ((long long)dst_h * 256 * (long long)src_h >= (1LL << 32)) return CVS_ERR_INVALID_ARG;   // the 8.8 row index is 32-bit (:1022)
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    RenderArgs a;
    for (int p = 0; p < 3; p++) {
        if (!dst[p] || !src[p] || row_bytes[p] < 0 || dst_linesize[p] < row_bytes[p] || src_linesize[p] < row_bytes[p])
            return CVS_ERR_INVALID_ARG;
        a.dst[p] = dst[p]; a.src[p] = src[p];
        a.dst_ls[p] = dst_linesize[p]; a.src_ls[p] = src_linesize[p]; a.row_bytes[p] = row_bytes[p];
    }
    a.dst_h = dst_h; a.src_h = src_h; a.is420 = src_is_420 != 0; a.interlaced = src_interlaced != 0;
    a.tff = src_top_field_first != 0; a.second = second_field != 0; a.field = (int32_t)field;
    CVS_CUDA(launch_render_field(a, ctx->stream));
    ctx->launches++;
    return CVS_OK;
}

int cvs422_synchronize(cvs422_ctx *ctx) {
    if (!ctx) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    return check_status(ctx);
}

int cvs422_set_stream(cvs422_ctx *ctx, void *cuda_stream) {
    if (!ctx) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    CVS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; }
    else { CVS_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    return CVS_OK;
}

int cvs422_rng_seek(cvs422_ctx *ctx, unsigned long long draws_consumed) {
    if (!ctx) return CVS_ERR_INVALID_ARG;
    ctx->cur.seek(draws_consumed);
    return CVS_OK;
}

unsigned long long cvs422_rng_tell(const cvs422_ctx *ctx) { return ctx ? (unsigned long long)ctx->cur.pos() : 0ULL; }

unsigned long long cvs422_draws_per_field(const cvs422_params *p, int w, int h, unsigned field) {
    if (!p || w <= 0 || h <= 0) return 0;
    const unsigned long long nl = (h > (int)field) ? (unsigned long long)((h - (int)field + 1) / 2) : 0ULL;
    unsigned long long n = 0;
    if (p->video_noise != 0) n += nl * (unsigned long long)w;
    if (p->vhs_head_switching && p->vhs_head_switching_phase_noise != 0) n += 4;
    if (p->video_chroma_noise != 0) n += nl * 2ULL * (unsigned long long)(w / 2);
    if (p->video_chroma_phase_noise != 0) n += nl;
    if (p->video_chroma_loss != 0) n += nl;
    return n;
}

unsigned long long cvs422_kernel_launches(const cvs422_ctx *ctx) { return ctx ? ctx->launches : 0ULL; }

int cvs422_kernel_time_reset(cvs422_ctx *ctx) {
    if (!ctx) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    CVS_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->ev_used = 0;
    return CVS_OK;
}

int cvs422_kernel_time_query(cvs422_ctx *ctx, double *total_ms, int *launches) {
    if (!ctx || !total_ms || !launches) return CVS_ERR_INVALID_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return CVS_ERR_CUDA;
    CVS_CUDA(cudaStreamSynchronize(ctx->stream));
    double t = 0;
    for (size_t i = 0; i < ctx->ev_used; i++) {
        float ms = 0;
        CVS_CUDA(cudaEventElapsedTime(&ms, ctx->ev_pool[i].first, ctx->ev_pool[i].second));
        t += ms;
    }
    *total_ms = t;
    *launches = (int)ctx->ev_used;
    return CVS_OK;
}

}  // extern "C"
