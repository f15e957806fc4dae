#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 scanline engine (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --gpus N ...          # the reference's own CPU code, host cores

Metric (BASELINE.json): fields/s at 1920x1080, VHS-SP preset (`-vhs -vhs-speed sp`).  One "frame"
of the reference's output is one FIELD: one composite_layer() call (ffmpeg_ntsc.cpp:2229) = one
output picture of ffmpeg_ntsc (540 processed scanlines, then line-doubled).  full-frames/s = value/2.

A step = one pass of the hot path over one batch of synthetic pictures:
  value  device-resident: `batch` pictures already in HBM, one cvs_composite_fields_device call/step
         (k_headswitch + k_fields), timed with CUDA events on the launching stream, max over ranks.
         The working set (batch x 8.3 MB x 2) is far larger than L2, so no L2 flush is needed.
         value_sustained: the same loop run for >= 2 s.
  e2e    same metric through the C ABI with HOST (pinned) buffers: cvs_composite_fields_host_async, H2D
         and D2H copies inside the timed region.  e2e.copy_ceiling: the same copies (sizes, pitches,
         streams) with no kernel between them, i.e. what this box can move at N ranks.
  parity the first fields of the stream, through the same host-buffer call, against the reference's own
         composite_layer() (oracle/_ref/libref.so) on the host: max |delta| per 8-bit channel.
  roofline  k_fields alone: algorithmic bytes (8 B per processed pixel) / mean launch duration from
         CUDA events recorded around every launch in the timed region (cvs_kernel_time_query).
  config.other_workloads  the other BASELINE configurations and the 4:2:2 sibling path, measured the same
         way (device-resident, kernel events), the whole field loop host to host (cvs_field_loop_host), and
         the two picture conversions either side of the path (`conversions`: libswscale's bytes), N = 1 only.
  cpu_baseline  the reference's own composite_layer() (oracle/_ref/libref.so, extracted at build
         time) on ONE host thread -- the reference is single-threaded -- median of 3 bounded samples.
Multi-GPU: fields are independent given the rand() position (closed-form seek), so rank r of N takes
the r-th contiguous chunk of every step's N*batch fields; no data-path collective.  The only
collective is one NCCL broadcast of the parameter block at start-up.  scaling = weak.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BARS = [0xC0C0C0, 0xC0C000, 0x00C0C0, 0x00C000, 0xC000C0, 0xC00000, 0x0000C0, 0x000000]
PRESETS = {"sp": ["-vhs", "-vhs-speed", "sp"], "ep": ["-vhs", "-vhs-speed", "ep"], "lp": ["-vhs", "-vhs-speed", "lp"],
           "comp": []}
KERNEL_NAME = {"sp": "VHS, 9", "lp": "VHS, 12", "ep": "VHS, 14", "comp": "composite"}


def metric_name(w, h, preset):
    what = "VHS-" + preset.upper() if preset != "comp" else "composite-only"
    return "fields/s at %dx%d %s (1 field = 1 composite_layer call = 1 output picture of ffmpeg_ntsc)" % (w, h, what)


def workload_name(w, h, preset):
    return "%dx%d %s (%s), synthetic perturbed colour bars" % (
        w, h, "VHS-" + preset.upper() if preset != "comp" else "composite only", " ".join(PRESETS[preset]) or "no switches")


# ----------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md recipe), sampled during the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# the reference's own CPU code (oracle/_ref) -- cpu_baseline leg, parity check and --impl reference arm.
# Nothing here imports the product package when oracle/_ref/libref.so exists: the parameter block is set by the
# harness's own ref_set_preset(), so the reference arm's process loads the reference's code and nothing else.
# ----------------------------------------------------------------------------------------------
def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def _load_cpu_checker(preset):
    """(call(dst, src, w, h, field, fieldno), kind): the extracted reference if it was built, else the oracle port."""
    ref = os.path.join(ROOT, "oracle", "_ref", "libref.so")
    if os.path.exists(ref):
        lib = C.CDLL(ref)
        if lib.ref_set_preset(preset.encode()) != 0:
            raise RuntimeError("unknown preset " + preset)
        lib.ref_srand(1)

        def call(dst, src, w, h, field, fieldno):
            lib.ref_composite_layer(dst.ctypes.data_as(C.c_void_p), 4 * w, src.ctypes.data_as(C.c_void_p), 4 * w,
                                    w, h, 0, 0, field, C.c_ulonglong(fieldno))
        return call, "reference"
    # the port: needs the parameter block (product library) -- only when the reference's code is not available
    import composite_video_simulator_b200 as cvs
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(orc):
        subprocess.check_call(["make", "liboracle.so"], cwd=os.path.join(ROOT, "oracle"), stdout=subprocess.DEVNULL)
    lib = C.CDLL(orc)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    p = cvs.params_from_argv(PRESETS[preset])
    g = helpers.OracleRng()
    lib.oracle_rng_seed(C.byref(g), 1)

    def call(dst, src, w, h, field, fieldno):
        lib.oracle_composite_layer(C.byref(p), C.byref(g), dst.ctypes.data_as(C.c_void_p), 4 * w,
                                   src.ctypes.data_as(C.c_void_p), 4 * w, w, h, 0, 0, field, C.c_ulonglong(fieldno))
    return call, "port"


def cpu_stream_frame(w, h, k):
    import numpy as np
    x = np.arange(w, dtype=np.uint32)[None, :]
    y = np.arange(h, dtype=np.uint32)[:, None]
    xs = (x + np.uint32(7 * k)) % np.uint32(w)
    base = np.broadcast_to(np.array(BARS, dtype=np.uint32)[(xs * np.uint32(8)) // np.uint32(w)], (h, w))
    pert = ((x * np.uint32(2654435761)) ^ (y * np.uint32(40503)) ^ np.uint32((k * 97) & 0xFFFFFFFF)) >> np.uint32(29)
    out = np.zeros((h, w), dtype=np.uint32)
    for sh in (0, 8, 16):
        c = ((base >> np.uint32(sh)) & np.uint32(0xFF)) + pert
        out |= np.minimum(c, 255).astype(np.uint32) << np.uint32(sh)
    return out


def cpu_run_fields(preset, w, h, nfields, first=0, keep=False):
    """nfields composite_layer() calls of the CPU checker on one thread, fields first.. of the synthetic stream,
    from the default seed.  Returns (seconds, kind, pictures or None); every field goes into a fresh picture when
    `keep` (the parity check), else into one reused picture as the reference's loop has it."""
    import numpy as np
    call, kind = _load_cpu_checker(preset)
    frames = [cpu_stream_frame(w, h, first + k) for k in range(nfields if keep else min(nfields, 4))]
    dst = np.zeros((h, w), dtype=np.uint32)
    pics = []
    t0 = time.perf_counter()
    for k in range(nfields):
        if keep:
            dst = np.zeros((h, w), dtype=np.uint32)
            pics.append(dst)
        call(dst, frames[k % len(frames)], w, h, ((first + k) & 1) ^ 1, first + k)
    return time.perf_counter() - t0, kind, (pics if keep else None)


def _ref_worker(arg):
    preset, w, h, nfields, first = arg
    dt, kind, _ = cpu_run_fields(preset, w, h, nfields, first)
    return dt, kind


def run_reference_arm(args):
    """The reference's CPU implementation on all host cores: P independent single-threaded processes
    (the reference has no threading), each a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    w, h, preset = args.width, args.height, args.preset
    cores = os.cpu_count() or 1
    per_proc = args.ref_fields_per_proc
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        kind = "reference"
        t_w = time.perf_counter()
        for _ in range(max(1, args.warmup)):
            pool.map(_ref_worker, [(preset, w, h, 1, 0)] * cores)
        t_field = (time.perf_counter() - t_w) / max(1, args.warmup)      # one field per process, all cores busy
        # bounded sample: keep the K timed steps within about two minutes of CPU wall time
        per_proc = max(1, min(per_proc, int(120.0 / (max(1, args.steps) * max(t_field, 1e-3)))))
        t0 = time.perf_counter()
        for s in range(args.steps):
            res = pool.map(_ref_worker, [(preset, w, h, per_proc, s * per_proc)] * cores)
            kind = res[0][1]
        dt = time.perf_counter() - t0
    total = cores * per_proc * args.steps
    val = total / dt
    line = {
        "impl": "reference", "metric": metric_name(w, h, preset), "value": val, "unit": "fields/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(w, h, preset),
                   "fields_per_step": cores * per_proc, "processes": cores, "cpu_model": cpu_model()},
        "cpu_baseline": {"value": val, "unit": "fields/s", "cores": cores, "kind": kind, "cpu_model": cpu_model(),
                         "sample": "%d processes x %d fields per step, %d steps, %s" % (cores, per_proc, args.steps, workload_name(w, h, preset))},
        "e2e": {"value": val, "unit": "fields/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------
def make_device_stream(torch, w, h, n, first, device):
    """Synthetic parity stream generated on the device: int32 BGRA, alpha 0."""
    x = torch.arange(w, device=device, dtype=torch.int64)[None, :]
    y = torch.arange(h, device=device, dtype=torch.int64)[:, None]
    bars = torch.tensor(BARS, device=device, dtype=torch.int64)
    out = torch.empty((n, h, w), device=device, dtype=torch.int32)
    for i in range(n):
        k = first + i
        xs = (x + 7 * k) % w
        base = bars[(xs * 8) // w].expand(h, w)
        pert = (((x * 2654435761) & 0xFFFFFFFF) ^ ((y * 40503) & 0xFFFFFFFF) ^ ((k * 97) & 0xFFFFFFFF)) >> 29
        pic = torch.zeros((h, w), device=device, dtype=torch.int64)
        for sh in (0, 8, 16):
            c = torch.clamp(((base >> sh) & 0xFF) + pert, max=255)
            pic |= c << sh
        out[i] = pic.to(torch.int32)
    return out


def bind_near_gpu(torch, local_rank):
    """Run this rank (and the threads and pinned buffers it creates from here on: first touch) on the CPUs of the
    GPU's NUMA node, as a host application feeding a GPU over PCIe would.  Returns a description, or None."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        dom, bus, devn = getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, devn)
        cpus = set()
        for part in open(path).read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cur = os.sched_getaffinity(0)
        near = cpus & cur
        if not near or near == cur:
            return None
        os.sched_setaffinity(0, near)
        return "%d CPUs of the GPU's NUMA node" % len(near)
    except Exception:
        return None


def hbm_peak():
    peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
            src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    return peak, src


def kernel_source_hash():
    """Identifies the kernel an ncu traffic capture belongs to (profiles/ncu_traffic.json is refused when stale)."""
    hsh = hashlib.sha256()
    for fn in ("lane_pipeline.cuh", "scanline_kernels.cuh"):
        with open(os.path.join(ROOT, "composite_video_simulator_b200", "csrc", fn), "rb") as f:
            hsh.update(f.read())
    return hsh.hexdigest()[:16]


class Timed:
    """K steps on `stream` between two CUDA events, barrier + synchronize on both sides, max over ranks."""

    def __init__(self, torch, dist, dev, stream):
        self.torch, self.dist, self.dev, self.stream = torch, dist, dev, stream
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e1 = torch.cuda.Event(enable_timing=True)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, eng, step, first, count):
        """-> (ms of the slowest rank, mean kernel ms per launch of the slowest rank, launches)"""
        torch = self.torch
        eng.synchronize()
        self.barrier()
        eng.kernel_time_reset()
        l0 = eng.kernel_launches()
        self.e0.record(self.stream)
        for i in range(count):
            step(first + i)
        self.e1.record(self.stream)
        self.e1.synchronize()
        eng.synchronize()
        self.barrier()
        kms, kn = eng.kernel_time_query()
        t = torch.tensor([self.e0.elapsed_time(self.e1), kms / max(kn, 1)], device=self.dev, dtype=torch.float64)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), eng.kernel_launches() - l0


def measure_other_bgra(torch, cvs, sharding, timed, dev, local_rank, name, w, h, preset, max_batch, steps, warmup, peak):
    """One more BASELINE configuration, device-resident, timed like the headline (N = 1)."""
    params = cvs.params_from_argv(PRESETS[preset])
    with cvs.Engine(params=params, device=local_rank, max_w=w, max_h=h, max_batch=max_batch) as eng:
        B = eng.preferred_batch(w, h, max_batch)
        eng.set_stream(timed.stream.cuda_stream)
        nsrc = min(B, 16)                                       # 16 distinct pictures, repeated: generation is slow in torch
        src16 = make_device_stream(torch, w, h, nsrc, 0, dev)
        src = src16.repeat((B + nsrc - 1) // nsrc, 1, 1)[:B].contiguous()
        dst = torch.zeros_like(src)

        def step(i):
            eng.rng_seek(sharding.stream_position(params, w, h, i * B))
            eng.composite_fields_device(dst, src, B, h, w, i * B)

        for i in range(warmup):
            step(i)
        ms, kms, launches = timed.run(eng, step, warmup, steps)
        eng.set_stream(0)
    nl = (h + 1) // 2
    alg = 8.0 * w * nl * B
    ach = alg / (kms / 1e3) / 1e9
    del src, dst
    torch.cuda.empty_cache()
    return {"workload": workload_name(w, h, preset), "value": B * steps / (ms / 1e3), "unit": "fields/s",
            "fields_per_step": B, "steps": steps, "kernel": "k_fields<float, %s, tv>" % KERNEL_NAME[preset],
            "kernel_ms_per_launch": kms, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "algorithmic_bytes_per_launch": alg}}


def measure_yuv422(torch, timed, dev, local_rank, w, h, argv, max_batch, steps, warmup, peak):
    """The 4:2:2 sibling path (include/cvs_yuv422.h, ffmpeg_to_composite.cpp:629-952), device-resident, in place."""
    import numpy as np
    from composite_video_simulator_b200 import yuv422
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    n = max_batch
    lc = w // 2 + 16
    Y0, U0, V0 = helpers.yuv422_frame(w, h, 0, 32)
    planes = [torch.from_numpy(np.ascontiguousarray(np.broadcast_to(p, (n,) + p.shape))).to(dev).contiguous()
              for p in (Y0, U0[:, :lc], V0[:, :lc])]
    with yuv422.Yuv422Engine(argv, device=local_rank, max_w=w, max_h=h, max_batch=n) as eng:
        eng.set_stream(timed.stream.cuda_stream)
        fno = [0]

        def step(i):
            eng.process_fields_device(planes[0], planes[1], planes[2], w, fno[0])
            fno[0] += n

        for i in range(warmup):
            step(i)
        ms, kms, launches = timed.run(eng, step, warmup, steps)
        eng.set_stream(0)
    alg = 2.0 * 2 * w * ((h + 1) // 2) * n                      # Y + U/2 + V/2 read and written
    ach = alg / (kms / 1e3) / 1e9
    del planes
    torch.cuda.empty_cache()
    return {"workload": "%dx%d planar 4:2:2, ffmpeg_to_composite %s, in place" % (w, h, " ".join(argv)),
            "value": n * steps / (ms / 1e3), "unit": "fields/s", "fields_per_step": n, "steps": steps,
            "kernel": "cvs422::k_yuv422_fast (fp64, bit-exact; four role warps per group of 31 rows)", "kernel_ms_per_launch": kms, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "algorithmic_bytes_per_launch": alg}}


def measure_field_loop(torch, cvs, local_rank, w, h, preset, n, calls):
    """The whole field loop around the seam through cvs_field_loop_host: NV12 decoder pictures (one per two fields) in
    pinned host memory -> scale -> composite_layer -> line doubling -> YUV 4:2:0 in pinned host memory.  Synchronous
    calls, wall clock; H2D and D2H inside."""
    import numpy as np
    params = cvs.params_from_argv(PRESETS[preset])
    nsrc = n // 2
    cw, ch = w // 2, h // 2
    Ys = torch.randint(16, 236, (nsrc, h, w), dtype=torch.uint8).pin_memory()
    UVs = torch.randint(16, 241, (nsrc, ch, 2 * cw), dtype=torch.uint8).pin_memory()
    y = torch.zeros((n, h, w), dtype=torch.uint8).pin_memory()
    u = torch.zeros((n, ch, cw), dtype=torch.uint8).pin_memory()
    v = torch.zeros((n, ch, cw), dtype=torch.uint8).pin_memory()
    idx = [k // 2 for k in range(n)]
    with cvs.Engine(params=params, device=local_rank, max_w=w, max_h=h, max_batch=n) as eng:
        l0 = eng.kernel_launches()
        eng.field_loop_host([Ys.numpy(), UVs.numpy()], w, h, 3, w, h, n, 0, y.numpy(), u.numpy(), v.numpy(), True, idx)
        t0 = time.perf_counter()
        for c in range(calls):
            eng.field_loop_host([Ys.numpy(), UVs.numpy()], w, h, 3, w, h, n, (c + 1) * n, y.numpy(), u.numpy(), v.numpy(), True, idx)
        dt = time.perf_counter() - t0
        launches = eng.kernel_launches() - l0
    h2d = nsrc * (w * h + ch * 2 * cw)
    d2h = n * (w * h + 2 * ch * cw)
    return {"workload": "%dx%d NV12 decoder pictures (1 per 2 fields) -> scale -> %s -> line doubling -> YUV 4:2:0, host to host, "
                        "%d fields per call" % (w, h, workload_name(w, h, preset), n),
            "value": calls * n / dt, "unit": "fields/s", "h2d_bytes_per_call": h2d, "d2h_bytes_per_call": d2h,
            "pcie_bytes_per_field": (h2d + d2h) / n, "gpu_launches": launches,
            "checksum": int(y.numpy()[-1].astype(np.uint32).sum() & 0xFFFFFFFF)}


def measure_conversions(torch, cvs, local_rank, peak, steps=10):
    """The two picture conversions either side of the path (libswscale's bytes, tests/test_swscale_pin.py) as
    device-resident streaming kernels: 1080p pictures, CUDA events on the launching stream, algorithmic bytes = 4 B/px BGRA
    + the planar bytes of the other side."""
    out = {}
    w, h, n = 1920, 1080, 128
    src = torch.randint(0, 1 << 24, (n, h, w), dtype=torch.int32, device="cuda")
    with cvs.Engine([], device=local_rank, max_w=w, max_h=h, max_batch=1) as eng:
        st = torch.cuda.Stream()
        eng.set_stream(st.cuda_stream)

        def timed_launches(fn):
            with torch.cuda.stream(st):
                for _ in range(3):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(steps):
                    fn()
                e1.record(st)
                e1.synchronize()
            return e0.elapsed_time(e1) / steps

        for v420 in (True, False):
            ch = h // 2 if v420 else h
            y = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
            u = torch.empty((n, ch, w // 2), dtype=torch.uint8, device="cuda")
            v = torch.empty_like(u)
            ms = timed_launches(lambda: eng.bgra_to_yuv_device(y, u, v, src, w, h, n, fmt420=v420))
            gbs = n * w * h * (4 + (1.5 if v420 else 2.0)) / (ms / 1e3) / 1e9
            out["bgra_to_%s_1080p" % ("yuv420p" if v420 else "yuv422p")] = {
                "kernel": "k_bgra_to_yuv420_tiled" if v420 else "k_bgra_to_yuv422_fast", "value": n / (ms / 1e3), "unit": "pictures/s",
                "kernel_ms_per_launch": ms, "pictures_per_launch": n, "gpu_launches": steps,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak if peak else None}}
            del y, u, v
        for sw, sh in ((720, 480), (1920, 1080)):
            ns = 64
            Y = torch.randint(0, 256, (ns, sh, sw), dtype=torch.uint8, device="cuda")
            UV = torch.randint(0, 256, (ns, sh // 2, sw), dtype=torch.uint8, device="cuda")
            dst = torch.empty((ns, h, w), dtype=torch.int32, device="cuda")
            ms = timed_launches(lambda: eng.scale_to_bgra_device(dst, w, h, [Y, UV], [sw, sw], sw, sh, 3, n=ns))
            gbs = ns * (sw * sh * 1.5 + w * h * 4.0) / (ms / 1e3) / 1e9
            out["nv12_%dx%d_to_bgra_1080p" % (sw, sh)] = {
                "kernel": "k_sws_yuv_to_bgra", "value": ns / (ms / 1e3), "unit": "pictures/s", "kernel_ms_per_launch": ms,
                "pictures_per_launch": ns, "gpu_launches": steps,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak if peak else None}}
            del Y, UV, dst
        eng.set_stream(0)
    return out


def copy_ceiling(torch, dist, dev, w, h, Be, steps, chunk, contiguous=False):
    """What the box moves with the e2e path's copies alone: per field the rows of one parity up (pitch 2 x stride ->
    the device picture) and down, in chunks on two streams, no kernel in between.  Same sizes, pitches and pinned
    buffers per rank as the e2e measurement; all ranks at once.  -> GB/s per direction (aggregate over ranks).
    contiguous=True moves the same number of bytes per field as ONE linear copy each way (the link itself)."""
    from cuda.bindings import runtime as rt  # cuda-python (image); plumbing only
    nl = h // 2
    hsrc = torch.empty((Be, h, w), dtype=torch.int32).pin_memory()
    hdst = torch.empty((Be, h, w), dtype=torch.int32).pin_memory()
    dsrc = torch.empty((Be, h, w), dtype=torch.int32, device=dev)
    ddst = torch.empty((Be, h, w), dtype=torch.int32, device=dev)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    H2D, D2H = rt.cudaMemcpyKind.cudaMemcpyHostToDevice, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost
    pic = 4 * w * h

    def one_step():
        for k0 in range(0, Be, chunk):
            for k in range(k0, min(k0 + chunk, Be)):
                f = (k & 1) ^ 1
                if contiguous:
                    rt.cudaMemcpyAsync(dsrc.data_ptr() + k * pic, hsrc.data_ptr() + k * pic, 4 * w * nl, H2D, s_in.cuda_stream)
                else:
                    rt.cudaMemcpy2DAsync(dsrc.data_ptr() + k * pic + f * 4 * w, 8 * w, hsrc.data_ptr() + k * pic + f * 4 * w,
                                         8 * w, 4 * w, nl, H2D, s_in.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(s_in)
            s_out.wait_event(ev)
            for k in range(k0, min(k0 + chunk, Be)):
                f = (k & 1) ^ 1
                if contiguous:
                    rt.cudaMemcpyAsync(hdst.data_ptr() + k * pic, ddst.data_ptr() + k * pic, 4 * w * nl, D2H, s_out.cuda_stream)
                else:
                    rt.cudaMemcpy2DAsync(hdst.data_ptr() + k * pic + f * 4 * w, 8 * w, ddst.data_ptr() + k * pic + f * 4 * w,
                                         8 * w, 4 * w, nl, D2H, s_out.cuda_stream)

    one_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    world = 1
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        world = dist.get_world_size()
    del hsrc, hdst, dsrc, ddst
    return world * steps * Be * nl * 4 * w / float(t[0]) / 1e9


def run_own_arm(args):
    import numpy as np
    import torch
    import composite_video_simulator_b200 as cvs
    from composite_video_simulator_b200.params import CvsParams
    from composite_video_simulator_b200 import sharding

    W, H, preset = args.width, args.height, args.preset
    argv = PRESETS[preset]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_near_gpu(torch, local_rank) if args.numa_bind else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # the only collective of the whole job: rank 0 broadcasts the parameter block (the preset tables)
    pbytes = torch.zeros(C.sizeof(CvsParams), dtype=torch.uint8, device=dev)
    if rank == 0:
        p0 = cvs.params_from_argv(argv)
        pbytes.copy_(torch.frombuffer(bytearray(bytes(p0)), dtype=torch.uint8))
    if dist is not None:
        dist.broadcast(pbytes, src=0)
    params = CvsParams.from_buffer_copy(bytes(pbytes.cpu().numpy().tobytes()))

    nl = (H + 1) // 2
    eng = cvs.Engine(params=params, device=local_rank, max_w=W, max_h=H, max_batch=max(args.batch, args.e2e_batch))
    eng.set_noise_mode(args.noise == "fast")
    # fields per step: the largest batch <= --batch that fills whole waves of the GPU (a lane = a scanline,
    # all tasks equally long: a partial last wave is pure loss); identical on every rank
    B = eng.preferred_batch(W, H, args.batch) if args.wave_align else args.batch
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)
    timed = Timed(torch, dist, dev, stream)

    src = make_device_stream(torch, W, H, B, rank * B, dev)
    dst = torch.zeros_like(src)
    torch.cuda.synchronize()

    def step(i):
        base, _ = sharding.chunk(i, rank, world, B)   # this rank's contiguous chunk of the global field stream
        eng.rng_seek(sharding.stream_position(params, W, H, base))
        eng.composite_fields_device(dst, src, B, H, W, base)

    for i in range(args.warmup):
        step(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_max, kernel_ms, launches = timed.run(eng, step, args.warmup, args.steps)
    value = world * B * args.steps / (ms_max / 1e3)

    # ---- multi-GPU identity: the last rank's last chunk, recomputed by rank 0 on its own GPU, must be the same bytes
    identity = None
    if world > 1:
        last = args.warmup + args.steps - 1
        mine = torch.stack([dst.view(torch.int64).sum(), dst[:, 1::2].to(torch.int64).sum()])
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        if rank == 0:
            base, _ = sharding.chunk(last, world - 1, world, B)
            src_o = make_device_stream(torch, W, H, B, (world - 1) * B, dev)
            dst_o = torch.zeros_like(src_o)
            eng.rng_seek(sharding.stream_position(params, W, H, base))
            eng.composite_fields_device(dst_o, src_o, B, H, W, base)
            eng.synchronize()
            again = torch.stack([dst_o.view(torch.int64).sum(), dst_o[:, 1::2].to(torch.int64).sum()])
            identity = {"what": "rank %d's chunk of the last step (%d fields from field %d) recomputed on rank 0's GPU"
                                % (world - 1, B, base), "byte_checksums_equal": bool(torch.equal(again, sums[world - 1]))}
            del src_o, dst_o
        timed.barrier()

    # ---- sustained: the same loop for >= 2 s ----
    sustained = None
    if args.sustain_s > 0:
        n_s = max(args.steps, int(args.sustain_s * 1e3 / (ms_max / args.steps)) + 1)
        ms_s, kms_s, _ = timed.run(eng, step, args.warmup, n_s)
        sustained = {"value": world * B * n_s / (ms_s / 1e3), "unit": "fields/s", "steps": n_s, "seconds": ms_s / 1e3,
                     "kernel_ms_per_launch": kms_s}

    # ---- side measurement (not the headline): the other per-pixel noise mode, same steps, same timing ----
    # (cvs_set_noise_mode: CVS_NOISE_FAST draws the per-pixel noise from counter generators, within +-1 LSB)
    other = None
    if args.noise_side:
        eng.set_noise_mode(args.noise != "fast")
        for i in range(args.warmup):
            step(i)
        ms_o, kms_o, _ = timed.run(eng, step, args.warmup, args.steps)
        alg_o = 8.0 * W * nl * B
        other = {"noise": "exact" if args.noise == "fast" else "fast (counter generators, +-1 LSB)",
                 "value": world * B * args.steps / (ms_o / 1e3), "unit": "fields/s", "kernel_ms_per_launch": kms_o,
                 "roofline_frac": alg_o / (kms_o / 1e3) / 1e9 / hbm_peak()[0]}
        eng.set_noise_mode(args.noise == "fast")
        eng.kernel_time_reset()

    # ---- e2e: host buffers through the C ABI (H2D + kernels + D2H inside the timed region) ----
    Be = args.e2e_batch
    hsrc = [torch.empty((Be, H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
    hdst = [torch.zeros((Be, H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
    first = src[:Be].cpu() if Be <= B else make_device_stream(torch, W, H, Be, rank * Be, dev).cpu()
    for t_ in hsrc:
        t_.copy_(first)
    hs_np = [t_.numpy().view(np.uint32) for t_ in hsrc]
    hd_np = [t_.numpy().view(np.uint32) for t_ in hdst]

    def e2e_step(i):
        # every step uploads its own Be pictures from pinned host memory and downloads its Be results;
        # consecutive steps are queued asynchronously over two buffer sets (cvs_composite_fields_host_async)
        # and the timed region ends with a full synchronize
        base, _ = sharding.chunk(i, rank, world, Be)
        eng.rng_seek(sharding.stream_position(params, W, H, base))
        # (the engine orders call i after the downloads of call i-2, which used the same device buffers;
        # the host buffers are only read back after the final synchronize)
        eng.composite_fields_host_async(hd_np[i % 2], hs_np[i % 2], base)

    for i in range(2):
        e2e_step(i)
    eng.synchronize()
    timed.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, args.steps // 2)
    for i in range(e2e_steps):
        e2e_step(i)
    eng.synchronize()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * Be * e2e_steps / float(t[0])
    e2e_checksum = int(hd_np[(e2e_steps - 1) % 2][0, 1::2].sum() & 0xFFFFFFFF)    # the step's result is read on the host
    clocks = sampler.stop() if rank == 0 else None           # sampled over the timed regions above

    # ---- the copies alone (all ranks together): what the box can move ----
    ceiling = ceiling_lin = None
    if args.copy_probe:
        try:
            ceiling = copy_ceiling(torch, dist, dev, W, H, Be, e2e_steps, 32)
            ceiling_lin = copy_ceiling(torch, dist, dev, W, H, Be, e2e_steps, 32, contiguous=True)
        except Exception as ex:       # cuda-python missing: the probe is optional
            ceiling = ceiling_lin = None
            sys.stderr.write("copy probe unavailable: %r\n" % (ex,))
    del hsrc, hdst

    # ---- parity of what was timed: the first fields of the stream through the same host-buffer entry point,
    # against the reference's own composite_layer() on the host (rank 0; position 0 of rand()) ----
    parity = None
    if rank == 0 and args.parity_fields > 0:
        npar = args.parity_fields
        par_src = np.stack([cpu_stream_frame(W, H, k) for k in range(npar)])
        par_dst = np.zeros_like(par_src)
        eng.rng_seek(0)
        eng.composite_fields_host(par_dst, par_src, 0)
        _, kind, want = cpu_run_fields(preset, W, H, npar, 0, keep=True)
        d = np.abs(np.stack(want).view(np.uint8).astype(np.int16) - par_dst.view(np.uint8).astype(np.int16))
        rows = np.zeros((npar, H), dtype=bool)
        for k in range(npar):
            rows[k, ((k & 1) ^ 1)::2] = True
        nvals = int(rows.sum()) * W * 4
        parity = {"fields": npar, "max_abs_delta": int(d.max()), "frac_moved": float((d > 0).sum()) / nvals,
                  "values_off_by_more_than_1": int((d > 1).sum()), "against": kind,
                  "through": "cvs_composite_fields_host (the e2e entry point), noise mode %s" % args.noise}

    if rank != 0:
        if dist is not None:
            # the other workloads run on rank 0 only; wait so that the process group is torn down together
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_fields) ----
    peak, peak_src = hbm_peak()
    alg_bytes = 8.0 * W * nl * B                               # 4 B read + 4 B written per processed pixel
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    traffic, traffic_note = None, "no ncu capture of this kernel build (profiles/ncu_traffic.json missing or stale)"
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)        # one `ncu --set full` capture (scripts/round_end_gpu.sh), stamped with the kernel sources' hash
        key = "%s_%dx%d" % (preset, W, H)
        if tj.get("kernel_source_sha256") == kernel_source_hash() and key in tj.get("captures", {}):
            cap = tj["captures"][key]
            traffic = cap["dram_bytes"] / cap["fields"] * B
            traffic_note = "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (%d fields), scaled to %d" % (cap["fields"], B)
    except Exception:
        pass

    # ---- the other BASELINE configurations and the 4:2:2 path (N = 1) ----
    others = None
    if world == 1 and args.other_workloads:
        eng.set_stream(0)
        del src, dst
        torch.cuda.empty_cache()
        st2 = max(5, args.steps // 2)
        others = {}
        for name, (w2, h2, pr2, mb) in (("ep_1080p", (1920, 1080, "ep", 960)), ("comp_2160p", (3840, 2160, "comp", 240)),
                                        ("sp_480p", (720, 480, "sp", 3072))):
            if (w2, h2, pr2) == (W, H, preset):
                continue
            others[name] = measure_other_bgra(torch, cvs, sharding, timed, dev, local_rank, name, w2, h2, pr2, mb, st2,
                                              args.warmup, peak)
        # side measurements: a failure here is reported in its key and never takes the headline line with it
        def side(key, fn):
            try:
                others[key] = fn()
            except Exception as e:      # noqa: BLE001
                others[key] = {"error": "%s: %s" % (type(e).__name__, e)}
                torch.cuda.empty_cache()
        side("field_loop_nv12_to_yuv420p_1080p", lambda: measure_field_loop(torch, cvs, local_rank, 1920, 1080, "sp", 256, 3))
        side("conversions", lambda: measure_conversions(torch, cvs, local_rank, peak))
        side("yuv422_sp_1080p", lambda: measure_yuv422(torch, timed, dev, local_rank, 1920, 1080, ["-vhs", "-vhs-speed", "sp"],
                                                      339, st2, args.warmup, peak))

    # ---- CPU baseline: the reference's own code, one thread, median of 3 bounded samples ----
    cpu = None
    if world == 1 and args.cpu_fields > 0:
        per = max(2, args.cpu_fields // 3)
        runs = []
        kind = "reference"
        for r in range(3):
            dtc, kind, _ = cpu_run_fields(preset, W, H, per, r * per)
            runs.append(per / dtc)
        cpu = {"value": statistics.median(runs), "unit": "fields/s", "cores": 1, "kind": kind, "cpu_model": cpu_model(),
               "host_cores": os.cpu_count(), "samples": runs,
               "sample": "median of 3 runs of %d consecutive fields of the synthetic stream, %s, 1 thread"
                         % (per, workload_name(W, H, preset))}

    line = {
        "metric": metric_name(W, H, preset), "value": value, "unit": "fields/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "value_sustained": sustained,
        "config": {"workload": workload_name(W, H, preset) + ", %d fields per GPU per step, generated on device" % B,
                   "fields_per_step_per_gpu": B, "full_frames_per_s": value / 2,
                   "l2": "inputs larger than L2 (%.0f MB read + %.0f MB written per step)" % (alg_bytes / 2e6, alg_bytes / 2e6),
                   "noise": ("exact glibc rand() replay" if args.noise == "exact" else
                             "fast mode: per-pixel noise from counter generators (+-1 LSB), per-line draws exact"),
                   "other_noise_mode": other, "host_affinity": numa or "unchanged",
                   "parallelism": "fields sharded by contiguous chunk, no data-path collective",
                   "other_workloads": others},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "fields/s", "h2d_bytes_per_step": Be * nl * 4 * W,
                "d2h_bytes_per_step": Be * nl * 4 * W, "fields_per_step_per_gpu": Be, "result_checksum": e2e_checksum,
                "gbs_each_way": e2e_value * nl * 4 * W / 1e9,
                "copy_ceiling_gbs_each_way": ceiling, "copy_ceiling_contiguous_gbs_each_way": ceiling_lin,
                "frac_of_copy_ceiling": (e2e_value * nl * 4 * W / 1e9 / ceiling) if ceiling else None},
        "parity": parity,
        "multi_gpu_identity": identity,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "k_fields<float, %s, tv>" % KERNEL_NAME[preset], "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_per_launch": kernel_ms},
        "cpu_baseline": cpu,
    }
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def _claim_stdout():
    """Exactly ONE line may reach stdout (the JSON result).  Libraries print there too (NCCL's version
    banner does), so file descriptor 1 is pointed at stderr for the whole run and the result line is
    written to the saved descriptor at the end."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--batch", type=int, default=960,
                    help="upper bound of fields per GPU per step (device-resident); 960 -> 915 fields = 9 whole waves of the GPU: "
                         "every launch ends with a tail of partly idle SMs, 3 waves per launch measured 3.5 %% slower than 9")
    ap.add_argument("--no-wave-align", dest="wave_align", action="store_false",
                    help="use --batch as is instead of the wave-aligned batch cvs_preferred_batch() suggests")
    ap.add_argument("--e2e-batch", type=int, default=128, help="fields per GPU per step (host buffers; 4 pinned buffers of this many pictures per rank)")
    ap.add_argument("--cpu-fields", type=int, default=72, help="fields of the single-thread CPU baseline (3 samples of a third each)")
    ap.add_argument("--parity-fields", type=int, default=8, help="fields compared with the reference's CPU code (0 = skip)")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the sustained run (0 = skip)")
    ap.add_argument("--ref-fields-per-proc", type=int, default=4)
    ap.add_argument("--no-numa-bind", dest="numa_bind", action="store_false",
                    help="do not move the process to the CPUs of the GPU's NUMA node")
    ap.add_argument("--noise", default="exact", choices=["exact", "fast"],
                    help="per-pixel noise source of the measured runs (cvs_set_noise_mode); the headline is exact")
    ap.add_argument("--no-noise-side", dest="noise_side", action="store_false",
                    help="skip the side measurement of the other noise mode")
    ap.add_argument("--no-other-workloads", dest="other_workloads", action="store_false",
                    help="skip config.other_workloads (the other BASELINE configurations and the 4:2:2 path)")
    ap.add_argument("--no-copy-probe", dest="copy_probe", action="store_false", help="skip e2e.copy_ceiling")
    ap.add_argument("--quick", action="store_true", help="kernel experiments: headline loop only")
    ap.add_argument("--width", type=int, default=1920, help="experiments only: the headline metric is 1920x1080")
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--preset", default="sp", choices=["sp", "ep", "lp", "comp"],
                    help="sp = the headline VHS-SP workload; the others are for kernel experiments")
    args = ap.parse_args()
    if args.quick:
        args.other_workloads = args.copy_probe = False
        args.cpu_fields = args.parity_fields = 0
        args.sustain_s = 0
        args.e2e_batch = min(args.e2e_batch, 16)
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    try:
        return run_own_arm(args)
    except BaseException:
        import traceback
        sys.stderr.write("bench.py rank %s failed:\n%s\n" % (os.environ.get("RANK", "0"), traceback.format_exc()))
        sys.stderr.flush()
        raise


if __name__ == "__main__":
    sys.exit(main())
