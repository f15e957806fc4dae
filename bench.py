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
  e2e    same metric through the C ABI with HOST (pinned) buffers: cvs_composite_fields_host, H2D
         and D2H copies inside the timed region.
  roofline  k_fields alone: algorithmic bytes (8 B per processed pixel) / mean launch duration from
         CUDA events recorded around every launch in the timed region (cvs_kernel_time_query).
  cpu_baseline  the reference's own composite_layer() (oracle/_ref/libref.so, extracted at build
         time) on ONE host thread -- the reference is single-threaded -- on a bounded sample.
Multi-GPU: fields are independent given the rand() position (closed-form seek), so rank r of N takes
the r-th contiguous chunk of every step's N*batch fields; no data-path collective.  The only
collective is one NCCL broadcast of the parameter block at start-up.  scaling = weak.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
ARGV = ["-vhs", "-vhs-speed", "sp"]
METRIC = "fields/s at 1920x1080 VHS-SP (1 field = 1 composite_layer call = 1 output picture of ffmpeg_ntsc)"
BARS = [0xC0C0C0, 0xC0C000, 0x00C0C0, 0x00C000, 0xC000C0, 0xC00000, 0x0000C0, 0x000000]


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
# the reference's own CPU code (oracle/_ref) -- cpu_baseline leg and --impl reference arm
# ----------------------------------------------------------------------------------------------
def _load_cpu_checker():
    """(lib, kind): the extracted reference if it was built, else the oracle port."""
    import composite_video_simulator_b200  # noqa: F401  (params struct)
    ref = os.path.join(ROOT, "oracle", "_ref", "libref.so")
    if os.path.exists(ref):
        return C.CDLL(ref), "reference"
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(orc):
        subprocess.check_call(["make", "liboracle.so"], cwd=os.path.join(ROOT, "oracle"), stdout=subprocess.DEVNULL)
    return C.CDLL(orc), "port"


def _cpu_stream_frame(k):
    import numpy as np
    x = np.arange(W, dtype=np.uint32)[None, :]
    y = np.arange(H, dtype=np.uint32)[:, None]
    xs = (x + np.uint32(7 * k)) % np.uint32(W)
    base = np.broadcast_to(np.array(BARS, dtype=np.uint32)[(xs * np.uint32(8)) // np.uint32(W)], (H, W))
    pert = ((x * np.uint32(2654435761)) ^ (y * np.uint32(40503)) ^ np.uint32((k * 97) & 0xFFFFFFFF)) >> np.uint32(29)
    out = np.zeros((H, W), dtype=np.uint32)
    for sh in (0, 8, 16):
        c = ((base >> np.uint32(sh)) & np.uint32(0xFF)) + pert
        out |= np.minimum(c, 255).astype(np.uint32) << np.uint32(sh)
    return out


def _cpu_run_fields(nfields, first=0):
    """Run nfields composite_layer() calls of the CPU checker on one thread; returns seconds."""
    import numpy as np
    import composite_video_simulator_b200 as cvs
    lib, kind = _load_cpu_checker()
    p = cvs.params_from_argv(ARGV)
    frames = [_cpu_stream_frame(first + k) for k in range(min(nfields, 4))]
    dst = np.zeros((H, W), dtype=np.uint32)
    if kind == "reference":
        lib.ref_set_params(C.byref(p))
        t0 = time.perf_counter()
        for k in range(nfields):
            src = frames[k % len(frames)]
            lib.ref_composite_layer(dst.ctypes.data_as(C.c_void_p), 4 * W, src.ctypes.data_as(C.c_void_p), 4 * W,
                                    W, H, 0, 0, ((first + k) & 1) ^ 1, C.c_ulonglong(first + k))
        return time.perf_counter() - t0, kind
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    g = helpers.OracleRng()
    lib.oracle_rng_seed(C.byref(g), 1)
    t0 = time.perf_counter()
    for k in range(nfields):
        src = frames[k % len(frames)]
        lib.oracle_composite_layer(C.byref(p), C.byref(g), dst.ctypes.data_as(C.c_void_p), 4 * W,
                                   src.ctypes.data_as(C.c_void_p), 4 * W, W, H, 0, 0, ((first + k) & 1) ^ 1,
                                   C.c_ulonglong(first + k))
    return time.perf_counter() - t0, kind


def _ref_worker(arg):
    nfields, first = arg
    dt, kind = _cpu_run_fields(nfields, first)
    return dt, kind


def run_reference_arm(args):
    """The reference's CPU implementation on all host cores: P independent single-threaded processes
    (the reference has no threading), each a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_proc = args.ref_fields_per_proc
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        kind = "reference"
        t_w = time.perf_counter()
        for _ in range(max(1, args.warmup)):
            pool.map(_ref_worker, [(1, 0)] * cores)
        t_field = (time.perf_counter() - t_w) / max(1, args.warmup)      # one field per process, all cores busy
        # bounded sample: keep the K timed steps within about two minutes of CPU wall time
        per_proc = max(1, min(per_proc, int(120.0 / (max(1, args.steps) * max(t_field, 1e-3)))))
        t0 = time.perf_counter()
        for s in range(args.steps):
            res = pool.map(_ref_worker, [(per_proc, s * per_proc)] * cores)
            kind = res[0][1]
        dt = time.perf_counter() - t0
    total = cores * per_proc * args.steps
    val = total / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "fields/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "1920x1080 VHS-SP (-vhs -vhs-speed sp), synthetic perturbed colour bars",
                   "fields_per_step": cores * per_proc, "processes": cores},
        "cpu_baseline": {"value": val, "unit": "fields/s", "cores": cores, "kind": kind,
                         "sample": "%d processes x %d fields per step, %d steps, 1080p VHS-SP" % (cores, per_proc, args.steps)},
        "e2e": {"value": val, "unit": "fields/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------
def make_device_stream(torch, n, first, device):
    """Synthetic parity stream generated on the device: int32 BGRA, alpha 0."""
    x = torch.arange(W, device=device, dtype=torch.int64)[None, :]
    y = torch.arange(H, device=device, dtype=torch.int64)[:, None]
    bars = torch.tensor(BARS, device=device, dtype=torch.int64)
    out = torch.empty((n, H, W), device=device, dtype=torch.int32)
    for i in range(n):
        k = first + i
        xs = (x + 7 * k) % W
        base = bars[(xs * 8) // W].expand(H, W)
        pert = (((x * 2654435761) & 0xFFFFFFFF) ^ ((y * 40503) & 0xFFFFFFFF) ^ ((k * 97) & 0xFFFFFFFF)) >> 29
        pic = torch.zeros((H, W), device=device, dtype=torch.int64)
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


def run_own_arm(args):
    import numpy as np
    import torch
    import composite_video_simulator_b200 as cvs
    from composite_video_simulator_b200.params import CvsParams

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
        p0 = cvs.params_from_argv(ARGV)
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
    from composite_video_simulator_b200 import sharding

    src = make_device_stream(torch, B, rank * B, dev)
    dst = torch.zeros_like(src)
    torch.cuda.synchronize()

    def step(i):
        base, _ = sharding.chunk(i, rank, world, B)   # this rank's contiguous chunk of the global field stream
        eng.rng_seek(sharding.stream_position(params, W, H, base))
        eng.composite_fields_device(dst, src, B, H, W, base)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    eng.synchronize()
    barrier()
    eng.kernel_time_reset()
    launches0 = eng.kernel_launches()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(stream)
    e1.synchronize()
    eng.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.kernel_launches() - launches0
    kms, kn = eng.kernel_time_query()
    t = torch.tensor([ms, kms / max(kn, 1)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, kernel_ms = float(t[0]), float(t[1])
    value = world * B * args.steps / (ms_max / 1e3)

    # ---- side measurement (not the headline): the other per-pixel noise mode, same steps, same timing ----
    # (cvs_set_noise_mode: CVS_NOISE_FAST draws the per-pixel noise from counter generators, within +-1 LSB)
    other = None
    if args.noise_side:
        eng.set_noise_mode(args.noise != "fast")
        for i in range(args.warmup):
            step(i)
        eng.synchronize()
        barrier()
        eng.kernel_time_reset()
        e0.record(stream)
        for i in range(args.steps):
            step(args.warmup + i)
        e1.record(stream)
        e1.synchronize()
        eng.synchronize()
        barrier()
        kms2, kn2 = eng.kernel_time_query()
        t = torch.tensor([e0.elapsed_time(e1), kms2 / max(kn2, 1)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        other = {"noise": "exact" if args.noise == "fast" else "fast (counter generators, +-1 LSB)",
                 "value": world * B * args.steps / (float(t[0]) / 1e3), "unit": "fields/s", "kernel_ms_per_launch": float(t[1])}
        eng.set_noise_mode(args.noise == "fast")
        eng.kernel_time_reset()

    # ---- e2e: host buffers through the C ABI (H2D + kernels + D2H inside the timed region) ----
    Be = args.e2e_batch
    hsrc = [torch.empty((Be, H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
    hdst = [torch.zeros((Be, H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
    first = src[:Be].cpu() if Be <= B else make_device_stream(torch, Be, rank * Be, dev).cpu()
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
    barrier()
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
    hd_np = hd_np[(e2e_steps - 1) % 2]
    e2e_checksum = int(hd_np[0, 1::2].sum() & 0xFFFFFFFF)    # the step's result is read on the host
    clocks = sampler.stop() if rank == 0 else None           # sampled over both timed regions

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_fields) ----
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    alg_bytes = 8.0 * W * nl * B                               # 4 B read + 4 B written per processed pixel
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    traffic = None
    try:
        if args.preset != "sp" or (W, H) != (1920, 1080):
            raise KeyError("the ncu traffic capture is of the headline workload only")
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)        # measured by one `ncu --set full` capture, scaled to this launch size
            traffic = tj["k_fields_sp_dram_bytes_per_launch_64_fields"] / tj["fields_per_launch_in_capture"] * B
    except Exception:
        pass

    # ---- CPU baseline: the reference's own code, one thread, bounded sample ----
    cpu = None
    if world == 1 and args.cpu_fields > 0:
        dtc, kind = _cpu_run_fields(args.cpu_fields)
        cpu = {"value": args.cpu_fields / dtc, "unit": "fields/s", "cores": 1, "kind": kind,
               "sample": "%d consecutive 1080p VHS-SP fields of the synthetic stream, 1 thread, %.1f s" % (args.cpu_fields, dtc)}

    line = {
        "metric": METRIC, "value": value, "unit": "fields/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%dx%d %s (%s), %d fields per GPU per step, synthetic perturbed colour bars "
                               "generated on device" % (W, H, "VHS-" + args.preset.upper() if args.preset != "comp" else "composite only",
                                                        " ".join(ARGV) or "no switches", B),
                   "fields_per_step_per_gpu": B, "full_frames_per_s": value / 2,
                   "l2": "inputs larger than L2 (%.0f MB read + %.0f MB written per step)" % (alg_bytes / 2e6, alg_bytes / 2e6),
                   "noise": ("exact glibc rand() replay" if args.noise == "exact" else
                             "fast mode: per-pixel noise from counter generators (+-1 LSB), per-line draws exact"),
                   "other_noise_mode": other, "host_affinity": numa or "unchanged", "parallelism": "fields sharded by contiguous chunk, no data-path collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "fields/s", "h2d_bytes_per_step": Be * nl * 4 * W,
                "d2h_bytes_per_step": Be * nl * 4 * W, "fields_per_step_per_gpu": Be, "result_checksum": e2e_checksum},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "k_fields<float, %s, tv>" % {"sp": "VHS, 9", "lp": "VHS, 12", "ep": "VHS, 14", "comp": "composite"}[args.preset], "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_per_launch": kernel_ms},
        "cpu_baseline": cpu,
    }
    emit(line)
    if dist is not None:
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
    ap.add_argument("--batch", type=int, default=320, help="upper bound of fields per GPU per step (device-resident)")
    ap.add_argument("--no-wave-align", dest="wave_align", action="store_false",
                    help="use --batch as is instead of the wave-aligned batch cvs_preferred_batch() suggests")
    ap.add_argument("--e2e-batch", type=int, default=128, help="fields per GPU per step (host buffers; 4 pinned buffers of this many pictures per rank)")
    ap.add_argument("--cpu-fields", type=int, default=64, help="fields of the single-thread CPU baseline sample")
    ap.add_argument("--ref-fields-per-proc", type=int, default=4)
    ap.add_argument("--no-numa-bind", dest="numa_bind", action="store_false",
                    help="do not move the process to the CPUs of the GPU's NUMA node")
    ap.add_argument("--noise", default="exact", choices=["exact", "fast"],
                    help="per-pixel noise source of the measured runs (cvs_set_noise_mode); the headline is exact")
    ap.add_argument("--no-noise-side", dest="noise_side", action="store_false",
                    help="skip the side measurement of the other noise mode")
    ap.add_argument("--width", type=int, default=1920, help="experiments only: the headline metric is 1920x1080")
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--preset", default="sp", choices=["sp", "ep", "lp", "comp"],
                    help="sp = the headline VHS-SP workload; the others are for kernel experiments")
    args = ap.parse_args()
    global ARGV, W, H
    W, H = args.width, args.height
    ARGV = {"sp": ["-vhs", "-vhs-speed", "sp"], "ep": ["-vhs", "-vhs-speed", "ep"], "lp": ["-vhs", "-vhs-speed", "lp"],
            "comp": []}[args.preset]
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
