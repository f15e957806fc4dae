"""Throughput of the 4:2:2 path (k_yuv422) on device-resident pictures: fields/s and algorithmic GB/s.
Secondary bench (the headline metric is bench.py's BGRA path); prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from composite_video_simulator_b200 import yuv422  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--batch", type=int, default=339)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--argv", default="-vhs -vhs-speed sp")
    ap.add_argument("--cpu-fields", type=int, default=8)
    a = ap.parse_args()
    w, h, n = a.width, a.height, a.batch
    import helpers
    ly, lc = w + 32, w // 2 + 16
    Y0, U0, V0 = helpers.yuv422_frame(w, h, 0, 32)
    base = [torch.from_numpy(np.ascontiguousarray(np.broadcast_to(p, (n,) + p.shape))).cuda() for p in (Y0, U0[:, :lc], V0[:, :lc])]
    base[1] = base[1].contiguous(); base[2] = base[2].contiguous()
    work = [b.clone() for b in base]
    with yuv422.Yuv422Engine(a.argv.split(), max_w=w, max_h=h, max_batch=n) as eng:
        st = torch.cuda.Stream()
        eng.set_stream(st.cuda_stream)
        fno = 0
        with torch.cuda.stream(st):
            for _ in range(a.warmup):
                eng.process_fields_device(work[0], work[1], work[2], w, fno); fno += n
            eng.synchronize()
            eng.kernel_time_reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(a.steps):
                eng.process_fields_device(work[0], work[1], work[2], w, fno); fno += n
            e1.record(st)
            eng.synchronize()
        ms = e0.elapsed_time(e1)
        kms, kl = eng.kernel_time_query()
        eng.set_stream(0)
    # end to end through the host-pointer entry point: pinned host pictures, H2D + D2H inside the timed region
    ne = min(n, 64)
    hp = [torch.empty((ne,) + tuple(b.shape[1:]), dtype=torch.uint8).pin_memory() for b in base]
    for t, b in zip(hp, base):
        t.copy_(b[:ne].cpu())
    hn = [t.numpy() for t in hp]
    import time
    with yuv422.Yuv422Engine(a.argv.split(), max_w=w, max_h=h, max_batch=ne) as eng:
        eng.process_fields_host(hn[0], hn[1], hn[2], w, 0)
        t0 = time.perf_counter()
        for i in range(3):
            eng.process_fields_host(hn[0], hn[1], hn[2], w, ne * (i + 1))
        e2e = 3 * ne / (time.perf_counter() - t0)
    # the reference's own code on one host core (oracle/_ref/libref422.so), a few fields
    cpu = None
    ref = helpers.load_ref422()
    if ref is not None and a.cpu_fields > 0:
        p = helpers.params422(*a.argv.split())
        t0 = time.perf_counter()
        cached = helpers.yuv422_frame(w, h, 0, 32)
        t0 = time.perf_counter()
        helpers.run_ref422(ref, p, w, h, a.cpu_fields, frames=lambda k: tuple(x.copy() for x in cached))
        cpu = a.cpu_fields / (time.perf_counter() - t0)
    fields = n * a.steps
    bytes_per_field = 2 * 2 * w * ((h + 1) // 2)          # Y + U + V = 2 B per pixel, read + written
    print(json.dumps({"path": "yuv422", "metric": "fields_per_s", "value": fields / (ms / 1e3), "ms_per_step": ms / a.steps,
                      "kernel_ms_per_step": kms / max(kl, 1), "algorithmic_GBps": bytes_per_field * n / (kms / max(kl, 1) / 1e3) / 1e9,
                      "e2e_fields_per_s": e2e, "cpu_reference_fields_per_s_1thread": cpu,
                      "config": {"workload": "%dx%d %s" % (w, h, a.argv), "batch": n}}))


if __name__ == "__main__":
    main()
