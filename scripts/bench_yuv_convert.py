"""Bandwidth of cvs_bgra_to_yuv_device on device-resident pictures: frames/s and GB/s
(algorithmic bytes: 4 B/px read + 1.5 B/px (4:2:0) or 2 B/px (4:2:2) written).  Prints one JSON line per format."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import composite_video_simulator_b200 as cvs  # noqa: E402

w, h, n, steps = 1920, 1080, 128, 20
src = torch.randint(0, 1 << 24, (n, h, w), dtype=torch.int32, device="cuda")
peak = None
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
with cvs.Engine([], max_w=w, max_h=h, max_batch=1) as eng:
    st = torch.cuda.Stream()
    eng.set_stream(st.cuda_stream)
    for v420 in (True, False):
        ch = h // 2 if v420 else h
        y = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
        u = torch.empty((n, ch, w // 2), dtype=torch.uint8, device="cuda")
        v = torch.empty_like(u)
        with torch.cuda.stream(st):
            for _ in range(3):
                eng.bgra_to_yuv_device(y, u, v, src, w, h, n, fmt420=v420)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(steps):
                eng.bgra_to_yuv_device(y, u, v, src, w, h, n, fmt420=v420)
            e1.record(st)
            e1.synchronize()
        ms = e0.elapsed_time(e1) / steps
        bytes_ = n * w * h * (4 + (1.5 if v420 else 2.0))
        gbs = bytes_ / (ms / 1e3) / 1e9
        print(json.dumps({"kernel": "k_bgra_to_yuv", "format": "yuv420p" if v420 else "yuv422p", "frames_per_s": n / (ms / 1e3),
                          "ms_per_launch": ms, "pictures_per_launch": n, "algorithmic_GBps": gbs,
                          "frac_of_measured_hbm_peak": (gbs / peak) if peak else None}))

# the input side: cvs_scale_to_bgra_device (frame_copy_scale), NV12 720x480 -> BGRA 1920x1080 and 1080p -> 1080p
with cvs.Engine([], max_w=w, max_h=h, max_batch=1) as eng:
    st = torch.cuda.Stream()
    eng.set_stream(st.cuda_stream)
    for sw, sh in ((720, 480), (1920, 1080), (3840, 2160)):
        ns = 64
        Y = torch.randint(0, 256, (ns, sh, sw), dtype=torch.uint8, device="cuda")
        UV = torch.randint(0, 256, (ns, sh // 2, sw), dtype=torch.uint8, device="cuda")
        dst = torch.empty((ns, h, w), dtype=torch.int32, device="cuda")
        with torch.cuda.stream(st):
            for _ in range(3):
                eng.scale_to_bgra_device(dst, w, h, [Y, UV], [sw, sw], sw, sh, 3, n=ns)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(steps):
                eng.scale_to_bgra_device(dst, w, h, [Y, UV], [sw, sw], sw, sh, 3, n=ns)
            e1.record(st)
            e1.synchronize()
        ms = e0.elapsed_time(e1) / steps
        bytes_ = ns * (sw * sh * 1.5 + w * h * 4.0)
        gbs = bytes_ / (ms / 1e3) / 1e9
        print(json.dumps({"kernel": "k_sws_yuv_to_bgra", "format": "nv12 %dx%d -> bgra %dx%d" % (sw, sh, w, h),
                          "frames_per_s": ns / (ms / 1e3), "algorithmic_GBps": gbs, "frac_of_measured_hbm_peak": gbs / peak if peak else None}))
