"""Generates tests/golden/yuv422_*.npz from the reference's OWN composite_video_process()
(oracle/_ref/libref422.so, built from /root/reference/ffmpeg_to_composite.cpp by oracle/Makefile).
Run in the build container (the reference tree is not on the GPU box): python tests/golden/make_golden_yuv422.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402

CASES = [
    ("comp_160x60", 160, 60, 2, []),
    ("vhs_sp_168x61", 168, 61, 3, ["-vhs"]),
    ("vhs_ep_pal_phase_172x64", 172, 64, 2, ["-tvstd", "pal", "-vhs", "-vhs-speed", "ep", "-width", "172"]),
    ("catv_recomb_164x40", 164, 40, 2, ["-comp-catv2", "-yc-recomb", "1", "-comp-phase", "90"]),
]

if __name__ == "__main__":
    ref = helpers.load_ref422()
    assert ref is not None, "needs /root/reference"
    for name, w, h, n, argv in CASES:
        p = helpers.params422(*argv)
        out = helpers.run_ref422(ref, p, w, h, n)
        d = {"argv": " ".join(argv), "w": w, "h": h, "n": n}
        for k in range(n):
            d["Y%d" % k], d["U%d" % k], d["V%d" % k] = out[k]
        np.savez_compressed(os.path.join(HERE, "yuv422_%s.npz" % name), **d)
        print("wrote", name)
