import sys, os
sys.path.insert(0, "tests")
import numpy as np, helpers
import composite_video_simulator_b200 as cvs
from composite_video_simulator_b200 import yuv422
w, h, n = 100, 67, 5
p = helpers.params("-vhs")
src = np.stack([helpers.noise_frame(w, h, 40 + k) for k in range(n)])
for fast in (False, True):
    out = np.zeros((n, h, w), dtype=np.uint32)
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=n) as eng:
        eng.set_noise_mode(fast)
        eng.composite_fields_host(out, src, 0)
        one = np.zeros((h, w), dtype=np.uint32)
        eng.composite_layer(one, src[0], 1, 0)
    print("bgra", fast, int(out.sum() & 0xffffffff))
w, h = 720, 480
p = helpers.params("-vhs", "-vhs-speed", "ep")
out = np.zeros((2, h, w), dtype=np.uint32)
src = np.stack([helpers.stream_frame(w, h, k) for k in range(2)])
with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=2) as eng:
    eng.composite_fields_host(out, src, 0)
print("bgra 480p", int(out.sum() & 0xffffffff))
w, h, n = 64, 67, 5
fr = [helpers.yuv422_frame(w, h, k, 4) for k in range(n)]
Y, U, V = [np.stack([f[i] for f in fr]) for i in range(3)]
with yuv422.Yuv422Engine(["-vhs"], max_w=w, max_h=h, max_batch=n) as eng:
    eng.process_fields_host(Y, U, V, w, 0)
print("yuv422", int(Y.sum()))
