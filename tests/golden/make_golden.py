"""Generate the committed golden fixtures from the REFERENCE's own code.

Run in the build container (needs /root/reference, from which oracle/Makefile extracts
composite_layer() into the git-ignored oracle/_ref/):

    python tests/golden/make_golden.py

For each case the fixture stores the argv, geometry, number of fields and the final dst picture
(uint32[h, w]) after n sequential composite_layer() calls on helpers.stream_frame(w, h, k) with the
default-seeded rand() -- produced by oracle/_ref/libref.so, i.e. by the reference source itself, not by
anything written in this repository.  Small geometries keep the fixtures a few tens of KB.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402

CASES = {
    "comp_160x120": (160, 120, 4, []),
    "sp_160x120": (160, 120, 4, ["-vhs", "-vhs-speed", "sp"]),
    "lp_164x122": (164, 122, 3, ["-vhs", "-vhs-speed", "lp"]),
    "ep_160x120": (160, 120, 4, ["-vhs", "-vhs-speed", "ep"]),
    "ep_full_101x67": (101, 67, 3, ["-vhs", "-vhs-speed", "ep", "-out-composite-lowpass-lite", "0"]),
    "catv_sp_160x120": (160, 120, 2, ["-vhs", "-comp-catv2", "-subcarrier-amp", "45"]),
    "pal_svideo_160x144": (160, 144, 2, ["-tvstd", "pal", "-vhs", "-vhs-svideo", "1"]),
    "phase90_hs_160x480": (160, 480, 2, ["-vhs", "-comp-phase", "90", "-vhs-head-switching-phase", "0.002"]),
}


def main():
    ref = helpers.load_ref()
    assert ref is not None, "needs /root/reference"
    for name, (w, h, n, argv) in CASES.items():
        p = helpers.params(*argv)
        out = helpers.run_ref(ref, p, lambda k: helpers.stream_frame(w, h, k), n, w, h)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), argv=np.array(argv, dtype="U64"), w=w, h=h, n=n, dst=out)
        print(name, out.shape, "%016x" % helpers.load_oracle().oracle_fnv1a64(out.ctypes.data, out.nbytes))


if __name__ == "__main__":
    main()
