"""In-tree build of the product library and of the parity checkers.

``build_product()``  nvcc -> composite_video_simulator_b200/libcvs_ntsc.so   (sm_100a)
``build_oracle()``   gcc  -> oracle/liboracle.so, and oracle/_ref/libref.so when the reference
                     tree is present (test infrastructure only)
``build_emu()``      g++  -> tests/_build/libemu.so (CPU emulation of the lane pipeline; tests only)
``build_emu422()``   g++  -> tests/_build/libemu422.so (the same for the 4:2:2 pipeline)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "composite_video_simulator_b200", "csrc")


def _run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: %s (in %s)" % (" ".join(cmd), cwd))
    return r.stdout


def build_product(jobs=8):
    return _run(["make", "-j%d" % jobs], CSRC)


def build_oracle():
    return _run(["make"], os.path.join(ROOT, "oracle"))


def build_emu(extra=()):
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libemu%s.so" % "".join(e.replace("-D", "_").replace("=", "") for e in extra))
    srcs = [os.path.join(ROOT, "tests", "emu_harness.cpp")] + [
        os.path.join(CSRC, f) for f in ("field_plan.cpp", "glibc_rand.cpp", "cvs_params.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("lane_pipeline.cuh", "field_plan.h", "glibc_rand.h")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    _run(["g++", "-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared"] + list(extra) + ["-o", out] + srcs, ROOT)
    return out


def build_emu422():
    """CPU emulation of the 4:2:2 lane pipeline (tests only)."""
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libemu422.so")
    srcs = [os.path.join(ROOT, "tests", "emu422_harness.cpp")] + [
        os.path.join(CSRC, f) for f in ("yuv422_plan.cpp", "field_plan.cpp", "glibc_rand.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("yuv422_pipeline.cuh", "yuv422_plan.h", "lane_pipeline.cuh",
                                                   "field_plan.h", "glibc_rand.h")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    _run(["g++", "-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", out] + srcs, ROOT)
    return out


if __name__ == "__main__":
    build_product()
    build_oracle()
    print("built", os.path.join(ROOT, "composite_video_simulator_b200", "libcvs_ntsc.so"))
