"""tools/cvs_ffmpeg_ntsc.cpp (the FFmpeg-linked host, ffmpeg_ntsc.cpp main() :1923-2331) cannot be linked
here -- the image has no FFmpeg development headers -- so two things are checked: (1) its FFmpeg branch is valid C++
against declarations of the API subset it uses (tests/ffmpeg_decl/, written for this check) and against the real
include/cvs_ntsc.h; (2) the no-FFmpeg branch builds against the product library and reports what is missing."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "cvs_ffmpeg_ntsc.cpp")

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")


def test_ffmpeg_branch_is_valid_against_the_declared_api():
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-DCVS_WITH_FFMPEG",
                        "-I" + os.path.join(ROOT, "tests", "ffmpeg_decl"), SRC], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_stub_branch_builds_and_says_what_is_missing(tmp_path):
    lib = os.path.join(ROOT, "composite_video_simulator_b200")
    if not os.path.exists(os.path.join(lib, "libcvs_ntsc.so")):
        pytest.skip("product library not built")
    exe = str(tmp_path / "cvs_ffmpeg_ntsc")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-o", exe, SRC, "-L" + lib, "-lcvs_ntsc", "-Wl,-rpath," + lib],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "-i", "a", "-o", "b"], capture_output=True, text=True)
    assert r.returncode == 2 and "FFmpeg" in r.stderr
