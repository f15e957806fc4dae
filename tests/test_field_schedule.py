"""tools/field_schedule.h -- which decoded picture every output field shows, the bookkeeping of the reference's main loop
(ffmpeg_ntsc.cpp:2146-2283) that tools/cvs_ffmpeg_ntsc.cpp batches for the GPU -- driven on the CPU against a direct
restatement of the reference's rule (tests/field_schedule_harness.cpp): 29.97p, 59.94p, 3:2 pulldown, late starts (black
fields first), missing pts, gaps and 200 random schedules, for batches of 1 .. 64 fields."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_field_schedule_follows_the_reference_rule(tmp_path):
    exe = str(tmp_path / "field_schedule_harness")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-O1", "-o", exe, os.path.join(ROOT, "tests", "field_schedule_harness.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
