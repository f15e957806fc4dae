"""The C-ABI library loads, exports every symbol include/cvs_ntsc.h declares, and fails loudly
(CVS_ERR_CUDA, no CPU fallback) when no CUDA device is usable."""
import ctypes as C
import os
import re

import pytest

import composite_video_simulator_b200 as cvs
from composite_video_simulator_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "cvs_ntsc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cvs_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTED_SYMBOLS) == names


def test_struct_layout_matches_header():
    p = cvs.default_params()
    assert p.struct_size == C.sizeof(cvs.CvsParams)
    assert _lib.load().cvs_abi_version() == 1


def test_strerror():
    lib = _lib.load()
    for st in range(0, -9, -1):
        assert lib.cvs_strerror(st)


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="this check is for machines without a CUDA device")
def test_no_cpu_fallback():
    with pytest.raises(cvs.CvsError) as e:
        cvs.Engine(["-vhs"])
    assert e.value.status == -3


def test_product_does_not_link_the_oracle():
    out = os.popen("ldd %s" % _lib.LIB_PATH).read()
    assert "oracle" not in out and "libref" not in out and "libemu" not in out
    syms = os.popen("nm -D --defined-only %s" % _lib.LIB_PATH).read()
    assert "oracle_" not in syms and "emu_" not in syms


# ---- include/cvs_yuv422.h -------------------------------------------------------------------------

def declared_functions_yuv422():
    src = open(os.path.join(ROOT, "include", "cvs_yuv422.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cvs422_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_yuv422_symbol_is_exported():
    from composite_video_simulator_b200 import yuv422
    lib = C.CDLL(_lib.LIB_PATH)
    names = declared_functions_yuv422()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(yuv422.EXPORTED_SYMBOLS) == names
    yuv422.lib()                                   # every signature binds


def test_yuv422_struct_layout_matches_header():
    from composite_video_simulator_b200 import yuv422
    # 24 int32 fields + 6 doubles, no implicit padding (reserved0 keeps the doubles 8-byte aligned)
    assert C.sizeof(yuv422.Yuv422Params) == 24 * 4 + 6 * 8
    p = yuv422.params_default()
    assert (p.output_width, p.output_height, p.video_noise, p.subcarrier_amplitude) == (720, 480, 2, 50)
    assert abs(p.vhs_out_sharpen_chroma - 0.85) < 1e-15 and p.vhs_out_sharpen == 1.5


@pytest.mark.skipif(not _no_gpu(), reason="this check is for machines without a CUDA device")
def test_yuv422_no_cpu_fallback():
    from composite_video_simulator_b200 import yuv422
    with pytest.raises(yuv422.Yuv422Error) as e:
        yuv422.Yuv422Engine(["-vhs"])
    assert e.value.status == -3
