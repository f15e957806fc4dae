"""Shared test helpers: checker libraries, synthetic streams, hashing, comparison.

Only tests (and bench.py / __graft_entry__.smoke) may touch oracle/; the product never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import composite_video_simulator_b200 as cvs  # noqa: E402
from composite_video_simulator_b200 import build as cvs_build  # noqa: E402
from composite_video_simulator_b200.params import CvsParams  # noqa: E402

ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class OracleRng(C.Structure):
    _fields_ = [("r", C.c_uint32 * 31), ("f", C.c_int), ("b", C.c_int), ("pos", C.c_ulonglong)]


def load_oracle():
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "ntsc_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "liboracle.so"], cwd=ORACLE_DIR, stdout=subprocess.DEVNULL)
    lib = C.CDLL(path)
    lib.oracle_draws_per_field.restype = C.c_ulonglong
    lib.oracle_fnv1a64.restype = C.c_ulonglong
    lib.oracle_fnv1a64.argtypes = [C.c_void_p, C.c_ulonglong]
    lib.oracle_rng_next.restype = C.c_uint32
    return lib


def load_ref():
    """The reference's own composite_layer(), extracted at build time (never committed)."""
    path = os.path.join(ORACLE_DIR, "_ref", "libref.so")
    if not os.path.exists(path):
        if os.path.exists("/root/reference/ffmpeg_ntsc.cpp"):
            subprocess.check_call(["make", "ref"], cwd=ORACLE_DIR, stdout=subprocess.DEVNULL)
        else:
            return None
    lib = C.CDLL(path)
    return lib


def load_refaudio():
    """The reference's own composite_audio_process(), extracted at build time (never committed)."""
    path = os.path.join(ORACLE_DIR, "_ref", "librefaudio.so")
    if not os.path.exists(path):
        if os.path.exists("/root/reference/ffmpeg_ntsc.cpp"):
            subprocess.check_call(["make", "ref"], cwd=ORACLE_DIR, stdout=subprocess.DEVNULL)
        else:
            return None
    return C.CDLL(path)


def audio_signal(n, channels, seed):
    """Deterministic test PCM [n, channels] int16: a sweep, a loud low tone (drives the limiter) and LCG noise."""
    t = np.arange(n, dtype=np.float64) / 44100.0
    x = 0.45 * np.sin(2 * np.pi * (200.0 + 6000.0 * t) * t) + 0.7 * np.sin(2 * np.pi * 55.0 * t)
    s = np.uint64(seed * 2654435761 + 12345)
    noise = np.empty(n * channels, dtype=np.float64)
    v = int(s) & 0xFFFFFFFF
    for i in range(n * channels):
        v = (v * 1664525 + 1013904223) & 0xFFFFFFFF
        noise[i] = ((v >> 8) / float(1 << 24) - 0.5) * 0.3
    out = np.empty((n, channels), dtype=np.float64)
    for c in range(channels):
        out[:, c] = x * (1.0 if c == 0 else 0.8) + noise[c::channels]
    return np.clip(np.round(out * 32768.0), -32768, 32767).astype(np.int16)


def run_refaudio(lib, p, pcm_packets):
    """The reference on a list of int16 packets [n, ch] (fresh filters, default-seeded rand()); returns the
    processed packets and the number of rand() draws it made."""
    lib.refaudio_setup(C.byref(p))
    lib.refaudio_srand(1)
    out = []
    for pk in pcm_packets:
        a = np.ascontiguousarray(pk.copy())
        lib.refaudio_process(a.ctypes.data_as(C.c_void_p), C.c_uint(a.shape[0]))
        out.append(a)
    return out, int(lib.refaudio_rand())


def load_emu():
    """CPU emulation of the lane pipeline.  CVS_EMU_KT=4 builds it with the 4-pixel step (experiments)."""
    kt = os.environ.get("CVS_EMU_KT")
    lib = C.CDLL(cvs_build.build_emu(("-DCVS_KT=%s" % kt,) if kt else ()))
    return lib


def params(*argv):
    return cvs.params_from_argv(list(argv))


BARS = [0xC0C0C0, 0xC0C000, 0x00C0C0, 0x00C000, 0xC000C0, 0xC00000, 0x0000C0, 0x000000]   # 75% bars


def bars_frame(w, h):
    x = np.arange(w)
    row = np.array(BARS, dtype=np.uint32)[(x * 8) // w]
    return np.ascontiguousarray(np.broadcast_to(row, (h, w))).copy()


def stream_frame(w, h, k):
    """Parity stream of SURVEY.md 8(d): bars rotated by 7k px plus a per-pixel hash perturbation."""
    x = np.arange(w, dtype=np.uint32)[None, :]
    y = np.arange(h, dtype=np.uint32)[:, None]
    xs = (x + np.uint32(7 * k)) % np.uint32(w)
    base = np.broadcast_to(np.array(BARS, dtype=np.uint32)[(xs * np.uint32(8)) // np.uint32(w)], (h, w))
    pert = ((x * np.uint32(2654435761)) ^ (y * np.uint32(40503)) ^ np.uint32((k * 97) & 0xFFFFFFFF)) >> np.uint32(29)
    out = np.zeros((h, w), dtype=np.uint32)
    for sh in (0, 8, 16):
        c = ((base >> np.uint32(sh)) & np.uint32(0xFF)) + pert
        out |= np.minimum(c, 255).astype(np.uint32) << np.uint32(sh)
    return out


def noise_frame(w, h, seed):
    rng = np.random.RandomState(seed)
    return (rng.randint(0, 1 << 24, size=(h, w), dtype=np.int64)).astype(np.uint32)


def fnv1a64(arr):
    """FNV-1a-64 of the bytes of arr (the known-answer hash of SURVEY.md App. D)."""
    a = np.ascontiguousarray(arr)
    return int(load_oracle().oracle_fnv1a64(a.ctypes.data, a.nbytes))


def load_golden():
    """name -> (argv, w, h, n, dst) fixtures produced by the reference's own code (make_golden.py)."""
    out = {}
    for fn in sorted(os.listdir(GOLDEN_DIR)):
        if fn.endswith(".npz") and not fn.startswith(("yuv422_", "audio_", "swscale_")):   # (those have their own loaders)
            z = np.load(os.path.join(GOLDEN_DIR, fn))
            out[fn[:-4]] = ([str(a) for a in z["argv"]], int(z["w"]), int(z["h"]), int(z["n"]), z["dst"])
    return out


def run_oracle(lib, p, frames, n, w, h, dst=None, g=None, interlaced=0, tff=0, field_fn=None):
    """n sequential composite_layer() calls on the oracle; frames(k) -> uint32[h,w]."""
    if dst is None:
        dst = np.zeros((h, w), dtype=np.uint32)
    if g is None:
        g = OracleRng()
        lib.oracle_rng_seed(C.byref(g), 1)
    for k in range(n):
        src = frames(k)
        field = field_fn(k) if field_fn else (k & 1) ^ 1
        rc = lib.oracle_composite_layer(C.byref(p), C.byref(g), dst.ctypes.data_as(C.c_void_p), dst.strides[0],
                                        src.ctypes.data_as(C.c_void_p), src.strides[0], w, h, interlaced, tff,
                                        field, C.c_ulonglong(k))
        assert rc == 0
    return dst, g


def run_ref(lib, p, frames, n, w, h, interlaced=0, tff=0):
    dst = np.zeros((h, w), dtype=np.uint32)
    lib.ref_set_params(C.byref(p))
    lib.ref_srand(1)
    for k in range(n):
        src = frames(k)
        lib.ref_composite_layer(dst.ctypes.data_as(C.c_void_p), 4 * w, src.ctypes.data_as(C.c_void_p), 4 * w,
                                w, h, interlaced, tff, (k & 1) ^ 1, C.c_ulonglong(k))
    return dst


def run_emu(lib, p, frames, n, w, h, precision, general=0, interlaced=0, tff=0, noise_fast=0):
    dst = np.zeros((h, w), dtype=np.uint32)
    pos = C.c_ulonglong(0)
    for k in range(n):
        src = frames(k)
        rc = lib.emu_composite_layer_ex(C.byref(p), precision, C.byref(pos), dst.ctypes.data_as(C.c_void_p), 4 * w,
                                        src.ctypes.data_as(C.c_void_p), 4 * w, w, h, interlaced, tff, (k & 1) ^ 1,
                                        C.c_ulonglong(k), general, noise_fast)
        assert rc == 0, rc
    return dst, pos.value


def channel_diff(a, b):
    """(max |delta|, #values differing, #values differing by > 1) over the 8-bit channels."""
    d = np.abs(a.view(np.uint8).astype(np.int16) - b.view(np.uint8).astype(np.int16))
    return int(d.max()) if d.size else 0, int((d > 0).sum()), int((d > 1).sum())


# ---- the 4:2:2 sibling path (include/cvs_yuv422.h) ----------------------------------------------

def load_oracle422():
    path = os.path.join(ORACLE_DIR, "libyuv422oracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("yuv422_oracle.c", "yuv422_oracle.h", "ntsc_oracle.c")]
    if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "libyuv422oracle.so"], cwd=ORACLE_DIR, stdout=subprocess.DEVNULL)
    lib = C.CDLL(path)
    lib.oracle422_draws_per_field.restype = C.c_ulonglong
    return lib


def load_ref422():
    """The reference's own composite_video_process() / render_field(), extracted at build time."""
    path = os.path.join(ORACLE_DIR, "_ref", "libref422.so")
    if not os.path.exists(path):
        if os.path.exists("/root/reference/ffmpeg_to_composite.cpp"):
            subprocess.check_call(["make", "ref"], cwd=ORACLE_DIR, stdout=subprocess.DEVNULL)
        else:
            return None
    return C.CDLL(path)


def params422(*argv):
    from composite_video_simulator_b200 import yuv422
    return yuv422.params_from_argv(list(argv))


def yuv422_frame(w, h, k, pad=0):
    """Synthetic 4:2:2 picture k: (Y[h, w+pad], U[h, w/2+pad], V[h, w/2+pad]) uint8.  Luma = bars rotated
    by 7k px with a hash perturbation, chroma = saturated bars + perturbation; the pad columns hold a
    recognisable pattern (the reference reads two luma bytes past each row)."""
    x = np.arange(w + pad, dtype=np.uint32)[None, :]
    y = np.arange(h, dtype=np.uint32)[:, None]
    xs = (x + np.uint32(7 * k)) % np.uint32(max(w, 1))
    bar = (xs * np.uint32(8)) // np.uint32(max(w, 1))
    ylev = np.array([180, 162, 131, 112, 84, 65, 35, 16], dtype=np.uint32)
    ulev = np.array([128, 44, 156, 72, 184, 100, 212, 128], dtype=np.uint32)
    vlev = np.array([128, 142, 44, 58, 198, 212, 114, 128], dtype=np.uint32)
    h1 = ((x * np.uint32(2654435761)) ^ (y * np.uint32(40503)) ^ np.uint32((k * 97) & 0xFFFFFFFF)) >> np.uint32(28)
    Y = np.clip(ylev[bar] + h1, 0, 255).astype(np.uint8)
    cw = w // 2
    xc = np.arange(cw + pad, dtype=np.uint32)[None, :]
    cbar = (((2 * xc + np.uint32(7 * k)) % np.uint32(max(w, 1))) * np.uint32(8)) // np.uint32(max(w, 1))
    h2 = ((xc * np.uint32(2246822519)) ^ (y * np.uint32(3266489917)) ^ np.uint32((k * 131) & 0xFFFFFFFF)) >> np.uint32(29)
    U = np.clip(ulev[cbar] + h2, 0, 255).astype(np.uint8)
    V = np.clip(vlev[cbar] + (h2 ^ np.uint32(5)), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(Y), np.ascontiguousarray(U), np.ascontiguousarray(V)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def run_ref422(ref, p, w, h, n, pad=2, first=0, seed=1, frames=None):
    """n fields through the reference's composite_video_process(); each field works on a fresh picture
    (frame k); returns the list of (Y, U, V) after processing."""
    ref.ref422_set_params(C.byref(p))
    ref.ref422_srand(seed)
    out = []
    for k in range(first, first + n):
        Y, U, V = (frames(k) if frames else yuv422_frame(w, h, k, pad))
        # slack after the last row: the reference reads 2 bytes past it
        Yb = np.zeros(Y.size + 16, dtype=np.uint8)
        Yb[:Y.size] = Y.ravel()
        ref.ref422_composite_video_process(_ptr(Yb), C.c_int(Y.shape[1]), _ptr(U), C.c_int(U.shape[1]),
                                           _ptr(V), C.c_int(V.shape[1]), C.c_int(w), C.c_int(h),
                                           C.c_uint((k & 1) ^ 1), C.c_ulonglong(k))
        out.append((Yb[:Y.size].reshape(Y.shape).copy(), U, V))
    return out


def run_oracle422(orc, p, w, h, n, pad=2, first=0, frames=None, rng=None):
    g = rng
    if g is None:
        g = OracleRng()
        orc.oracle_rng_seed(C.byref(g), 1)
    out = []
    for k in range(first, first + n):
        Y, U, V = (frames(k) if frames else yuv422_frame(w, h, k, pad))
        rc = orc.oracle422_composite_video_process(C.byref(p), C.byref(g), _ptr(Y), C.c_int(Y.shape[1]),
                                                   _ptr(U), C.c_int(U.shape[1]), _ptr(V), C.c_int(V.shape[1]),
                                                   C.c_int(w), C.c_int(h), C.c_uint((k & 1) ^ 1), C.c_ulonglong(k))
        assert rc == 0
        out.append((Y, U, V))
    return out, g


def load_emu422():
    return C.CDLL(cvs_build.build_emu422())


def run_emu422(emu, p, w, h, n, pad=2, first=0, force_general=0, frames=None, rng_pos=0):
    out = []
    pos = rng_pos
    for k in range(first, first + n):
        Y, U, V = (frames(k) if frames else yuv422_frame(w, h, k, pad))
        nd = C.c_ulonglong(0)
        rc = emu.emu422_process(C.byref(p), C.c_ulonglong(pos), _ptr(Y), C.c_int(Y.shape[1]), _ptr(U), C.c_int(U.shape[1]),
                                _ptr(V), C.c_int(V.shape[1]), C.c_int(w), C.c_int(h), C.c_uint((k & 1) ^ 1),
                                C.c_ulonglong(k), C.c_int(force_general), C.byref(nd))
        assert rc == 0, rc
        pos += nd.value
        out.append((Y, U, V))
    return out, pos


# ---- the reference's field loop with its audio step: one libc rand() stream for both ----------------

def audio_due(frames_done, current, ntsc=True):
    """tools/cvs_ntsc_raw.cpp's schedule: an audio packet that starts at or before field `current` is shown
    (current * 1001 / 60000 s, PAL current / 50 s) runs before that field."""
    num, den = (1001, 60000) if ntsc else (1, 50)
    return frames_done * den <= current * num * 44100


def reference_av_loop(ref, refaudio, oracle, p, frames, pcm, w, h, fields_per_frame, delay, packet, bob=True):
    """The reference's main loop (ffmpeg_ntsc.cpp:2140-2284) on its OWN code: libref.so's composite_layer() and
    librefaudio.so's composite_audio_process() both call libc rand(), i.e. they share one stream exactly as inside the
    reference program.  Returns (pictures [nfields, h, w], processed pcm)."""
    ref.ref_set_params(C.byref(p))
    refaudio.refaudio_setup(C.byref(p))
    ref.ref_srand(1)
    ch = refaudio.refaudio_channels()
    pcm = np.ascontiguousarray(pcm.copy())
    done, total = 0, pcm.shape[0]
    ring = [np.zeros((h, w), dtype=np.uint32) for _ in range(delay)]
    out, idx = [], 0

    def audio_packet():
        nonlocal done
        n = min(packet, total - done)
        blk = np.ascontiguousarray(pcm[done:done + n])
        refaudio.refaudio_process(blk.ctypes.data_as(C.c_void_p), C.c_uint(n))
        pcm[done:done + n] = blk
        done += n

    for current in range(len(frames) * fields_per_frame):
        while done < total and audio_due(done, current, bool(p.output_ntsc)):
            audio_packet()
        src = frames[current // fields_per_frame]
        f = (current & 1) ^ 1
        pic = ring[idx]
        ref.ref_composite_layer(pic.ctypes.data_as(C.c_void_p), 4 * w, src.ctypes.data_as(C.c_void_p), 4 * w,
                                w, h, 0, 0, f, C.c_ulonglong(current))
        if bob:
            oracle.oracle_bob(pic.ctypes.data_as(C.c_void_p), 4 * w, w, h, f)
        out.append(pic.copy())
        idx = (idx + 1) % delay
    while done < total:
        audio_packet()
    assert ch == pcm.shape[1]
    return np.stack(out), pcm


# ---- the picture conversions either side of the path (oracle/convert_oracle.c; unpinned: libswscale is absent) ----

def load_convert_oracle():
    path = os.path.join(ORACLE_DIR, "libconvertoracle.so")
    src = os.path.join(ORACLE_DIR, "convert_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "libconvertoracle.so"], cwd=ORACLE_DIR, stdout=subprocess.DEVNULL)
    return C.CDLL(path)


def oracle_bgra_to_yuv(bgra, v420):
    """bgra: uint32[h, w] -> (y, u, v) uint8 planes."""
    lib = load_convert_oracle()
    h, w = bgra.shape
    cw, ch = (w + 1) // 2, ((h + 1) // 2 if v420 else h)
    src = np.ascontiguousarray(bgra)
    y = np.zeros((h, w), np.uint8)
    u = np.zeros((ch, cw), np.uint8)
    v = np.zeros((ch, cw), np.uint8)
    rc = lib.oracle_bgra_to_yuv(_ptr(y), w, _ptr(u), cw, _ptr(v), cw, _ptr(src), src.strides[0], w, h, 1 if v420 else 0)
    assert rc == 0
    return y, u, v


def oracle_sws_bgra_route(img, sw, sh, dw, dh):
    """The library's own route for a BGRA source at another size, for every geometry (oracle_sws_bgra_to_bgra)."""
    lib = load_convert_oracle()
    src = np.ascontiguousarray(img)
    dst = np.zeros((dh, dw), np.uint32)
    rc = lib.oracle_sws_bgra_to_bgra(_ptr(dst), 4 * dw, dw, dh, _ptr(src), src.strides[0], sw, sh)
    assert rc == 0, rc
    return dst


def oracle_sws_filter(srcn, dstn, one, max_taps=16):
    """The bilinear filter bank of one axis as libswscale builds it (oracle/convert_oracle.c) -> (pos[dstn], coef[dstn, taps])."""
    lib = load_convert_oracle()
    pos = np.zeros(dstn, np.int32)
    coef = np.zeros(dstn * max_taps, np.int32)
    n = lib.oracle_sws_bilinear_filter(srcn, dstn, one, _ptr(pos), _ptr(coef), max_taps)
    assert n > 0, n
    return pos, coef[:dstn * n].reshape(dstn, n).copy()


def oracle_scale_to_bgra(planes, sw, sh, fmt, dw, dh):
    """planes: list of uint8 arrays (BGRA: one [sh, sw*4]); -> uint32[dh, dw]."""
    lib = load_convert_oracle()
    ps = [np.ascontiguousarray(p) for p in planes] + [None] * (3 - len(planes))
    dst = np.zeros((dh, dw), np.uint32)
    rc = lib.oracle_scale_to_bgra(_ptr(dst), 4 * dw, dw, dh, *[(_ptr(p) if p is not None else None) for p in ps],
                                  *[(p.strides[0] if p is not None else 0) for p in ps], sw, sh, fmt)
    assert rc == 0, rc
    return dst
