"""GPU parity: the CUDA path, called through the C ABI, against the oracle on the same inputs.

fp32 (production): max |delta| <= 1 LSB per 8-bit channel (north_star's bar).
fp64 (validation mode): bit-exact with the reference arithmetic.
"""
import os

import numpy as np
import pytest

import helpers
import composite_video_simulator_b200 as cvs

pytestmark = pytest.mark.gpu

CASES = [
    (720, 480, 4, []),
    (720, 480, 4, ["-vhs", "-vhs-speed", "sp"]),
    (720, 480, 3, ["-vhs", "-vhs-speed", "ep"]),
    (724, 482, 3, ["-vhs", "-vhs-speed", "lp"]),
    (101, 67, 2, ["-vhs", "-vhs-speed", "lp", "-out-composite-lowpass-lite", "0"]),
    (720, 480, 2, ["-vhs", "-comp-catv", "-subcarrier-amp", "40"]),
    (720, 480, 2, ["-vhs", "-vhs-svideo", "1", "-comp-phase", "90"]),
    (720, 480, 2, ["-vhs", "-vhs-head-switching-phase", "0.001", "-vhs-head-switching-point", "0.5"]),
    (1920, 1080, 2, ["-vhs", "-vhs-speed", "sp"]),
    (1920, 1080, 2, ["-vhs", "-vhs-speed", "ep"]),          # BASELINE config 3
    (3840, 2160, 2, []),                                    # BASELINE config 4: composite only, 4K
    (3840, 2160, 1, ["-vhs", "-vhs-speed", "lp"]),          # head-switch shift beyond the in-kernel ring: pre-pass
    # every switch the oracle sweep (tests/test_oracle_vs_ref.py) pins against the reference's code, on the GPU:
    (720, 480, 3, ["-vhs", "-chroma-dropout", "30000"]),                       # a16: ~30 % of the rows lose chroma (:1891-1901)
    (720, 480, 4, ["-vhs", "-comp-phase", "270", "-comp-phase-offset", "1"]),  # line phase -(y>>1) (:1473-1480)
    (720, 480, 3, ["-vhs", "-comp-phase", "0"]),
    (720, 480, 3, ["-comp-phase", "90", "-comp-phase-offset", "3"]),
    (720, 480, 2, ["-vhs", "-nocolor-subcarrier"]),                            # :1715
    (720, 480, 2, ["-vhs", "-in-composite-lowpass", "0", "-out-composite-lowpass", "0"]),
    (720, 480, 3, ["-vhs", "-vhs-chroma-vblend", "0"]),                        # :1843
    (720, 480, 2, ["-vhs", "-vhs-svideo", "1", "-vhs-chroma-vblend", "0"]),
    (720, 480, 2, ["-comp-catv3", "-chroma-noise", "5"]),                      # chroma noise without VHS (:1718)
    (720, 576, 2, ["-tvstd", "pal", "-vhs"]),                                  # PAL: no vertical blend, 312.5-line head switch
    (33, 21, 3, ["-vhs", "-vhs-speed", "ep"]),                                 # a row shorter than the pipeline depth
    (1024, 512, 2, ["-vhs"]),                                                  # 2 MiB pictures: last row ends the mapping
]


@pytest.mark.parametrize("w,h,n,argv", CASES)
@pytest.mark.parametrize("use_double", [False, True])
def test_seam_matches_oracle(oracle, w, h, n, argv, use_double):
    p = helpers.params(*argv)
    frames = lambda k: helpers.stream_frame(w, h, k)
    want, g = helpers.run_oracle(oracle, p, frames, n, w, h)
    got = np.zeros((h, w), dtype=np.uint32)
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=4) as eng:
        eng.set_precision(use_double)
        for k in range(n):
            eng.composite_layer(got, frames(k), (k & 1) ^ 1, k)
        assert eng.rng_tell() == g.pos
    mx, nd, n2 = helpers.channel_diff(want, got)
    if use_double:
        assert mx == 0, (mx, nd)
    else:
        assert mx <= 1 and n2 == 0, (mx, nd, n2)
        assert nd <= 0.005 * want.size * 4


def test_batch_equals_sequential(oracle):
    w, h, n = 720, 480, 6
    p = helpers.params("-vhs")
    src = np.stack([helpers.stream_frame(w, h, k) for k in range(n)])
    seq = np.zeros((n, h, w), dtype=np.uint32)
    bat = np.zeros((n, h, w), dtype=np.uint32)
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=n) as eng:
        for k in range(n):
            eng.composite_layer(seq[k], src[k], (k & 1) ^ 1, k)
        pos = eng.rng_tell()
        eng.rng_seek(0)
        eng.composite_fields_host(bat, src, 0)
        assert eng.rng_tell() == pos
    assert np.array_equal(seq, bat)


@pytest.mark.parametrize("w,h,n,argv", [(100, 67, 7, ["-vhs"]),                 # 34 / 33 rows per parity: packed, two fields meet in a warp
                                        (64, 65, 5, ["-vhs", "-vhs-speed", "ep"]),  # 33 / 32 rows
                                        (64, 61, 5, ["-vhs"]),                   # 31 / 30 rows: too short to pack (per-field warps)
                                        (720, 480, 9, [])])
def test_packed_rows_batch_matches_oracle(oracle, w, h, n, argv):
    """A batch is cut into warps of 31 rows across field boundaries (scanline_kernels.cuh, `packed`): fp64 must
    stay bit-exact, whichever rows and fields meet in a warp, and equal to the per-field mapping."""
    p = helpers.params(*argv)
    frames = lambda k: helpers.noise_frame(w, h, 40 + k)
    src = np.stack([frames(k) for k in range(n)])
    got = {}
    for packed in ("1", "0"):
        os.environ["CVS_PACKED_ROWS"] = packed
        try:
            out = np.zeros((n, h, w), dtype=np.uint32)
            with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=n) as eng:
                eng.set_precision(True)
                eng.composite_fields_host(out, src, 0)
            got[packed] = out
        finally:
            os.environ.pop("CVS_PACKED_ROWS", None)
    assert np.array_equal(got["1"], got["0"])
    seq = np.zeros((n, h, w), dtype=np.uint32)
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=1) as eng:
        eng.set_precision(True)
        for k in range(n):
            eng.composite_layer(seq[k], src[k], (k & 1) ^ 1, k)
    assert np.array_equal(got["1"], seq)
    # and the sequential calls against the oracle (one destination picture, as the reference's loop uses it)
    acc = np.zeros((h, w), dtype=np.uint32)
    want, _ = helpers.run_oracle(oracle, p, frames, n, w, h)
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=1) as eng:
        eng.set_precision(True)
        for k in range(n):
            eng.composite_layer(acc, frames(k), (k & 1) ^ 1, k)
    assert np.array_equal(want, acc)


@pytest.mark.parametrize("name", sorted(helpers.load_golden().keys()))
def test_golden_fixtures(name):
    """Fixtures produced by the reference's own code (tests/golden/make_golden.py)."""
    argv, w, h, n, want = helpers.load_golden()[name]
    for use_double in (True, False):
        got = np.zeros((h, w), dtype=np.uint32)
        with cvs.Engine(argv=argv, max_w=w, max_h=h, max_batch=2) as eng:
            eng.set_precision(use_double)
            for k in range(n):
                eng.composite_layer(got, helpers.stream_frame(w, h, k), (k & 1) ^ 1, k)
        mx, nd, n2 = helpers.channel_diff(want, got)
        assert (mx == 0) if use_double else (mx <= 1 and n2 == 0), (use_double, mx, nd, n2)


def test_untouched_rows_and_alpha():
    w, h = 160, 120
    src = helpers.stream_frame(w, h, 0) | np.uint32(0xFF000000)       # alpha set in the source
    dst = np.full((h, w), 0xDEADBEEF, dtype=np.uint32)
    with cvs.Engine(["-vhs"], max_w=w, max_h=h, max_batch=1) as eng:
        eng.composite_layer(dst, src, 1, 0)
    assert (dst[0::2] == 0xDEADBEEF).all()                            # other parity untouched (:1910)
    assert (dst[1::2] >> 24 == 0).all()                               # alpha = 0 (:1914)


def test_invalid_geometry_leaves_dst_untouched():
    w, h = 64, 32
    src = helpers.stream_frame(w, h, 0)
    dst = np.full((h, w), 7, dtype=np.uint32)
    with cvs.Engine([], max_w=w, max_h=h, max_batch=1) as eng:
        with pytest.raises(cvs.CvsError) as e:
            eng.composite_layer(dst, src, 0, 0, dst_stride=4 * w - 4)   # stride < 4*w, :1580
        assert e.value.status == -1
        assert eng.rng_tell() == 0
        with pytest.raises(cvs.CvsError) as e:
            eng.composite_layer(np.zeros((h * 2, w), np.uint32), np.zeros((h * 2, w), np.uint32), 0, 0)
        assert e.value.status == -5                                   # larger than the context capacity
    assert (dst == 7).all()


def test_interlaced_source_rows(oracle):
    w, h, n = 160, 121, 3
    p = helpers.params("-vhs")
    frames = lambda k: helpers.noise_frame(w, h, k)
    want, _ = helpers.run_oracle(oracle, p, frames, n, w, h, interlaced=1, tff=1)
    got = np.zeros((h, w), dtype=np.uint32)
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=1) as eng:
        eng.set_precision(True)
        for k in range(n):
            eng.composite_layer(got, frames(k), (k & 1) ^ 1, k, src_interlaced=True, src_top_field_first=True)
    assert np.array_equal(want, got)


def test_device_pointers_and_batch_split():
    """Device-resident form with torch tensors on a torch stream; one batch == several smaller batches;
    unaligned row strides take the scalar load/store path and give the same pixels."""
    import torch
    w, h, n = 720, 480, 8
    p = helpers.params("-vhs", "-vhs-speed", "lp")
    src = torch.from_numpy(np.stack([helpers.stream_frame(w, h, k) for k in range(n)]).view(np.int32)).cuda()
    a = torch.zeros_like(src)
    b = torch.zeros_like(src)
    st = torch.cuda.Stream()
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=n) as eng:
        eng.set_stream(st.cuda_stream)
        eng.composite_fields_device(a, src, n, h, w, 0)
        pos = eng.rng_tell()
        eng.rng_seek(0)
        for k0 in range(0, n, 3):
            m = min(3, n - k0)
            eng.composite_fields_device(b[k0:], src[k0:], m, h, w, k0)
        assert eng.rng_tell() == pos
        eng.synchronize()
        assert torch.equal(a, b)
        # misaligned: pictures of width w-1 viewed inside the same buffers (stride 4*w, base + 4 bytes)
        eng.rng_seek(0)
        c = torch.zeros_like(src)
        eng.composite_fields_device(c.data_ptr() + 4, src.data_ptr() + 4, n, h, w - 1, 0, dst_pic_stride=4 * w * h,
                                    dst_stride=4 * w, src_pic_stride=4 * w * h, src_stride=4 * w)
        eng.synchronize()
    host = np.zeros((h, w - 1), dtype=np.uint32)
    srcn = src.cpu().numpy().view(np.uint32)
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=1) as eng:
        eng.composite_layer(host, np.ascontiguousarray(srcn[0][:, 1:]), 1, 0)
    assert np.array_equal(c.cpu().numpy().view(np.uint32)[0][1::2, 1:], host[1::2])


def test_1080p_vhs_sp_256_fields(oracle):
    """BASELINE config 2: 1920x1080 VHS-SP, +-1 LSB vs the CPU path on 256 synthetic frames.  The oracle
    runs the 256 calls serially into one reused picture; the GPU runs them as 4 batches."""
    import ctypes as C
    w, h, n, B = 1920, 1080, 256, 64
    p = helpers.params("-vhs", "-vhs-speed", "sp")
    g = helpers.OracleRng()
    oracle.oracle_rng_seed(C.byref(g), 1)
    want = np.zeros((h, w), dtype=np.uint32)
    hist = np.zeros(3, dtype=np.int64)
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=B) as eng:
        for k0 in range(0, n, B):
            src = np.stack([helpers.stream_frame(w, h, k) for k in range(k0, k0 + B)])
            got = np.zeros_like(src)
            eng.composite_fields_host(got, src, k0)
            for i in range(B):
                k = k0 + i
                f = (k & 1) ^ 1
                oracle.oracle_composite_layer(C.byref(p), C.byref(g), want.ctypes.data_as(C.c_void_p), 4 * w,
                                              src[i].ctypes.data_as(C.c_void_p), 4 * w, w, h, 0, 0, f, C.c_ulonglong(k))
                d = np.abs(want[f::2].view(np.uint8).astype(np.int16) - got[i][f::2].view(np.uint8).astype(np.int16))
                hist += np.bincount(np.minimum(d.ravel(), 2), minlength=3)
        assert eng.rng_tell() == g.pos
    print("delta histogram (0, 1, >=2):", hist.tolist())
    assert hist[2] == 0
    assert hist[1] <= 0.002 * hist.sum()


@pytest.mark.parametrize("w,h,n,argv", [(720, 480, 3, ["-vhs"]), (724, 482, 2, ["-vhs", "-vhs-speed", "ep"]),
                                        (720, 480, 2, []), (101, 67, 2, ["-vhs", "-comp-catv", "-vhs-svideo", "1"])])
def test_gpu_fp32_is_bit_identical_to_cpu_emulation(emu, w, h, n, argv):
    """The fp32 kernels fuse exactly where lane_pipeline.cuh says so (-fmad=false), so the GPU and the
    host emulation of the same source agree bit for bit -- the CPU parity suite therefore speaks for
    the GPU arithmetic as well."""
    p = helpers.params(*argv)
    frames = lambda k: helpers.stream_frame(w, h, k)
    want, _ = helpers.run_emu(emu, p, frames, n, w, h, precision=0)
    got = np.zeros((h, w), dtype=np.uint32)
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=1) as eng:
        for k in range(n):
            eng.composite_layer(got, frames(k), (k & 1) ^ 1, k)
    assert np.array_equal(want, got)


@pytest.mark.parametrize("w,h,n,argv", [(720, 480, 4, ["-vhs"]), (724, 482, 3, ["-vhs", "-vhs-speed", "ep"]),
                                        (720, 480, 2, []), (1920, 1080, 2, ["-vhs", "-vhs-speed", "sp"]),
                                        (101, 67, 3, ["-vhs", "-comp-catv", "-vhs-svideo", "1"]),
                                        (3840, 2160, 1, ["-vhs", "-vhs-speed", "lp"])])   # head-switch pre-pass rows
def test_fast_noise_mode(oracle, emu, w, h, n, argv):
    """cvs_set_noise_mode(CVS_NOISE_FAST): per-pixel noise from counter generators (the kern_n_* instantiations).
    Within +-1 LSB of the reference, bit-identical to the CPU emulation of the same mode, the same pictures from a
    batch as from sequential calls, and the rand() position advances exactly as in exact mode."""
    p = helpers.params(*argv)
    frames = lambda k: helpers.stream_frame(w, h, k)
    want, g = helpers.run_oracle(oracle, p, frames, n, w, h)
    got = np.zeros((h, w), dtype=np.uint32)
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=n) as eng:
        eng.set_noise_mode(True)
        for k in range(n):
            eng.composite_layer(got, frames(k), (k & 1) ^ 1, k)
        assert eng.rng_tell() == g.pos
        mx, nd, n2 = helpers.channel_diff(want, got)
        assert mx <= 1 and n2 == 0, (mx, nd, n2)
        assert nd <= 0.03 * want.size * 4, nd
        if w <= 1920:
            emu_out, _ = helpers.run_emu(emu, p, frames, n, w, h, precision=0, noise_fast=1)
            assert np.array_equal(emu_out, got)
        # batch == sequential in this mode too (the generators are seeded per field and row)
        src = np.stack([frames(k) for k in range(n)])
        bat = np.zeros((n, h, w), dtype=np.uint32)
        eng.rng_seek(0)
        eng.composite_fields_host(bat, src, 0)
        seq = np.zeros((n, h, w), dtype=np.uint32)
        eng.rng_seek(0)
        for k in range(n):
            eng.composite_layer(seq[k], src[k], (k & 1) ^ 1, k)
        assert np.array_equal(bat, seq)
        # and the fp64 validation mode ignores it: bit-exact
        eng.set_precision(True)
        eng.rng_seek(0)
        exact = np.zeros((h, w), dtype=np.uint32)
        for k in range(n):
            eng.composite_layer(exact, frames(k), (k & 1) ^ 1, k)
        assert np.array_equal(exact, want)


@pytest.mark.parametrize("w,h", [(160, 120), (164, 121)])
def test_fused_bob_matches_reference_loop(oracle, w, h):
    """composite_layer + the line doubling of the main loop (ffmpeg_ntsc.cpp:2229-2257) over a reused
    picture, against the engine with the doubling fused into the store (host and device forms)."""
    import ctypes as C
    import torch
    n = 4
    p = helpers.params("-vhs")
    frames = [helpers.stream_frame(w, h, k) for k in range(n)]
    g = helpers.OracleRng()
    oracle.oracle_rng_seed(C.byref(g), 1)
    want = np.full((h, w), 0x00123456, dtype=np.uint32)
    got = want.copy()
    dev = torch.from_numpy(want.view(np.int32).copy()).cuda()
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=1) as eng, cvs.Engine(params=p, max_w=w, max_h=h, max_batch=1) as eng2:
        eng.set_precision(True)
        eng2.set_precision(True)
        eng.set_bob(True)
        eng2.set_bob(True)
        for k in range(n):
            f = (k & 1) ^ 1
            oracle.oracle_composite_layer(C.byref(p), C.byref(g), want.ctypes.data_as(C.c_void_p), 4 * w,
                                          frames[k].ctypes.data_as(C.c_void_p), 4 * w, w, h, 0, 0, f, C.c_ulonglong(k))
            oracle.oracle_bob(want.ctypes.data_as(C.c_void_p), 4 * w, w, h, f)
            eng.composite_layer(got, frames[k], f, k)
            assert np.array_equal(want, got), k
            src = torch.from_numpy(frames[k].view(np.int32)).cuda()
            eng2.composite_fields_device(dev, src, 1, h, w, k)
            eng2.synchronize()
            assert np.array_equal(want, dev.cpu().numpy().view(np.uint32)), k


def test_async_host_calls_overlap_correctly():
    """A stream of cvs_composite_fields_host_async batches (two host buffer sets, device buffers
    alternate inside the engine) equals the synchronous calls."""
    w, h, n, calls = 320, 240, 5, 5
    p = helpers.params("-vhs")
    src = [np.stack([helpers.stream_frame(w, h, c * n + k) for k in range(n)]) for c in range(calls)]
    want = [np.zeros((n, h, w), dtype=np.uint32) for _ in range(calls)]
    got = [np.zeros((n, h, w), dtype=np.uint32) for _ in range(calls)]
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=n) as eng:
        for c in range(calls):
            eng.composite_fields_host(want[c], src[c], c * n)
        pos = eng.rng_tell()
        eng.rng_seek(0)
        for c in range(calls):
            eng.composite_fields_host_async(got[c], src[c], c * n)
        eng.synchronize()
        assert eng.rng_tell() == pos
    for c in range(calls):
        assert np.array_equal(want[c], got[c]), c


def test_preferred_batch_is_wave_aligned():
    """cvs_preferred_batch: the largest batch <= max whose warp tasks fill whole waves of the GPU (packed mapping:
    b fields of 540 rows are ceil(540 b / 31) tasks)."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    with cvs.Engine(["-vhs"], max_w=1920, max_h=1080, max_batch=320) as eng:
        b = eng.preferred_batch(1920, 1080, 320)
        assert 1 <= b <= 320
        tasks = lambda n: (540 * n + 30) // 31
        # some residency of 1..4 CTAs x 4 warps per SM makes b the wave-aligned batch
        ok = False
        for ctas in (1, 2, 3, 4):
            slots = sms * ctas * 4
            waves = tasks(320) // slots
            ok |= waves >= 1 and tasks(b) <= waves * slots < tasks(b + 1)
        assert ok, b
        assert eng.preferred_batch(1920, 1080, 1) == 1      # less than one wave: unchanged
        assert eng.preferred_batch(1920, 1080, 10**6) <= 320  # clamped to the context's capacity


# SURVEY.md App. D known answers (FNV-1a-64 of the reused dst picture after n fields of 75 % colour bars, default
# seed), reproduced by the CUDA path in the fp64 validation mode: the same hashes the reference's own code gives.
KAT = [
    (720, 480, "sp", 4, 0xadc46d5e6368519b),
    (720, 480, "sp", 60, 0x6efec46688da8257),        # BASELINE config 1
    (720, 480, "comp", 60, 0xdddc5e14c4d9b4a9),
    (1920, 1080, "sp", 32, 0x69cd6d98c07172ab),
    (1920, 1080, "ep", 16, 0x34b7815ed20c013d),
    (3840, 2160, "comp", 8, 0x950d9b8328d39f81),
]
KAT_MODES = {"sp": ["-vhs", "-vhs-speed", "sp"], "ep": ["-vhs", "-vhs-speed", "ep"], "comp": []}


@pytest.mark.parametrize("w,h,mode,n,want", KAT)
def test_known_answer_hashes_fp64(w, h, mode, n, want):
    bars = helpers.bars_frame(w, h)
    dst = np.zeros((h, w), dtype=np.uint32)
    with cvs.Engine(KAT_MODES[mode], max_w=w, max_h=h, max_batch=1) as eng:
        eng.set_precision(True)
        for k in range(n):
            eng.composite_layer(dst, bars, (k & 1) ^ 1, k)
        assert eng.rng_tell() == sum(cvs.draws_per_field(eng.params, w, h, (k & 1) ^ 1) for k in range(n))
    assert helpers.fnv1a64(dst) == want


def test_config1_480p_vhs_60_colour_bar_fields_fp32(oracle):
    """BASELINE config 1 in production arithmetic: 60 colour-bar fields at 720x480 `-vhs`, within +-1 LSB of the
    CPU path, as one batch and with the fields in a reused picture as the reference's loop has them."""
    w, h, n = 720, 480, 60
    p = helpers.params("-vhs")
    bars = helpers.bars_frame(w, h)
    want, g = helpers.run_oracle(oracle, p, lambda k: bars, n, w, h)
    assert helpers.fnv1a64(want) == 0x6efec46688da8257
    src = np.ascontiguousarray(np.broadcast_to(bars, (n, h, w)))
    got = np.zeros((n, h, w), dtype=np.uint32)
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=n) as eng:
        eng.composite_fields_host(got, src, 0)
        assert eng.rng_tell() == g.pos
    # the reused picture after 60 calls = odd rows of field 58 (field parity 1) and even rows of field 59
    last = np.zeros((h, w), dtype=np.uint32)
    last[1::2] = got[58][1::2]
    last[0::2] = got[59][0::2]
    mx, nd, n2 = helpers.channel_diff(want, last)
    assert mx <= 1 and n2 == 0 and nd <= 0.005 * want.size * 4, (mx, nd, n2)


def test_failed_call_consumes_no_draws():
    """include/cvs_ntsc.h: a call that fails leaves the rand() position untouched (the reference returns before its
    first draw, ffmpeg_ntsc.cpp:1578-1583)."""
    w, h = 64, 480
    with cvs.Engine(["-vhs"], max_w=w, max_h=h, max_batch=2) as eng:
        eng.rng_seek(12345)
        src = np.stack([helpers.stream_frame(w, h, k) for k in range(2)])
        dst = np.zeros_like(src)
        with pytest.raises(ValueError):
            eng.composite_fields_host(dst[:, :, : w - 1], src, 0)              # shapes differ: rejected before the ABI
        assert eng.rng_tell() == 12345
        with pytest.raises(cvs.CvsError) as e:
            eng.composite_layer(dst[0], src[0], 2, 0)                          # field > 1
        assert e.value.status == -1 and eng.rng_tell() == 12345
        eng.composite_fields_host(dst, src, 0)
        assert eng.rng_tell() == 12345 + sum(cvs.draws_per_field(eng.params, w, h, f) for f in (1, 0))


def test_phase_table_survives_precision_switches(oracle):
    """The {sin, cos} table of the chroma phase noise has one device copy per precision; growing one must not let
    the other be overrun (a larger -chroma-phase-noise after a precision round trip)."""
    w, h = 160, 120
    frames = lambda k: helpers.stream_frame(w, h, k)
    with cvs.Engine(["-vhs", "-chroma-phase-noise", "2"], max_w=w, max_h=h, max_batch=1) as eng:
        got = np.zeros((h, w), dtype=np.uint32)
        eng.composite_layer(got, frames(0), 1, 0)                 # float table, 5 states
        p = helpers.params("-vhs", "-chroma-phase-noise", "40")
        eng.set_params(p)
        eng.set_precision(True)
        eng.rng_seek(0)
        eng.composite_layer(got, frames(0), 1, 0)                 # double table, 81 states
        eng.set_precision(False)
        eng.rng_seek(0)
        got = np.zeros((h, w), dtype=np.uint32)
        for k in range(3):
            eng.composite_layer(got, frames(k), (k & 1) ^ 1, k)   # float again, 81 states
    want, _ = helpers.run_oracle(oracle, p, frames, 3, w, h)
    mx, nd, n2 = helpers.channel_diff(want, got)
    assert mx <= 1 and n2 == 0, (mx, nd, n2)


def test_last_row_of_an_exactly_sized_device_buffer():
    """A field whose last row is the last row of the allocation (field 1, even height): the kernel must not read
    past the end of the row (run under compute-sanitizer by scripts/sanitizer_probe.py; here: it runs and equals
    the same picture processed inside a larger allocation)."""
    import torch
    w, h = 1024, 512                                  # 2 MiB per picture
    p = helpers.params("-vhs")
    pic = torch.from_numpy(helpers.stream_frame(w, h, 3).view(np.int32))
    exact = pic.cuda()
    big = torch.zeros((2, h, w), dtype=torch.int32, device="cuda")
    big[0] = exact
    outs = []
    with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=1) as eng:
        for src in (exact, big):
            dst = torch.zeros((h, w), dtype=torch.int32, device="cuda")
            eng.rng_seek(0)
            eng.composite_fields_device(dst, src, 1, h, w, 0)     # fieldno 0 -> field 1: rows 1, 3, .. h-1
            eng.synchronize()
            outs.append(dst.cpu())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("warm_px", ["0", "5"])
@pytest.mark.parametrize("w,h,n,argv", [(720, 480, 3, ["-vhs"]), (101, 67, 2, ["-vhs", "-vhs-speed", "ep"]),
                                        (3840, 2160, 1, ["-vhs", "-vhs-speed", "lp"])])    # (pre-pass rows warm up too)
def test_noise_warmup_second_chance(oracle, w, h, n, argv, warm_px):
    """CVS_ERR_NOISE_SYNC is not something a caller has to handle: a lane whose noise warm-up does not converge walks
    the generator backward and tries again in the kernel (scanline_kernels.cuh).  CVS_WARM_PX shortens the first
    attempt so that nearly every row takes that path; fp64 must still be the reference bit for bit."""
    p = helpers.params(*argv)
    frames = lambda k: helpers.noise_frame(w, h, 11 + k)
    want, g = helpers.run_oracle(oracle, p, frames, n, w, h)
    got = np.zeros((h, w), dtype=np.uint32)
    os.environ["CVS_WARM_PX"] = warm_px
    try:
        with cvs.Engine(params=p, max_w=w, max_h=h, max_batch=2) as eng:
            eng.set_precision(True)
            for k in range(n):
                eng.composite_layer(got, frames(k), (k & 1) ^ 1, k)
            assert eng.rng_tell() == g.pos
    finally:
        os.environ.pop("CVS_WARM_PX", None)
    assert np.array_equal(want, got)
