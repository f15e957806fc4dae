"""Multi-rank host logic on CPU (gloo, world_size 2): each rank computes its contiguous chunk of the
field stream with the CPU emulation of the kernels, seeking rand() by the closed form of
composite_video_simulator_b200.sharding; gathered, the chunks must equal the serial oracle run."""
import os
import sys

import numpy as np
import pytest

import helpers
from composite_video_simulator_b200 import sharding

W, H, B, STEPS = 96, 65, 3, 2            # odd height: the two parities draw different amounts
ARGV = ["-vhs", "-vhs-speed", "sp"]


def _frames(k):
    return helpers.stream_frame(W, H, k)


def _rank_main(rank, world, port, out_dir):
    import ctypes as C
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = helpers.load_emu()
    # rank 0 owns the parameter block and broadcasts it (the job's only collective)
    from composite_video_simulator_b200.params import CvsParams
    buf = torch.zeros(C.sizeof(CvsParams), dtype=torch.uint8)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(bytes(helpers.params(*ARGV))), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    p = CvsParams.from_buffer_copy(buf.numpy().tobytes())
    outs = []
    for step in range(STEPS):
        first, count = sharding.chunk(step, rank, world, B)
        pos = C.c_ulonglong(sharding.stream_position(p, W, H, first))
        for k in range(first, first + count):
            dst = np.zeros((H, W), dtype=np.uint32)
            rc = emu.emu_composite_layer(C.byref(p), 1, C.byref(pos), dst.ctypes.data_as(C.c_void_p), 4 * W,
                                         _frames(k).ctypes.data_as(C.c_void_p), 4 * W, W, H, 0, 0,
                                         sharding.field_parity(k), C.c_ulonglong(k), 0)
            assert rc == 0
            outs.append((k, dst))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), ks=np.array([k for k, _ in outs]),
             pics=np.stack([d for _, d in outs]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_serial(oracle, tmp_path):
    import ctypes as C
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = {}
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        for k, pic in zip(z["ks"], z["pics"]):
            got[int(k)] = pic
    n = world * B * STEPS
    assert sorted(got) == list(range(n))
    # serial reference: one oracle, fields in order, each into a fresh picture
    p = helpers.params(*ARGV)
    g = helpers.OracleRng()
    oracle.oracle_rng_seed(C.byref(g), 1)
    for k in range(n):
        dst = np.zeros((H, W), dtype=np.uint32)
        src = _frames(k)
        assert g.pos == sharding.stream_position(p, W, H, k)
        oracle.oracle_composite_layer(C.byref(p), C.byref(g), dst.ctypes.data_as(C.c_void_p), 4 * W,
                                      src.ctypes.data_as(C.c_void_p), 4 * W, W, H, 0, 0, sharding.field_parity(k),
                                      C.c_ulonglong(k))
        assert np.array_equal(dst, got[k]), k


# ---- the 4:2:2 path -------------------------------------------------------------------------------

W2, H2 = 98, 65
ARGV2 = ["-vhs", "-vhs-speed", "lp"]


def _rank_main_yuv422(rank, world, port, out_dir):
    import ctypes as C
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from composite_video_simulator_b200 import yuv422
    emu = helpers.load_emu422()
    buf = torch.zeros(C.sizeof(yuv422.Yuv422Params), dtype=torch.uint8)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(bytes(helpers.params422(*ARGV2))), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    p = yuv422.Yuv422Params.from_buffer_copy(buf.numpy().tobytes())
    ks, pics = [], []
    for step in range(STEPS):
        first, count = sharding.chunk(step, rank, world, B)
        out, _ = helpers.run_emu422(emu, p, W2, H2, count, first=first,
                                    rng_pos=sharding.stream_position_yuv422(p, W2, H2, first))
        for i, (Y, U, V) in enumerate(out):
            ks.append(first + i)
            pics.append(np.concatenate([Y.ravel(), U.ravel(), V.ravel()]))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), ks=np.array(ks), pics=np.stack(pics))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_serial_yuv422(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_rank_main_yuv422, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = {}
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        for k, pic in zip(z["ks"], z["pics"]):
            got[int(k)] = pic
    n = world * B * STEPS
    assert sorted(got) == list(range(n))
    p = helpers.params422(*ARGV2)
    want, g = helpers.run_oracle422(helpers.load_oracle422(), p, W2, H2, n)
    assert g.pos == sharding.stream_position_yuv422(p, W2, H2, n)
    for k in range(n):
        Y, U, V = want[k]
        assert np.array_equal(np.concatenate([Y.ravel(), U.ravel(), V.ravel()]), got[k]), k
