"""On-hardware multi-GPU identity: the same field stream computed by N ranks, one process
per GPU, each taking its contiguous chunk of every step and seeking rand() by the closed form of
composite_video_simulator_b200.sharding, must be byte-identical to ONE GPU running the stream serially -- and both
must agree with the serial CPU oracle.  Skipped on a single-GPU box."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from composite_video_simulator_b200 import sharding

pytestmark = pytest.mark.gpu

W, H, B, STEPS = 720, 481, 5, 2             # odd height: the two field parities draw different amounts
ARGV = ["-vhs", "-vhs-speed", "sp"]


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _rank_main(rank, world, port, out_dir, use_double):
    import torch
    import torch.distributed as dist
    import composite_video_simulator_b200 as cvs
    from composite_video_simulator_b200.params import CvsParams
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    # the job's only collective besides the gather of the results: rank 0 broadcasts the parameter block
    buf = torch.zeros(C.sizeof(CvsParams), dtype=torch.uint8, device=dev)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(bytes(helpers.params(*ARGV))), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    p = CvsParams.from_buffer_copy(buf.cpu().numpy().tobytes())
    mine = torch.zeros((STEPS, B, H, W), dtype=torch.int32, device=dev)
    with cvs.Engine(params=p, device=rank, max_w=W, max_h=H, max_batch=B) as eng:
        eng.set_precision(use_double)
        for step in range(STEPS):
            first, count = sharding.chunk(step, rank, world, B)
            src = torch.from_numpy(np.stack([helpers.stream_frame(W, H, k) for k in range(first, first + count)]).view(np.int32)).to(dev)
            eng.rng_seek(sharding.stream_position(p, W, H, first))
            eng.composite_fields_device(mine[step], src, count, H, W, first)
        eng.synchronize()
    gathered = [torch.zeros_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, gathered, dst=0)
    if rank == 0:
        # [rank][step][i] -> field index chunk(step, rank)[0] + i
        pics = {}
        for r in range(world):
            g = gathered[r].cpu().numpy().view(np.uint32)
            for step in range(STEPS):
                first, count = sharding.chunk(step, r, world, B)
                for i in range(count):
                    pics[first + i] = g[step, i]
        n = world * B * STEPS
        assert sorted(pics) == list(range(n))
        # the same stream on ONE GPU, serially, in two odd-sized batches
        serial = torch.zeros((n, H, W), dtype=torch.int32, device=dev)
        src = torch.from_numpy(np.stack([helpers.stream_frame(W, H, k) for k in range(n)]).view(np.int32)).to(dev)
        with cvs.Engine(params=p, device=0, max_w=W, max_h=H, max_batch=n) as eng:
            eng.set_precision(use_double)
            cut = n // 2 + 1
            eng.composite_fields_device(serial, src, cut, H, W, 0)
            eng.composite_fields_device(serial[cut:], src[cut:], n - cut, H, W, cut)
            eng.synchronize()
            assert eng.rng_tell() == sharding.stream_position(p, W, H, n)
        ser = serial.cpu().numpy().view(np.uint32)
        same = all(np.array_equal(ser[k], pics[k]) for k in range(n))
        np.savez(os.path.join(out_dir, "result.npz"), same=np.array(same), pics=np.stack([pics[k] for k in range(n)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("use_double", [False, True])
def test_n_gpus_equal_one_gpu_equal_oracle(oracle, tmp_path, use_double):
    world = _ngpus()
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus N)")
    world = min(world, 8)
    import torch.multiprocessing as mp
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(world, port, str(tmp_path), use_double), nprocs=world, join=True)
    z = np.load(os.path.join(str(tmp_path), "result.npz"))
    assert bool(z["same"]), "sharded pictures differ from the one-GPU serial run"
    # against the serial CPU oracle: every field into a fresh picture
    p = helpers.params(*ARGV)
    g = helpers.OracleRng()
    oracle.oracle_rng_seed(C.byref(g), 1)
    n = world * B * STEPS
    worst = 0
    for k in range(n):
        dst = np.zeros((H, W), dtype=np.uint32)
        src = helpers.stream_frame(W, H, k)
        oracle.oracle_composite_layer(C.byref(p), C.byref(g), dst.ctypes.data_as(C.c_void_p), 4 * W,
                                      src.ctypes.data_as(C.c_void_p), 4 * W, W, H, 0, 0, sharding.field_parity(k),
                                      C.c_ulonglong(k))
        mx, nd, n2 = helpers.channel_diff(dst, z["pics"][k])
        worst = max(worst, mx)
        assert (mx == 0) if use_double else (mx <= 1 and n2 == 0), (k, mx, nd, n2)
    print("multi-GPU identity: %d ranks, %d fields, max |delta| vs oracle %d" % (world, n, worst))
