"""Probe of the host<->device copy ceiling on the box (contiguous vs strided-2D, one- and two-way)."""
import ctypes, time, torch
rt = ctypes.CDLL("libcudart.so")
N = 256; H, W = 1080, 1920
host_in = torch.empty((N, H, W), dtype=torch.int32).pin_memory()
host_out = torch.empty((N, H, W), dtype=torch.int32).pin_memory()
dev_in = torch.empty((N, H, W), dtype=torch.int32, device="cuda")
dev_out = torch.empty((N, H, W), dtype=torch.int32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
GB = N * H * W * 4 / 1e9
def h2d():
    with torch.cuda.stream(s1): dev_in.copy_(host_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): host_out.copy_(dev_out, non_blocking=True)
def both(): h2d(); d2h()
print("contiguous H2D  %.1f GB/s" % (GB / timeit(h2d)))
print("contiguous D2H  %.1f GB/s" % (GB / timeit(d2h)))
t = timeit(both); print("contiguous both %.1f GB/s each way" % (GB / t))
# strided: every other row (the field rows), via torch strided copies
def h2d_f():
    with torch.cuda.stream(s1): dev_in[:, 1::2].copy_(host_in[:, 1::2], non_blocking=True)
def d2h_f():
    with torch.cuda.stream(s2): host_out[:, 1::2].copy_(dev_out[:, 1::2], non_blocking=True)
def both_f(): h2d_f(); d2h_f()
print("field-rows H2D  %.1f GB/s" % (GB / 2 / timeit(h2d_f)))
print("field-rows D2H  %.1f GB/s" % (GB / 2 / timeit(d2h_f)))
print("field-rows both %.1f GB/s each way" % (GB / 2 / timeit(both_f)))
