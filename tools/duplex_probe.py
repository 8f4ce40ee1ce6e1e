"""PCIe duplex probe (one GPU): H2D alone, D2H alone, both at once on two streams, and two H2D streams competing — the
ceilings the host-staged path (bench.py e2e) lives under.  Pinned host memory, 1 GiB per copy.  JSON on stdout."""
import json
import time

import torch

N = 1 << 30
h1 = torch.empty(N, dtype=torch.uint8).pin_memory()
h2 = torch.empty(N, dtype=torch.uint8).pin_memory()
d1 = torch.empty(N, dtype=torch.uint8, device="cuda")
d2 = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
out = {"async_engines": torch.cuda.get_device_properties(0).multi_processor_count and None}
try:
    import ctypes
    rt = ctypes.CDLL("libcudart.so")
    v = ctypes.c_int()
    rt.cudaDeviceGetAttribute(ctypes.byref(v), 40, 0)  # cudaDevAttrAsyncEngineCount
    out["async_engines"] = v.value
except Exception as e:  # noqa: BLE001
    out["async_engines"] = repr(e)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


def h2d():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def both():
    h2d()
    d2h()


def two_h2d():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)


t = timed(h2d); out["h2d_GBps"] = round(N / t / 1e9, 2)
t = timed(d2h); out["d2h_GBps"] = round(N / t / 1e9, 2)
t = timed(both); out["duplex_each_GBps"] = round(N / t / 1e9, 2); out["duplex_sum_GBps"] = round(2 * N / t / 1e9, 2)
t = timed(two_h2d); out["two_h2d_streams_sum_GBps"] = round(2 * N / t / 1e9, 2)
print(json.dumps(out))
