#!/usr/bin/env python
"""Task-sized codec calls from several threads: direct C-ABI calls (serialised per device) vs the host library's
group-commit queue (spark-s3-shuffle_b200/host/coalesce.h).  Every thread plays a map task that commits 200 partitions
of 671 KB (BASELINE config 2's shape) R times.    python tools/coalesce_bench.py > profiles/r1z_coalesce.json"""
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spark_s3_shuffle_b200 as pkg  # noqa: E402
from oracle import oracle  # noqa: E402  (synthetic records only)


def main():
    c, host = pkg.capi, pkg.host
    c.init(1)
    L, H = c.load(), host.load()
    d = host.S3ShuffleDispatcher({"spark.app.id": "bench", "spark.shuffle.s3.rootDir": "file:///tmp/b2s-coalesce"})
    nblk, blk, reps = 200, 6453 * 104, 6
    rows = []
    for threads in (1, 2, 4, 8):
        ctx = []
        for t in range(threads):
            cap = int(c.compress_bound(c.CODEC_LZ4BLOCK, 32768, blk))
            src, dst = c.HostBuffer(nblk * blk), c.HostBuffer(nblk * cap)   # pinned, as a JVM direct buffer would be
            src.array[:] = oracle.gen_terasort(t * 10_000_000, nblk * 6453)
            sp = (C.c_void_p * nblk)(*[src.ptr + i * blk for i in range(nblk)])
            dp = (C.c_void_p * nblk)(*[dst.ptr + i * cap for i in range(nblk)])
            ln = np.full(nblk, blk, dtype=np.uint64)
            cp = np.full(nblk, cap, dtype=np.uint64)
            ctx.append((src, dst, sp, dp, ln, cp, np.zeros(nblk, np.uint64), np.zeros(nblk, np.uint64), np.zeros(nblk, np.int32)))
        res = {}
        for mode in ("direct", "queue"):
            def task(t):
                src, dst, sp, dp, ln, cp, dl, ck, st = ctx[t]
                for _ in range(reps):
                    if mode == "direct":
                        rc = L.b2s_compress_batch(c.CODEC_LZ4BLOCK, 0, 32768, c.CHECKSUM_CRC32C, nblk, sp, ln.ctypes.data,
                                                  dp, cp.ctypes.data, dl.ctypes.data, ck.ctypes.data, st.ctypes.data)
                    else:
                        rc = H.b2sh_dispatcher_queue_compress(d._h, c.CODEC_LZ4BLOCK, 0, 32768, c.CHECKSUM_CRC32C, nblk, sp,
                                                              ln.ctypes.data, dp, cp.ctypes.data, dl.ctypes.data,
                                                              ck.ctypes.data, st.ctypes.data)
                    assert rc == 0 and not st.any()
            task(0)  # warm-up
            ths = [threading.Thread(target=task, args=(t,)) for t in range(threads)]
            t0 = time.perf_counter()
            for th in ths:
                th.start()
            for th in ths:
                th.join()
            dt = time.perf_counter() - t0
            res[mode] = {"wall_ms": round(dt * 1e3, 1), "ms_per_call": round(dt * 1e3 / reps, 2),
                         "GBps": round(threads * reps * nblk * blk / dt / 1e9, 2)}
        for x in ctx:
            x[0].free()
            x[1].free()
        rows.append({"task_threads": threads, "calls_per_thread": reps, "blocks_per_call": nblk,
                     "uncompressed_MB_per_call": round(nblk * blk / 1e6, 1), **res})
    print(json.dumps({"what": "map-task-sized b2s_compress_batch calls (LZ4Block + CRC32C, pinned host buffers) from N threads",
                      "queue": d.queueStatistics(), "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
