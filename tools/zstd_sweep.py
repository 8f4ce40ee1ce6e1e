#!/usr/bin/env python
"""BASELINE config 5: Zstandard, shuffle-block size sweep on one GPU (device-resident), GB/s of uncompressed bytes and
the fraction of the HBM roofline ((1+r)·U algorithmic bytes per step / time / measured peak).  Frames for the read leg
are written by libzstd level 3 on the host (what zstd-jni writes), the write leg uses the GPU encoder.
    python tools/zstd_sweep.py  > profiles/r1z_zstd_sweep.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import spark_s3_shuffle_b200 as pkg  # noqa: E402
import zstd_ref  # noqa: E402
from oracle import oracle  # noqa: E402  (synthetic records only)


def main():
    c = pkg.capi
    c.init(1)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    total = int(os.environ.get("B2S_SWEEP_MIB", "1024")) << 20  # per point (SURVEY asks 4 GiB; 1 GiB keeps the run short)
    rows = []
    for size in (4 << 10, 16 << 10, 64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20, 64 << 20):
        n = total // size
        recs = (size + 103) // 104
        base = oracle.gen_terasort(0, recs * min(n, 64)).tobytes()
        parts = [base[(i % 64) * recs * 104:(i % 64) * recs * 104 + size] for i in range(n)]
        src = np.frombuffer(b"".join(parts), dtype=np.uint8)
        off = (np.arange(n, dtype=np.uint64) * size)
        ln = np.full(n, size, dtype=np.uint64)
        d_src = c.dev_alloc(src.size)
        c.dev_memcpy(d_src, src.ctypes.data, src.size, 1)
        cap = int(c.compress_bound(c.CODEC_ZSTD, 32768, size)) * n
        d_cmp, d_out = c.dev_alloc(cap), c.dev_alloc(src.size)
        for _ in range(2):
            w = c.compress_dev(c.CODEC_ZSTD, d_src, off, ln, d_cmp, cap, 32768, level=3)
        tw = c.last_timing()["kernel_ms"]
        # read leg on libzstd level-3 frames
        frames = [zstd_ref.compress_stream(p, level=3) for p in parts[:64]]
        frames = [frames[i % 64] for i in range(n)]
        fsrc = np.frombuffer(b"".join(frames), dtype=np.uint8)
        fln = np.array([len(f) for f in frames], dtype=np.uint64)
        foff = np.concatenate(([0], np.cumsum(fln)[:-1])).astype(np.uint64)
        d_f = c.dev_alloc(fsrc.size)
        c.dev_memcpy(d_f, fsrc.ctypes.data, fsrc.size, 1)
        for _ in range(2):
            r = c.decompress_dev(c.CODEC_ZSTD, d_f, foff, fln, d_out, src.size)
        tr = c.last_timing()["kernel_ms"]
        assert not w["status"].any() and not r["status"].any() and r["total"] == src.size
        rg, rl = w["total"] / src.size, fsrc.size / src.size
        rows.append({"shuffle_block_bytes": size, "streams": n, "write_read_GBps": round(src.size / (tw + tr) / 1e6, 2),
                     "gpu_encode_GBps": round(src.size / tw / 1e6, 2), "gpu_encode_ratio": round(rg, 4),
                     "gpu_encode_roofline_frac": round((1 + rg) * src.size / (tw * 1e-3) / 1e9 / peak, 5),
                     "gpu_decode_libzstd3_GBps": round(src.size / tr / 1e6, 2), "libzstd3_ratio": round(rl, 4),
                     "gpu_decode_roofline_frac": round((1 + rl) * src.size / (tr * 1e-3) / 1e9 / peak, 5)})
        for p in (d_src, d_cmp, d_out, d_f):
            c.dev_free(p)
    doc = {"config": "BASELINE config 5: zstd level 3, shuffle-block size sweep 4 KiB .. 64 MiB, %d MiB per point, 1 x B200" % (total >> 20),
           "hbm_peak_GBps": peak, "note": "decode time includes the size pass (frames carry no content size); the read leg "
           "decodes frames written by libzstd level 3 (what zstd-jni writes), the write leg is the GPU encoder at level 3",
           "rows": rows}
    if "--bench-line" in sys.argv:  # bench.py --config 5: one JSON line in the bench contract's shape
        vals = sorted(r["write_read_GBps"] for r in rows)
        print(json.dumps({"metric": "shuffle write+read GB/s (compress+CRC) at 1/2/4/8 B200 vs JVM-LZ4 CPU baseline",
                          "value": vals[len(vals) // 2], "unit": "GB/s", "n_gpus": 1, "steps": 1, "warmup": 1,
                          "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "u8", "data": "synthetic",
                          "config": {"workload": doc["config"] + " (value = median over the block sizes)", "baseline_config": 5},
                          "sweep": doc}))
    else:
        print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
