"""Generates tests/golden/zstd_vectors.json — Zstandard frames produced by libzstd.so.1 (the library zstd-jni wraps),
in the shapes the JVM writer emits: one-shot frames with content size (levels 1 and 3) and streaming frames without it
(ZSTD_compressStream2 fed 32 KiB at a time, optionally flushed => several blocks, Repeat_Mode tables, treeless literals).
Inputs are the ones of make_golden.py (its deterministic inputs()), so only the frames and a CRC-32 of the input are stored.
Run from the repo root:  python tests/golden/make_golden_zstd.py
"""
import json
import os
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import zstd_ref  # noqa: E402  (ctypes binding of libzstd.so.1)

sys.path.insert(0, HERE)
from make_golden import inputs  # noqa: E402

out = {"_generator": "tests/golden/make_golden_zstd.py", "_libzstd": "libzstd.so.1 (1.5.5)", "cases": {}}
for name, x in sorted(inputs().items()):
    frames = {
        "oneshot_l1": zstd_ref.compress(x, 1),
        "oneshot_l3": zstd_ref.compress(x, 3),
        "stream_l1": zstd_ref.compress_stream(x, 1, 32768, 0),
        "stream_l3_flush2": zstd_ref.compress_stream(x, 3, 32768, 2),
    }
    for f in frames.values():
        assert zstd_ref.decompress(f) == x
    out["cases"][name] = {"input_len": len(x), "crc32": zlib.crc32(x), "frames": {k: v.hex() for k, v in frames.items()}}
json.dump(out, open(os.path.join(HERE, "zstd_vectors.json"), "w"), indent=0, sort_keys=True)
print("wrote", sum(len(v["frames"]) for v in out["cases"].values()), "frames")
