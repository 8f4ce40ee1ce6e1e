"""Parity on the data shapes of BASELINE.json configs 4 and 5 (configs 1-3 live in test_host_plumbing.py,
test_gpu_parity.py and bench.py):
  config 4 — "sql join workload, Snappy codec": UnsafeRow-shaped binary rows
  config 5 — "Zstd-3 codec, block-size sweep 4 KB-64 MB": one Zstandard stream per shuffle block, sizes swept
"""
import numpy as np
import pytest

import zstd_ref

pytestmark = pytest.mark.gpu


def unsafe_rows(n_rows, seed):
    """4-byte length + 8-byte null bitmap + five 8-byte fields (join key, two low-cardinality ints, two zipf longs)"""
    rng = np.random.default_rng(seed)
    rows = np.zeros((n_rows, 6), dtype=np.int64)
    rows[:, 1] = rng.integers(0, 10**7, n_rows)
    rows[:, 2] = rng.integers(0, 1000, n_rows)
    rows[:, 3] = rng.integers(0, 1000, n_rows)
    rows[:, 4] = rng.zipf(1.2, n_rows) % 100000
    rows[:, 5] = rng.zipf(1.2, n_rows) % 100000
    body = rows.view(np.uint8).reshape(n_rows, 48)
    ln = np.full((n_rows, 1), 48, dtype=np.int32).view(np.uint8).reshape(n_rows, 4)
    return np.concatenate([ln, body], axis=1).tobytes()


def test_config4_sql_rows_snappy(capi, oracle):
    pa = pytest.importorskip("pyarrow")
    parts = [unsafe_rows(12603, s) for s in range(40)]  # ~655 KiB shuffle blocks, as 50 GiB / (400 maps x 200 partitions)
    comp, cks, st = capi.compress_batch(capi.CODEC_SNAPPY_XERIAL, parts, 32768, capi.CHECKSUM_ADLER32)
    assert st == [0] * len(parts)
    ratio = sum(map(len, comp)) / sum(map(len, parts))
    assert ratio < 0.6
    for p, s, k in zip(parts[:6], comp[:6], cks[:6]):
        assert s == oracle.xerial_compress(p, 32768, compressor=1)
        assert oracle.xerial_decompress(s) == p and k == oracle.adler32(s)
    # the real snappy library writes, the GPU reads
    codec = pa.Codec("snappy")
    hdr = bytes([0x82]) + b"SNAPPY\x00" + (1).to_bytes(4, "big") + (1).to_bytes(4, "big")
    streams = []
    for p in parts:
        s = hdr
        for o in range(0, len(p), 32768):
            c = codec.compress(p[o:o + 32768]).to_pybytes()
            s += len(c).to_bytes(4, "big") + c
        streams.append(s)
    out, st, _ = capi.decompress_batch(capi.CODEC_SNAPPY_XERIAL, streams + comp)
    assert st == [0] * (2 * len(parts)) and out == parts + parts


@pytest.mark.parametrize("size", [4096, 65536, 1 << 20, 16 << 20])
def test_config5_zstd3_shuffle_block_size_sweep(capi, oracle, size):
    total = 32 << 20
    n = max(1, total // size)
    recs = (size + 103) // 104
    parts = [oracle.gen_terasort(i * recs, recs).tobytes()[:size] for i in range(n)]
    frames = [zstd_ref.compress_stream(p, level=3) for p in parts]      # what zstd-jni level 3 writes
    sizes, st = capi.decompressed_size_batch(capi.CODEC_ZSTD, frames)
    assert st == [0] * n and sizes == [size] * n
    out, st, _ = capi.decompress_batch(capi.CODEC_ZSTD, frames, dst_caps=sizes)
    assert st == [0] * n and out == parts
    comp, _, st = capi.compress_batch(capi.CODEC_ZSTD, parts, 32768)     # our writer, libzstd as the reader
    assert st == [0] * n
    for p, f in list(zip(parts, comp))[:: max(1, n // 8)]:
        assert zstd_ref.decompress(f) == p
