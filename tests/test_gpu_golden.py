"""GPU path against the committed golden vectors (tests/golden/vectors.json: liblz4 / zlib / xxhash / snappy outputs)."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))
NAMES = sorted(GOLDEN["cases"])


def test_gpu_decodes_golden_lz4block_streams(capi, oracle):
    streams = [bytes.fromhex(GOLDEN["cases"][k]["lz4block_stream_hex"]) for k in NAMES]
    sizes, st = capi.decompressed_size_batch(capi.CODEC_LZ4BLOCK, streams)
    assert st == [0] * len(NAMES) and sizes == [GOLDEN["cases"][k]["input_len"] for k in NAMES]
    # verify the stored per-stream checksum over the compressed bytes while decoding (ADLER32 = Spark's default)
    slices = [[(len(s), GOLDEN["cases"][k]["lz4block_stream_adler32"])] for s, k in zip(streams, NAMES)]
    out, st, _ = capi.decompress_batch(capi.CODEC_LZ4BLOCK, streams, capi.CHECKSUM_ADLER32, slices, dst_caps=sizes)
    assert st == [0] * len(NAMES)
    got_crc = capi.checksum_batch(capi.CHECKSUM_CRC32, out)
    got_adl = capi.checksum_batch(capi.CHECKSUM_ADLER32, out)
    for k, d, c, a in zip(NAMES, out, got_crc, got_adl):
        g = GOLDEN["cases"][k]
        assert c == g["crc32"] and a == g["adler32"], k
        if g["input_hex"] is not None:
            assert d.hex() == g["input_hex"]
        assert oracle.xxh32(d) == g["xxh32_seed9747b28c"]


def test_gpu_checksums_of_golden_streams(capi):
    streams = [bytes.fromhex(GOLDEN["cases"][k]["lz4block_stream_hex"]) for k in NAMES]
    assert capi.checksum_batch(capi.CHECKSUM_CRC32, streams) == [GOLDEN["cases"][k]["lz4block_stream_crc32"] for k in NAMES]
    assert capi.checksum_batch(capi.CHECKSUM_ADLER32, streams) == [GOLDEN["cases"][k]["lz4block_stream_adler32"] for k in NAMES]


def test_gpu_crc32c_rfc3720_vectors(capi):
    k = GOLDEN["kat"]
    msgs = [bytes(32), b"\xff" * 32, bytes(range(32)), bytes(range(31, -1, -1)), b"123456789"]
    want = [k["crc32c_32_zero_bytes"], k["crc32c_32_ff_bytes"], k["crc32c_32_incrementing"],
            k["crc32c_32_decrementing"], k["crc32c_123456789"]]
    assert capi.checksum_batch(capi.CHECKSUM_CRC32C, msgs) == want


def test_gpu_reencodes_golden_inputs(capi, oracle):
    """encode on the GPU what liblz4 encoded for the vectors; the restated JVM reader must give the same bytes back."""
    streams = [bytes.fromhex(GOLDEN["cases"][k]["lz4block_stream_hex"]) for k in NAMES]
    inputs = [oracle.lz4block_decompress(s) for s in streams]
    comp, cks, st = capi.compress_batch(capi.CODEC_LZ4BLOCK, inputs, 32768, capi.CHECKSUM_CRC32)
    assert st == [0] * len(NAMES)
    for k, x, c, s in zip(NAMES, inputs, comp, cks):
        assert oracle.lz4block_decompress(c) == x, k
        assert s == oracle.crc32(c)
        assert c == oracle.lz4block_compress(x, 32768, compressor=1), k
    assert comp[NAMES.index("empty")].hex() == GOLDEN["kat"]["lz4block_empty_stream_hex"]


def test_gpu_decodes_committed_libzstd_frames(capi):
    """tests/golden/zstd_vectors.json — frames written by libzstd.so.1 (what zstd-jni wraps), committed with the script
    that made them (tests/golden/make_golden_zstd.py): sized and decoded through the C ABI, CRC-32 of every output"""
    import json
    import os
    import zlib
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "zstd_vectors.json")))
    names, frames, want_len, want_crc = [], [], [], []
    for name, c in sorted(gold["cases"].items()):
        for kind, hexframe in sorted(c["frames"].items()):
            names.append((name, kind))
            frames.append(bytes.fromhex(hexframe))
            want_len.append(c["input_len"])
            want_crc.append(c["crc32"])
    sizes, st = capi.decompressed_size_batch(capi.CODEC_ZSTD, frames)
    assert st == [0] * len(frames) and sizes == want_len
    out, st, _ = capi.decompress_batch(capi.CODEC_ZSTD, frames)
    assert st == [0] * len(frames)
    for nm, o, n, crc in zip(names, out, want_len, want_crc):
        assert len(o) == n and zlib.crc32(o) == crc, nm
