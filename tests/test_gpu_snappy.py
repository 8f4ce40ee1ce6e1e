"""GPU parity for spark.io.compression.codec=snappy (xerial SnappyOutputStream framing over raw Snappy): the CUDA path
through the C ABI vs the oracle, the real snappy library (pyarrow) and the committed golden streams."""
import json
import os

import numpy as np
import pytest

from conftest import KINDS, corpus

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))


@pytest.mark.parametrize("kind", KINDS)
def test_gpu_encode_equals_cpu_specification_and_decodes_everywhere(capi, oracle, kind):
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("snappy")
    sizes = [0, 1, 11, 12, 13, 14, 63, 64, 65, 1000, 32767, 32768, 32769, 70001, 300000]
    parts = [corpus(oracle, kind, n, seed=i) for i, n in enumerate(sizes)]
    comp, cks, st = capi.compress_batch(capi.CODEC_SNAPPY_XERIAL, parts, 32768, capi.CHECKSUM_CRC32)
    assert st == [0] * len(parts)
    for p, s, k in zip(parts, comp, cks):
        assert s == oracle.xerial_compress(p, 32768, compressor=1), "kernel output differs from its CPU specification"
        assert oracle.xerial_decompress(s) == p
        assert k == oracle.crc32(s)
        ip, out = 16, b""
        while ip < len(s):  # every chunk through the real snappy library
            clen = int.from_bytes(s[ip:ip + 4], "big")
            out += codec.decompress(s[ip + 4:ip + 4 + clen], decompressed_size=min(32768, len(p) - len(out))).to_pybytes()
            ip += 4 + clen
        assert out == p


def test_empty_stream_is_the_16_byte_header(capi):
    comp, _, st = capi.compress_batch(capi.CODEC_SNAPPY_XERIAL, [b""], 32768)
    assert st == [0] and comp[0].hex() == GOLD["kat"]["xerial_header_hex"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("compressor", [0, 1])
def test_cpu_encode_gpu_decode(capi, oracle, kind, compressor):
    sizes = [0, 1, 13, 100, 5000, 32768, 32769, 200000]
    parts = [corpus(oracle, kind, n, seed=3 + i) for i, n in enumerate(sizes)]
    streams = [oracle.xerial_compress(p, 32768, compressor=compressor) for p in parts]
    slices = [[(len(s), oracle.adler32(s))] for s in streams]
    out, st, _ = capi.decompress_batch(capi.CODEC_SNAPPY_XERIAL, streams, capi.CHECKSUM_ADLER32, slices)
    assert st == [0] * len(parts)
    assert out == parts


def test_gpu_decodes_golden_streams_from_the_real_snappy_library(capi):
    names, streams, want = [], [], []
    for name, c in GOLD["cases"].items():
        names.append(name)
        streams.append(bytes.fromhex(c["xerial_stream_hex"]))
        want.append(c["input_len"])
    sizes, st = capi.decompressed_size_batch(capi.CODEC_SNAPPY_XERIAL, streams)
    assert st == [0] * len(streams) and sizes == want
    out, st, _ = capi.decompress_batch(capi.CODEC_SNAPPY_XERIAL, streams)
    assert st == [0] * len(streams)
    import zlib
    for name, o in zip(names, out):
        assert zlib.crc32(o) == GOLD["cases"][name]["crc32"], name


def test_concatenated_streams_and_batch_slices(capi, oracle):
    a, b = corpus(oracle, "text", 50000, 1), corpus(oracle, "terasort", 40000, 2)
    sa, sb = oracle.xerial_compress(a), oracle.xerial_compress(b, compressor=1)
    block = sa + sb  # ShuffleBlockBatchId: two partitions' streams back to back; the header re-occurs mid-stream
    slices = [[(len(sa), oracle.crc32(sa)), (len(sb), oracle.crc32(sb))]]
    out, st, bad = capi.decompress_batch(capi.CODEC_SNAPPY_XERIAL, [block], capi.CHECKSUM_CRC32, slices)
    assert st == [0] and out[0] == a + b
    slices[0][1] = (len(sb), oracle.crc32(sb) ^ 1)
    out, st, bad = capi.decompress_batch(capi.CODEC_SNAPPY_XERIAL, [block], capi.CHECKSUM_CRC32, slices,
                                         dst_caps=[len(a) + len(b)])
    assert st == [capi.E_CHECKSUM] and bad == [1]


def test_corrupt_streams_are_rejected(capi, oracle):
    x = corpus(oracle, "text", 40000, 5)
    s = bytearray(oracle.xerial_compress(x))
    cases = []
    t = bytearray(s); t[0] ^= 1; cases.append(bytes(t))                    # bad magic
    t = bytearray(s); t[16:20] = (len(s)).to_bytes(4, "big"); cases.append(bytes(t))   # chunk length past the end
    t = bytearray(s); t[20] = 0xFF; t[21] = 0xFF; t[22] = 0xFF; t[23] = 0xFF; t[24] = 0xFF; cases.append(bytes(t))  # bad varint
    cases.append(bytes(s[:-3]))                                            # truncated
    t = bytearray(s); t[23] = 0x0A; t[24] = 0xFF; t[25] = 0xFF; cases.append(bytes(t))   # copy before start of output
    for c in cases:
        ref_ok = True
        try:
            oracle.xerial_decompress(c)
        except IOError:
            ref_ok = False
        out, st, _ = capi.decompress_batch(capi.CODEC_SNAPPY_XERIAL, [c], dst_caps=[len(x) + 64])
        assert (st[0] == 0) == ref_ok, (st, ref_ok)
        if ref_ok:
            assert out[0] == oracle.xerial_decompress(c)


def test_many_small_streams_packed(capi, oracle):
    rng = np.random.default_rng(4)
    parts = [oracle.gen_terasort(int(rng.integers(0, 10**6)), int(rng.integers(1, 700))).tobytes() for _ in range(3000)]
    src = np.frombuffer(b"".join(parts), dtype=np.uint8)
    ln = np.array([len(p) for p in parts], dtype=np.uint64)
    off = np.concatenate(([0], np.cumsum(ln)[:-1])).astype(np.uint64)
    cap = sum(capi.compress_bound(capi.CODEC_SNAPPY_XERIAL, 32768, len(p)) for p in parts)
    dst = np.empty(cap, dtype=np.uint8)
    w = capi.compress_packed(capi.CODEC_SNAPPY_XERIAL, src, off, ln, dst, 32768, capi.CHECKSUM_CRC32C)
    assert not w["status"].any()
    back = np.empty(src.size, dtype=np.uint8)
    sb = np.arange(len(parts) + 1, dtype=np.uint32)
    r = capi.decompress_packed(capi.CODEC_SNAPPY_XERIAL, dst, w["dst_off"], w["dst_len"], back, capi.CHECKSUM_CRC32C,
                               sb, w["dst_len"], w["checksums"])
    assert not r["status"].any() and r["total"] == src.size
    assert np.array_equal(back, src)
    i = 1234
    s = dst[int(w["dst_off"][i]):int(w["dst_off"][i] + w["dst_len"][i])].tobytes()
    assert oracle.xerial_decompress(s) == parts[i]


def test_host_mirror_with_snappy_codec(tmp_path, oracle):
    import uuid
    import spark_s3_shuffle_b200 as pkg
    from shuffle_model import decode_pairs, encode_pairs, oracle_read_partition
    host = pkg.host
    d = host.S3ShuffleDispatcher({"spark.app.id": "app-" + uuid.uuid4().hex[:8],
                                  "spark.shuffle.s3.rootDir": "file://" + str(tmp_path) + "/s",
                                  "spark.io.compression.codec": "snappy",
                                  "spark.shuffle.checksum.algorithm": "CRC32"})
    i = np.arange(60000, dtype=np.int64)
    w = host.S3ShuffleMapOutputWriter(d, 0, 0, 4)
    for r in range(4):
        with w.getPartitionWriter(r) as s:
            s.write(encode_pairs(i[i % 4 == r] % 50, i[i % 4 == r]))
    w.commitAllPartitions()
    w.close()
    rd = host.S3ShuffleReader(d, 0, [0], 0, 4, True)
    blocks = rd.read()
    k = np.concatenate([decode_pairs(b)[0] for _, b in blocks])
    v = np.concatenate([decode_pairs(b)[1] for _, b in blocks])
    assert np.array_equal(np.sort(v), i) and np.array_equal(k, v % 50)
    rd.close()
    for r in range(4):  # the unmodified reference's reader (oracle arithmetic) consumes the GPU-written files
        got = oracle_read_partition(oracle, d, 0, [0], r, "CRC32", codec="snappy")
        assert np.array_equal(decode_pairs(got[0][1])[1], i[i % 4 == r])
    d.close()
