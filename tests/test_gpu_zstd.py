"""GPU parity for spark.io.compression.codec=zstd, read side: frames written by libzstd.so.1 (what zstd-jni wraps), in
the shapes Spark's ZStdCompressionCodec produces (streaming frames without content size), decoded by the CUDA kernels
through the C ABI and compared with the input / with libzstd's own decoder."""
import numpy as np
import pytest

import zstd_ref
from conftest import KINDS, corpus

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", KINDS)
def test_gpu_decodes_libzstd_streaming_frames(capi, oracle, kind):
    sizes = [0, 1, 100, 5000, 32768, 131072, 131073, 400000, 700000]
    parts = [corpus(oracle, kind, n, seed=i) for i, n in enumerate(sizes)]
    frames = [zstd_ref.compress_stream(p, level=1 if i % 2 else 3, chunk=32768, flush_every=(3 if i % 3 == 0 else 0))
              for i, p in enumerate(parts)]
    got_sizes, st = capi.decompressed_size_batch(capi.CODEC_ZSTD, frames)
    assert st == [0] * len(parts) and got_sizes == sizes
    slices = [[(len(f), oracle.crc32(f))] for f in frames]
    out, st, _ = capi.decompress_batch(capi.CODEC_ZSTD, frames, capi.CHECKSUM_CRC32, slices)
    assert st == [0] * len(parts)
    assert out == parts


def test_gpu_decodes_one_shot_frames_levels_1_to_3(capi, oracle):
    parts, frames = [], []
    for lvl in (1, 2, 3):
        for kind in ("terasort", "text", "runs", "random"):
            p = corpus(oracle, kind, 150000 + 1000 * lvl, seed=lvl)
            parts.append(p)
            frames.append(zstd_ref.compress(p, lvl))
    out, st, _ = capi.decompress_batch(capi.CODEC_ZSTD, frames)
    assert st == [0] * len(parts) and out == parts


def test_concatenated_frames_and_per_partition_checksums(capi, oracle):
    a, b = corpus(oracle, "text", 90000, 1), corpus(oracle, "terasort", 60000, 2)
    fa, fb = zstd_ref.compress_stream(a, 1), zstd_ref.compress_stream(b, 3)
    block = fa + fb  # ShuffleBlockBatchId: two partitions' streams back to back
    slices = [[(len(fa), oracle.adler32(fa)), (len(fb), oracle.adler32(fb))]]
    out, st, bad = capi.decompress_batch(capi.CODEC_ZSTD, [block], capi.CHECKSUM_ADLER32, slices)
    assert st == [0] and out[0] == a + b
    slices[0][0] = (len(fa), oracle.adler32(fa) ^ 4)
    out, st, bad = capi.decompress_batch(capi.CODEC_ZSTD, [block], capi.CHECKSUM_ADLER32, slices,
                                         dst_caps=[len(a) + len(b)])
    assert st == [capi.E_CHECKSUM] and bad == [0]


def test_corrupt_and_truncated_frames_are_rejected_like_libzstd(capi, oracle):
    d = corpus(oracle, "terasort", 80000, 6)
    f = zstd_ref.compress_stream(d, 1)
    bad = [f[:-1], f[:len(f) // 2], b"\x00" + f[1:], f[:4] + bytes([f[4] | 0x08]) + f[5:]]
    for c in bad:
        with pytest.raises(IOError):
            zstd_ref.decompress(c)
    out, st, _ = capi.decompress_batch(capi.CODEC_ZSTD, bad + [f], dst_caps=[len(d) + 64] * (len(bad) + 1))
    assert all(s == capi.E_CORRUPT for s in st[:-1]) and st[-1] == 0 and out[-1] == d
    out, st, _ = capi.decompress_batch(capi.CODEC_ZSTD, [f], dst_caps=[len(d) - 1])
    assert st == [capi.E_DST_TOO_SMALL]


def test_many_shuffle_blocks_packed(capi, oracle):
    rng = np.random.default_rng(9)
    parts = [oracle.gen_terasort(int(rng.integers(0, 10**6)), int(rng.integers(1, 3000))).tobytes() for _ in range(600)]
    frames = [zstd_ref.compress_stream(p, 1) for p in parts]
    src = np.frombuffer(b"".join(frames), dtype=np.uint8)
    ln = np.array([len(f) for f in frames], dtype=np.uint64)
    off = np.concatenate(([0], np.cumsum(ln)[:-1])).astype(np.uint64)
    total = sum(len(p) for p in parts)
    dst = np.empty(total, dtype=np.uint8)
    r = capi.decompress_packed(capi.CODEC_ZSTD, src, off, ln, dst)
    assert not r["status"].any() and r["total"] == total
    assert dst.tobytes() == b"".join(parts)


# ---------------------------------------------------------------------------------------------------------------------
# write side: the GPU encoder (raw literals + predefined-FSE sequences) against its CPU model and against libzstd
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def zmodel(tmp_path_factory):
    import ctypes as C
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path_factory.mktemp("zc") / "libzc.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out,
                           os.path.join(root, "tests", "native", "zstd_core_host.cpp")])
    L = C.CDLL(out)
    L.zc_compress_model.restype = C.c_longlong
    L.zc_compress_model_hlog.restype = C.c_longlong
    L.zc_compress_model_hlog.argtypes = [C.c_char_p, C.c_ulonglong, C.c_uint, C.c_char_p, C.c_ulonglong, C.c_int]

    def model(data, block_size=32768, hash_log=12):
        cap = len(data) + len(data) // 64 + 1024
        buf = C.create_string_buffer(cap)
        n = L.zc_compress_model_hlog(data, len(data), block_size, buf, cap, hash_log)
        assert n > 0
        return buf.raw[:n]
    return model


@pytest.mark.parametrize("kind", KINDS)
def test_gpu_zstd_encode_equals_cpu_model_and_libzstd_decodes_it(capi, oracle, zmodel, kind):
    sizes = [0, 1, 12, 13, 100, 1000, 32767, 32768, 32769, 70001, 300000]
    parts = [corpus(oracle, kind, n, seed=i) for i, n in enumerate(sizes)]
    comp, cks, st = capi.compress_batch(capi.CODEC_ZSTD, parts, 32768, capi.CHECKSUM_CRC32C)
    assert st == [0] * len(parts)
    for p, f, k in zip(parts, comp, cks):
        assert zstd_ref.decompress(f) == p, "libzstd cannot read the GPU-written frame"
        assert f == zmodel(p), "kernel output differs from its CPU model"
        assert k == oracle.crc32c(f)
    # and back through the GPU decoder
    out, st, _ = capi.decompress_batch(capi.CODEC_ZSTD, comp)
    assert st == [0] * len(parts) and out == parts


def test_zstd_level_selects_the_match_finder_table(capi, oracle, zmodel):
    """spark.io.compression.zstd.level reaches the kernels: 1 -> 2^11-entry hash table, 2 / unspecified -> 2^12,
    >= 3 -> 2^13; every level's frames equal the CPU model at that table size and are read by libzstd"""
    p = corpus(oracle, "text", 300000, 5) + corpus(oracle, "terasort", 200000, 6)
    sizes = {}
    for level, hlog in ((0, 12), (1, 11), (2, 12), (3, 13), (9, 13)):
        comp, _, st = capi.compress_batch(capi.CODEC_ZSTD, [p], 32768, level=level)
        assert st == [0] and zstd_ref.decompress(comp[0]) == p
        assert comp[0] == zmodel(p, 32768, hlog), level
        sizes[level] = len(comp[0])
    assert sizes[3] <= sizes[2] <= sizes[1] and sizes[3] < sizes[1]


@pytest.mark.parametrize("bs", [64, 4096, 65536])
def test_gpu_zstd_block_size_sweep(capi, oracle, zmodel, bs):
    p = corpus(oracle, "text", 200000, 3) + corpus(oracle, "terasort", 100000, 4)
    comp, _, st = capi.compress_batch(capi.CODEC_ZSTD, [p], bs)
    assert st == [0] and zstd_ref.decompress(comp[0]) == p and comp[0] == zmodel(p, bs)


def test_host_mirror_with_zstd_codec(tmp_path, oracle):
    import uuid
    import spark_s3_shuffle_b200 as pkg
    from shuffle_model import decode_pairs, encode_pairs
    host = pkg.host
    d = host.S3ShuffleDispatcher({"spark.app.id": "app-" + uuid.uuid4().hex[:8],
                                  "spark.shuffle.s3.rootDir": "file://" + str(tmp_path) + "/s",
                                  "spark.io.compression.codec": "zstd"})
    i = np.arange(80000, dtype=np.int64)
    w = host.S3ShuffleMapOutputWriter(d, 0, 0, 3)
    for r in range(3):
        with w.getPartitionWriter(r) as s:
            s.write(encode_pairs(i[i % 3 == r] % 40, i[i % 3 == r]))
    lens = w.commitAllPartitions()
    w.close()
    # the unmodified reference's reduce side = libzstd over the partition's byte range
    data = open(d.getPath("data", 0, 0), "rb").read()
    acc = np.concatenate(([0], np.cumsum(lens)))
    for r in range(3):
        dec = zstd_ref.decompress(data[int(acc[r]):int(acc[r + 1])])
        assert np.array_equal(decode_pairs(dec)[1], i[i % 3 == r])
    rd = host.S3ShuffleReader(d, 0, [0], 0, 3, False)
    blocks = rd.read()
    v = np.concatenate([decode_pairs(b)[1] for _, b in blocks])
    assert np.array_equal(np.sort(v), i)
    rd.close()
    d.close()
