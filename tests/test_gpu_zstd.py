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
