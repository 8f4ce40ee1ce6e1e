"""End-to-end jobs through the host mirror (writer -> .data/.index/.checksum -> reader) with the CUDA codec path
underneath — the GPU counterpart of the reference's only test file, src/test/scala/org/apache/spark/shuffle/
S3ShuffleManagerTest.scala (foldByKey :44, foldByKey_zeroBuffering :49, noMapSideCombine :56, forceSortShuffle :75,
combineByKey :103, teraSortLike :146).  Spark's RDD operators are restated with numpy; every shuffled byte goes
through b2s_compress_packed / b2s_decompress_packed via libb200shuffle_host.so.  Also checks interoperability with
the unmodified reference: GPU-written files are read by the ORACLE reader and oracle-written files by the GPU reader.
"""
import os
import uuid

import numpy as np
import pytest

import spark_s3_shuffle_b200 as pkg
from shuffle_model import decode_pairs, encode_pairs, oracle_read_partition

pytestmark = pytest.mark.gpu
host = pkg.host


def conf_for(tmp_path, **extra):
    conf = {
        "spark.app.id": "app-" + uuid.uuid4().hex[:12],
        "spark.shuffle.s3.rootDir": "file://" + str(tmp_path) + "/spark-s3-shuffle",
        "spark.shuffle.checksum.enabled": True,
        "spark.shuffle.checksum.algorithm": "ADLER32",
    }
    conf.update(extra)
    return conf


def shuffle_write(d, shuffle_id, map_parts, n_red, partitioner):
    """map_parts: list of (keys, values) per map task.  Returns the per-map partitionLengths."""
    out = []
    for m, (k, v) in enumerate(map_parts):
        w = host.S3ShuffleMapOutputWriter(d, shuffle_id, m, n_red)
        p = partitioner(k)
        for r in range(n_red):
            mask = p == r
            if not mask.any():
                continue
            with w.getPartitionWriter(r) as s:
                s.write(encode_pairs(k[mask], v[mask]))
        out.append(w.commitAllPartitions())
        w.close()
    return out


def shuffle_read(d, shuffle_id, n_maps, r0, r1, batch=False):
    rd = host.S3ShuffleReader(d, shuffle_id, list(range(n_maps)), r0, r1, batch)
    blocks = rd.read()
    ks = [decode_pairs(b)[0] for _, b in blocks] or [np.zeros(0, np.int64)]
    vs = [decode_pairs(b)[1] for _, b in blocks] or [np.zeros(0, np.int64)]
    nbytes = rd.remoteBytesRead
    rd.close()
    return np.concatenate(ks), np.concatenate(vs), nbytes, blocks


@pytest.mark.parametrize("alg", ["ADLER32", "CRC32", "CRC32C"])
def test_foldByKey(tmp_path, oracle, alg):
    """test/S3ShuffleManagerTest.scala:176-205: 10,000 ints, 3 maps, 5 reducers, per-key sums and the key set."""
    d = host.S3ShuffleDispatcher(conf_for(tmp_path, **{"spark.shuffle.checksum.algorithm": alg}))
    n, n_maps, n_red = 10_000, 3, 5
    i = np.arange(n, dtype=np.int64)
    parts = [(i[m::n_maps] % 7, i[m::n_maps]) for m in range(n_maps)]
    lens = shuffle_write(d, 0, parts, n_red, lambda k: k % n_red)
    sums = {}
    total_remote = 0
    for r in range(n_red):
        k, v, nbytes, _ = shuffle_read(d, 0, n_maps, r, r + 1)
        total_remote += nbytes
        for key in np.unique(k):
            assert key % n_red == r
            sums[int(key)] = int(v[k == key].sum())
    assert sorted(sums) == list(range(7))
    for key in range(7):
        assert sums[key] == int(i[i % 7 == key].sum())
    assert total_remote == int(sum(l.sum() for l in lens))  # incRemoteBytesRead parity (storage/S3ShuffleReader.scala:94)
    # interoperability: the unmodified reference's reduce side (oracle arithmetic) reads the GPU-written files
    for r in range(n_red):
        got = oracle_read_partition(oracle, d, 0, range(n_maps), r, alg)
        k = np.concatenate([decode_pairs(b)[0] for _, b in got] or [np.zeros(0, np.int64)])
        assert ((k % n_red) == r).all()
    d.removeShuffle(0)
    d.close()


def test_teraSortLike_and_forceSortShuffle(tmp_path):
    """:146-174 — 5 x 10,000 random (Int, Int), sortByKey(true, 4): range partitioner, sortedness across reducers."""
    d = host.S3ShuffleDispatcher(conf_for(tmp_path, **{"spark.shuffle.s3.bufferSize": 0}))  # zeroBuffering variant :49-54
    rng = np.random.default_rng(7)
    parts = [(rng.integers(-2**31, 2**31, 10_000), rng.integers(-2**31, 2**31, 10_000)) for _ in range(5)]
    allk = np.concatenate([p[0] for p in parts])
    bounds = np.quantile(allk, [0.25, 0.5, 0.75]).astype(np.int64)
    shuffle_write(d, 3, parts, 4, lambda k: np.searchsorted(bounds, k, side="right"))
    prev_max = -2**63
    seen = 0
    for r in range(4):
        k, v, _, _ = shuffle_read(d, 3, 5, r, r + 1)
        order = np.argsort(k, kind="stable")
        k = k[order]
        assert k.size == 0 or k[0] >= prev_max
        if k.size:
            prev_max = k[-1]
        seen += k.size
    assert seen == 50_000
    d.close()


def test_combineByKey_large(tmp_path):
    """:103-144 — 20 x 100,000 records; exact counts per key after the shuffle."""
    d = host.S3ShuffleDispatcher(conf_for(tmp_path))
    n_maps, per = 20, 100_000
    parts = []
    for m in range(n_maps):
        i = np.arange(m * per, (m + 1) * per, dtype=np.int64)
        parts.append((i % 1000, i))
    shuffle_write(d, 1, parts, 8, lambda k: k % 8)
    counts = np.zeros(1000, dtype=np.int64)
    for r in range(8):
        k, v, _, blocks = shuffle_read(d, 1, n_maps, r, r + 1)
        assert len(blocks) == n_maps
        counts += np.bincount(k, minlength=1000)
    assert (counts == n_maps * per // 1000).all()
    d.close()


def test_batch_fetch_multi_partition_blocks(tmp_path):
    """ShuffleBlockBatchId: one block spans several reduce partitions = concatenated codec streams, every slice
    verified separately (storage/S3ChecksumValidationStream.scala:22-27,68-86)."""
    d = host.S3ShuffleDispatcher(conf_for(tmp_path, **{"spark.shuffle.checksum.algorithm": "CRC32"}))
    i = np.arange(30_000, dtype=np.int64)
    parts = [(i[m::2], i[m::2] * 3) for m in range(2)]
    shuffle_write(d, 0, parts, 6, lambda k: k % 6)
    k, v, _, blocks = shuffle_read(d, 0, 2, 1, 5, batch=True)
    assert sorted(b[0] for b in blocks) == [(0, 1, 5), (1, 1, 5)]  # hand-off order is unspecified (LIFO)
    assert np.array_equal(np.sort(k), np.sort(i[(i % 6 >= 1) & (i % 6 < 5)]))
    assert np.array_equal(v, k * 3)
    d.close()


def test_corrupt_data_raises_the_reference_exceptions(tmp_path):
    d = host.S3ShuffleDispatcher(conf_for(tmp_path))
    i = np.arange(20_000, dtype=np.int64)
    shuffle_write(d, 0, [(i, i), (i, i + 1)], 3, lambda k: k % 3)
    acc = host.S3ShuffleHelper.getPartitionLengths(d, 0, 1)
    path = d.getPath("data", 0, 1)
    raw = bytearray(open(path, "rb").read())
    raw[int(acc[2]) + 40] ^= 0x10  # a payload byte of partition 2 of map 1
    open(path, "wb").write(raw)
    shuffle_read(d, 0, 2, 0, 2)  # untouched partitions still read
    with pytest.raises(host.SparkException, match=r"Invalid checksum detected for shuffle_0_1_2$"):
        shuffle_read(d, 0, 2, 2, 3)  # storage/S3ChecksumValidationStream.scala:72-74
    d.close()
    # with checksums disabled the codec's own block hash catches it: IOException("Stream is corrupted")
    d2 = host.S3ShuffleDispatcher(conf_for(tmp_path, **{"spark.shuffle.checksum.enabled": False,
                                                        "spark.app.id": "app-nochk"}))
    shuffle_write(d2, 0, [(i, i)], 1, lambda k: k * 0)
    path = d2.getPath("data", 0, 0)
    assert not os.path.exists(d2.getPath("checksum", 0, 0))
    raw = bytearray(open(path, "rb").read())
    raw[100] ^= 0xFF
    open(path, "wb").write(raw)
    with pytest.raises(host.IOException, match="Stream is corrupted"):
        shuffle_read(d2, 0, 1, 0, 1)
    d2.close()


def test_gpu_reader_consumes_files_written_by_the_reference_path(tmp_path, oracle):
    """Files laid out by the pass-through writer from oracle-compressed bytes (the unmodified reference's write side)
    are verified + decoded by the GPU reader."""
    d = host.S3ShuffleDispatcher(conf_for(tmp_path, **{"spark.shuffle.s3.gpu.enabled": False,
                                                       "spark.shuffle.checksum.algorithm": "CRC32"}))
    i = np.arange(50_000, dtype=np.int64)
    w = host.S3ShuffleMapOutputWriter(d, 0, 0, 3)
    cks = []
    for r in range(3):
        comp = oracle.lz4block_compress(encode_pairs(i[i % 3 == r], i[i % 3 == r]), 32768)
        cks.append(oracle.crc32(comp))
        with w.getPartitionWriter(r) as s:
            s.write(comp)
    w.commitAllPartitions(cks)
    w.close()
    for r in range(3):
        k, v, _, _ = shuffle_read(d, 0, 1, r, r + 1)
        assert np.array_equal(k, i[i % 3 == r]) and np.array_equal(v, k)
    d.close()


def test_empty_partitions_and_maps(tmp_path):
    d = host.S3ShuffleDispatcher(conf_for(tmp_path, **{"spark.shuffle.s3.alwaysCreateIndex": True}))
    k = np.array([4, 4, 4], dtype=np.int64)
    lens = shuffle_write(d, 0, [(k, k), (np.zeros(0, np.int64), np.zeros(0, np.int64))], 5, lambda x: x % 5)
    assert [int(x > 0) for x in lens[0]] == [0, 0, 0, 0, 1] and not lens[1].any()
    for r in range(5):
        kk, vv, nbytes, blocks = shuffle_read(d, 0, 2, r, r + 1)
        assert len(blocks) == (1 if r == 4 else 0)  # filterNot(maxBytes == 0), storage/S3ShuffleReader.scala:91
    d.close()


def test_single_spill_transfer_verifies_checksums_on_the_gpu(tmp_path, oracle):
    """the single-spill path with checksum-on-the-fly — per-partition checksums recomputed by the GPU
    over the spill file before it becomes the .data object."""
    d = host.S3ShuffleDispatcher(conf_for(tmp_path, **{"spark.shuffle.checksum.algorithm": "ADLER32"}))
    parts = [oracle.lz4block_compress(encode_pairs(np.arange(5000) % (r + 2), np.arange(5000)), 32768) for r in range(4)]
    lens = [len(p) for p in parts]
    cks = [oracle.adler32(p) for p in parts]
    spill = tmp_path / "spill.bin"
    spill.write_bytes(b"".join(parts))
    host.S3SingleSpillShuffleMapOutputWriter(d, 0, 0).transferMapSpillFile(spill, lens, cks, verifyOnTransfer=True)
    k, v, _, blocks = shuffle_read(d, 0, 1, 0, 4)
    assert len(blocks) == 4 and np.array_equal(np.sort(v), np.sort(np.tile(np.arange(5000), 4)))
    spill.write_bytes(b"".join(parts))
    bad = list(cks)
    bad[2] ^= 1
    with pytest.raises(host.SparkException, match=r"Invalid checksum detected for shuffle_0_1_2$"):
        host.S3SingleSpillShuffleMapOutputWriter(d, 0, 1).transferMapSpillFile(spill, lens, bad, verifyOnTransfer=True)
    d.close()
