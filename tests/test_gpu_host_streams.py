"""GPU tests of the stream-level host pieces (spark-s3-shuffle_b200/host/streams.h, codec_adapter.h):
  * the pipelined reader of SURVEY.md §8(f)-2 — S3BufferedPrefetchIterator feeding batches of completed blocks to one
    b2s_decompress_batch each (storage/S3ShuffleReader.scala:98-110, storage/S3BufferedPrefetchIterator.scala:196-212);
  * the Spark CompressionCodec seam of §8(f)-1 — compressedOutputStream / compressedInputStream for lz4, snappy, zstd;
  * re-entrancy of the C ABI from several task threads (SURVEY.md §8(b) threading row).
"""
import io
import threading
import uuid

import numpy as np
import pytest

import spark_s3_shuffle_b200 as pkg
import zstd_ref
from conftest import corpus
from shuffle_model import decode_pairs, encode_pairs

pytestmark = pytest.mark.gpu
host = pkg.host


def conf_for(tmp_path, **extra):
    conf = {
        "spark.app.id": "app-" + uuid.uuid4().hex[:12],
        "spark.shuffle.s3.rootDir": "file://" + str(tmp_path) + "/spark-s3-shuffle",
        "spark.shuffle.checksum.enabled": True,
        "spark.shuffle.checksum.algorithm": "CRC32",
    }
    conf.update(extra)
    return conf


def write_job(d, n_maps, n_red, per_map):
    for m in range(n_maps):
        i = np.arange(m * per_map, (m + 1) * per_map, dtype=np.int64)
        w = host.S3ShuffleMapOutputWriter(d, 0, m, n_red)
        for r in range(n_red):
            sel = i[i % n_red == r]
            if sel.size:
                with w.getPartitionWriter(r) as s:
                    s.write(encode_pairs(sel % 977, sel))
        w.commitAllPartitions()
        w.close()


@pytest.mark.parametrize("codec", ["lz4", "snappy", "zstd"])
def test_pipelined_reader_decodes_in_batches_under_the_buffer_budget(tmp_path, codec):
    budget = 192 * 1024
    d = host.S3ShuffleDispatcher(conf_for(tmp_path, **{"spark.io.compression.codec": codec,
                                                       "spark.shuffle.s3.maxBufferSizeTask": budget,
                                                       "spark.shuffle.s3.maxConcurrencyTask": 4}))
    n_maps, n_red, per = 24, 5, 40_000
    write_job(d, n_maps, n_red, per)
    rd = host.S3ShuffleReader(d, 0, list(range(n_maps)), 1, 4)
    rd.open()
    seen, values, nbatch = [], [], 0
    while True:
        blocks = rd.nextBatch()
        if blocks is None:
            break
        nbatch += 1
        for bid, data in blocks:
            seen.append(bid)
            k, v = decode_pairs(data)
            assert np.array_equal(k, v % 977)
            values.append(v)
    assert sorted(seen) == sorted((m, r, r + 1) for m in range(n_maps) for r in range(1, 4))
    allv = np.arange(n_maps * per, dtype=np.int64)
    assert np.array_equal(np.sort(np.concatenate(values)), allv[(allv % n_red >= 1) & (allv % n_red < 4)])
    st = rd.statistics()
    assert st["batches"] == nbatch and nbatch > 1, "the budget holds fewer compressed bytes than the task reads"
    assert st["peakMemoryUsage"] <= budget and st["numStreams"] == len(seen) and st["bytesRead"] == rd.remoteBytesRead
    assert st["line"].startswith("Statistics: Stage 0.0 TID 0 -- %d bytes" % rd.remoteBytesRead)
    rd.close()
    d.close()


def test_next_batch_honours_max_blocks_and_read_keeps_everything(tmp_path):
    d = host.S3ShuffleDispatcher(conf_for(tmp_path))
    write_job(d, 6, 4, 9_000)
    rd = host.S3ShuffleReader(d, 0, list(range(6)), 0, 4)
    rd.open()
    n = 0
    while True:
        b = rd.nextBatch(maxBlocks=3)
        if b is None:
            break
        assert 1 <= len(b) <= 3
        n += len(b)
    assert n == 24
    blocks = rd.read()            # open() + nextBatch() until the end, all blocks kept
    assert len(blocks) == 24
    v = np.sort(np.concatenate([decode_pairs(b)[1] for _, b in blocks]))
    assert np.array_equal(v, np.arange(54_000))
    rd.close()
    d.close()


def test_block_larger_than_the_task_budget_is_read_through(tmp_path):
    """bsize = min(maxBufferSize, maxBytes) (storage/S3BufferedPrefetchIterator.scala:125): the adaptor buffers the
    head of the block, the codec stream reads the tail straight from the block stream."""
    d = host.S3ShuffleDispatcher(conf_for(tmp_path, **{"spark.shuffle.s3.maxBufferSizeTask": 20_000}))
    rng = np.random.default_rng(3)
    k = rng.integers(0, 2**31 - 1, 60_000)
    w = host.S3ShuffleMapOutputWriter(d, 0, 0, 2)
    with w.getPartitionWriter(0) as s:
        s.write(encode_pairs(k, k))
    with w.getPartitionWriter(1) as s:
        s.write(encode_pairs(k[:10], k[:10]))
    lens = w.commitAllPartitions()
    w.close()
    assert lens[0] > 20_000 > lens[1]
    rd = host.S3ShuffleReader(d, 0, [0], 0, 2)
    got = dict(rd.read())
    assert np.array_equal(decode_pairs(got[(0, 0, 1)])[0], k) and np.array_equal(decode_pairs(got[(0, 1, 2)])[0], k[:10])
    rd.close()
    d.close()


def test_pipelined_reader_reports_a_bad_partition_like_the_reference(tmp_path):
    d = host.S3ShuffleDispatcher(conf_for(tmp_path))
    write_job(d, 3, 3, 8_000)
    acc = host.S3ShuffleHelper.getPartitionLengths(d, 0, 2)
    path = d.getPath("data", 0, 2)
    raw = bytearray(open(path, "rb").read())
    raw[int(acc[1]) + 30] ^= 0x04
    open(path, "wb").write(raw)
    rd = host.S3ShuffleReader(d, 0, [0, 1, 2], 0, 3)
    rd.open()
    with pytest.raises(host.SparkException, match=r"Invalid checksum detected for shuffle_0_2_1$"):
        while rd.nextBatch() is not None:
            pass
    rd.close()
    d.close()


# ---- CompressionCodec seam -------------------------------------------------------------------------------------
def reference_decode(oracle, codec, stream):
    """what the JVM reader of that codec accepts: concatenated streams included"""
    if codec == "lz4":
        return oracle.lz4block_decompress(stream)
    if codec == "snappy":
        return oracle.xerial_decompress(stream)
    return zstd_ref.decompress(stream, cap=1 << 26)


@pytest.mark.parametrize("codec", ["lz4", "snappy", "zstd"])
def test_compressed_output_stream_emits_reference_readable_streams(tmp_path, oracle, codec):
    d = host.S3ShuffleDispatcher(conf_for(tmp_path, **{"spark.io.compression.codec": codec,
                                                       "spark.shuffle.s3.gpu.codecBufferSize": "1m"}))
    c = host.B200CompressionCodec(d)
    assert c.supportsConcatenationOfSerializedStreams()   # what storage/S3ShuffleReader.scala:57-60 asks the codec
    data = corpus(oracle, "terasort", 3_300_000, 4) + corpus(oracle, "text", 1_234_567, 5)
    sink = io.BytesIO()
    out = c.compressedOutputStream(sink)
    rng = np.random.default_rng(8)
    at = 0
    while at < len(data):                                 # a serializer's irregular writes
        n = int(rng.integers(1, 200_000))
        out.write(data[at:at + n])
        at += n
    out.flush()
    out.close()
    stream = sink.getvalue()
    assert out.bytesIn == len(data) and out.bytesOut == len(stream)
    assert out.streams == -(-len(data) // (1 << 20))      # one complete stream per codecBufferSize + the tail
    assert reference_decode(oracle, codec, stream) == data
    inp = c.compressedInputStream(io.BytesIO(stream))      # and back through the adapter, in odd-sized pulls
    got = []
    while True:
        b = inp.read(77_777)
        if not b:
            break
        got.append(b)
    assert b"".join(got) == data
    inp.close()
    c.close()
    d.close()


@pytest.mark.parametrize("codec", ["lz4", "snappy", "zstd"])
def test_codec_streams_empty_and_reference_written_input(tmp_path, oracle, codec):
    d = host.S3ShuffleDispatcher(conf_for(tmp_path, **{"spark.io.compression.codec": codec}))
    c = host.B200CompressionCodec(d)
    sink = io.BytesIO()
    c.compressedOutputStream(sink).close()                 # nothing written: still a valid, self-terminated stream
    empty = sink.getvalue()
    assert len(empty) > 0 and reference_decode(oracle, codec, empty) == b""
    assert c.compressedInputStream(io.BytesIO(empty)).read() == b""
    data = corpus(oracle, "text", 500_000, 9)
    ref = {"lz4": lambda: oracle.lz4block_compress(data), "snappy": lambda: oracle.xerial_compress(data),
           "zstd": lambda: zstd_ref.compress_stream(data, level=1)}[codec]()
    assert c.compressedInputStream(io.BytesIO(ref)).read() == data   # streams of the unmodified JVM-side writers
    with pytest.raises(host.IOException, match="Stream is corrupted"):
        c.compressedInputStream(io.BytesIO(ref[: len(ref) // 2])).read()
    c.close()
    d.close()


# ---- threading ---------------------------------------------------------------------------------------------------
def test_c_abi_is_reentrant_across_task_threads(capi, oracle):
    """spark.executor.cores task threads call the codec concurrently (SURVEY.md §8(b) threading): every thread's
    batches must come back exact, with its own checksums."""
    errors = []

    def task(t):
        try:
            for rep in range(3):
                parts = [corpus(oracle, ("terasort", "text", "random")[(t + i) % 3], 40_000 + 9_000 * i + t, seed=t * 31 + i)
                         for i in range(6)]
                codec = (capi.CODEC_LZ4BLOCK, capi.CODEC_SNAPPY_XERIAL, capi.CODEC_ZSTD)[(t + rep) % 3]
                comp, cks, st = capi.compress_batch(codec, parts, 32768, capi.CHECKSUM_CRC32C)
                assert st == [0] * 6
                assert cks == [oracle.crc32c(s) for s in comp]
                slices = [[(len(s), k)] for s, k in zip(comp, cks)]
                out, st, _ = capi.decompress_batch(codec, comp, capi.CHECKSUM_CRC32C, slices)
                assert st == [0] * 6 and out == parts
        except BaseException as e:  # noqa: BLE001 - reported on the main thread
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=task, args=(t,)) for t in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


# ---- group commit across task threads ------------------------------------------------------------------------------
def test_group_commit_merges_the_calls_of_concurrent_task_threads(tmp_path, capi, oracle):
    """spark-s3-shuffle_b200/host/coalesce.h: calls submitted while the GPU is busy are merged into one C-ABI batch;
    every caller still gets exactly its own streams, checksums and status."""
    d = host.S3ShuffleDispatcher(conf_for(tmp_path))
    errors, n_threads, reps = [], 8, 6

    def task(t):
        try:
            for rep in range(reps):
                parts = [corpus(oracle, ("terasort", "text", "ints")[(t + i) % 3], 150_000 + 7_000 * i + 13 * t, seed=t * 17 + i)
                         for i in range(5)]
                comp, cks, st = d.queueCompress(capi.CODEC_LZ4BLOCK, parts, 32768, capi.CHECKSUM_CRC32,
                                                bound=lambda n: int(capi.compress_bound(capi.CODEC_LZ4BLOCK, 32768, n)))
                assert st == [0] * 5
                assert [oracle.lz4block_decompress(s) for s in comp] == parts      # the unmodified reader's arithmetic
                assert cks == [oracle.crc32(s) for s in comp]
                out, st = d.queueDecompress(capi.CODEC_LZ4BLOCK, comp, [len(p) for p in parts])
                assert st == [0] * 5 and out == parts
        except BaseException as e:  # noqa: BLE001 - reported on the main thread
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=task, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    st = d.queueStatistics()
    assert st["calls"] == n_threads * reps * 2 and st["streams"] == st["calls"] * 5
    assert st["batches"] < st["calls"] and st["maxMerged"] >= 2, st     # something was merged
    d.close()


def test_writer_and_reader_through_the_group_commit_queue(tmp_path):
    """spark.shuffle.s3.gpu.coalesce=true: four map tasks commit and four reduce tasks read concurrently through one
    dispatcher; files and results are the same as without the queue."""
    d = host.S3ShuffleDispatcher(conf_for(tmp_path, **{"spark.shuffle.s3.gpu.coalesce": True}))
    n_maps, n_red, per = 8, 4, 30_000
    errors = []

    def map_task(m):
        try:
            i = np.arange(m * per, (m + 1) * per, dtype=np.int64)
            w = host.S3ShuffleMapOutputWriter(d, 0, m, n_red)
            for r in range(n_red):
                sel = i[i % n_red == r]
                with w.getPartitionWriter(r) as s:
                    s.write(encode_pairs(sel % 977, sel))
            w.commitAllPartitions()
            w.close()
        except BaseException as e:  # noqa: BLE001
            errors.append(("map", m, repr(e)))

    ths = [threading.Thread(target=map_task, args=(m,)) for m in range(n_maps)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errors, errors
    got = {}

    def reduce_task(r):
        try:
            rd = host.S3ShuffleReader(d, 0, list(range(n_maps)), r, r + 1)
            got[r] = np.sort(np.concatenate([decode_pairs(b)[1] for _, b in rd.read()]))
            rd.close()
        except BaseException as e:  # noqa: BLE001
            errors.append(("reduce", r, repr(e)))

    ths = [threading.Thread(target=reduce_task, args=(r,)) for r in range(n_red)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert not errors, errors
    allv = np.arange(n_maps * per, dtype=np.int64)
    for r in range(n_red):
        assert np.array_equal(got[r], allv[allv % n_red == r])
    st = d.queueStatistics()
    assert st["calls"] >= n_maps + n_red
    # the same files without the queue: byte-identical .data objects
    d2 = host.S3ShuffleDispatcher(conf_for(tmp_path, **{"spark.app.id": "app-direct"}))
    i = np.arange(0, per, dtype=np.int64)
    w = host.S3ShuffleMapOutputWriter(d2, 0, 0, n_red)
    for r in range(n_red):
        with w.getPartitionWriter(r) as s:
            s.write(encode_pairs(i[i % n_red == r] % 977, i[i % n_red == r]))
    w.commitAllPartitions()
    w.close()
    assert open(d2.getPath("data", 0, 0), "rb").read() == open(d.getPath("data", 0, 0), "rb").read()
    d.close()
    d2.close()
