"""The read-side plumbing between the object store and the codec, mirrored in C++ (spark-s3-shuffle_b200/host/streams.h):
S3ShuffleBlockStream (storage/S3ShuffleBlockStream.scala), S3BufferedInputStreamAdaptor
(storage/S3BufferedInputStreamAdaptor.scala), S3BufferedPrefetchIterator + ThreadPredictor
(storage/S3BufferedPrefetchIterator.scala) and, on the write side, S3MeasureOutputStream
(shuffle/S3MeasureOutputStream.scala).  No GPU: the blocks are opaque bytes written in pass-through mode.
"""
import re
import uuid

import numpy as np
import pytest

import spark_s3_shuffle_b200 as pkg

host = pkg.host


def new_conf(tmp_path, **extra):
    conf = {
        "spark.app.id": "app-" + uuid.uuid4().hex[:12],
        "spark.shuffle.s3.rootDir": "file://" + str(tmp_path) + "/spark-s3-shuffle",
        "spark.shuffle.checksum.enabled": False,
        "spark.shuffle.s3.gpu.enabled": False,
    }
    conf.update(extra)
    return conf


def write_maps(d, sizes, seed=5):
    """sizes[m][r] bytes of noise per (map, reduce); returns {(m, r): bytes}"""
    rng = np.random.default_rng(seed)
    blocks = {}
    for m, row in enumerate(sizes):
        w = host.S3ShuffleMapOutputWriter(d, 0, m, len(row))
        for r, n in enumerate(row):
            b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            with w.getPartitionWriter(r) as s:
                s.write(b)
            blocks[(m, r)] = b
        lens = w.commitAllPartitions(np.zeros(len(row), dtype=np.int64))
        assert list(lens) == list(row)
        w.close()
    return blocks


def drain(it, close=True):
    got = []
    while it.hasNext():
        bid, data, h = it.next()
        got.append((bid, data))
        if close:
            it.closeStream(h)
    return got


def test_every_non_empty_block_is_delivered_exactly_once(tmp_path):
    d = host.S3ShuffleDispatcher(new_conf(tmp_path))
    rng = np.random.default_rng(1)
    sizes = [[int(x) for x in rng.integers(0, 40_000, 7)] for _ in range(9)]
    sizes[3][2] = sizes[3][3] = 0          # empty partitions are filtered (storage/S3ShuffleReader.scala:91)
    sizes[5] = [0] * 7                     # a map without output has no index at all
    blocks = write_maps(d, sizes)
    maps = [m for m in range(9) if m != 5]
    it = host.S3BufferedPrefetchIterator(d, 0, maps, 1, 6)
    got = drain(it)
    want = {(m, r, r + 1): blocks[(m, r)] for m in maps for r in range(1, 6) if sizes[m][r]}
    assert len(got) == len(want)           # result order is unspecified (LIFO hand-off, :146,:209): compare as a map
    assert dict(got) == want
    st = it.statistics()
    assert st["numStreams"] == len(want) and st["bytesRead"] == sum(len(v) for v in want.values())
    assert re.match(r"Statistics: Stage 0\.0 TID 0 -- \d+ bytes, \d+ ms waiting \(\d+ avg\), \d+ ms prefetching "
                    r"\(avg: \d+ ms - \d+ block size - .* MiB/s\)\. Total: \d+ ms - \d+% waiting\. \d+ active threads\.",
                    st["line"])
    it.close()


def test_batch_fetch_yields_one_block_per_map(tmp_path):
    d = host.S3ShuffleDispatcher(new_conf(tmp_path))
    sizes = [[100, 0, 300, 50], [0, 0, 0, 9]]
    blocks = write_maps(d, sizes)
    it = host.S3BufferedPrefetchIterator(d, 0, [0, 1], 0, 3, doBatchFetch=True)
    got = dict(drain(it))
    assert got == {(0, 0, 3): blocks[(0, 0)] + blocks[(0, 1)] + blocks[(0, 2)]}   # map 1's range [0,3) is empty
    it.close()


def test_memory_budget_bounds_the_bytes_in_flight(tmp_path):
    d = host.S3ShuffleDispatcher(new_conf(tmp_path))
    sizes = [[20_000] * 8 for _ in range(6)]
    write_maps(d, sizes)
    budget = 64 * 1024                      # room for three 20,000-byte blocks
    it = host.S3BufferedPrefetchIterator(d, 0, range(6), 0, 8, maxBufferSize=budget, maxThreads=4)
    held = []
    for _ in range(3):                      # hold three streams open: the prefetcher must stall, not overrun
        assert it.hasNext()
        held.append(it.next()[2])
    import time
    time.sleep(0.2)
    assert it.statistics()["peakMemoryUsage"] <= budget
    assert it.statistics()["numStreams"] == 3
    for h in held:
        it.closeStream(h)                   # onClose(bufferSize) returns the budget (adaptor :49-58, iterator :96-100)
    it.closeStream(held[0])                 # double close is ignored
    n = 3 + len(drain(it))
    assert n == 48
    st = it.statistics()
    assert st["peakMemoryUsage"] <= budget and 1 <= st["peakThreads"] <= 4 and st["activeThreads"] >= 1
    it.close()


def test_block_larger_than_the_budget_is_buffered_up_to_the_budget(tmp_path):
    d = host.S3ShuffleDispatcher(new_conf(tmp_path))
    blocks = write_maps(d, [[100_000, 10]])
    it = host.S3BufferedPrefetchIterator(d, 0, [0], 0, 2, maxBufferSize=30_000)
    got = dict(drain(it))
    # bsize = min(maxBufferSize, maxBytes) (:125): the adaptor holds the first 30,000 bytes, the rest is read through
    assert got[(0, 0, 1)] == blocks[(0, 0)][:30_000] and got[(0, 1, 2)] == blocks[(0, 1)]
    it.close()


def test_abandoning_the_iterator_with_open_streams_does_not_hang(tmp_path):
    d = host.S3ShuffleDispatcher(new_conf(tmp_path))
    write_maps(d, [[5_000] * 20 for _ in range(10)])
    it = host.S3BufferedPrefetchIterator(d, 0, range(10), 0, 20, maxBufferSize=16_000, maxThreads=3)
    it.next()
    it.next()
    it.close()                              # task cancelled: threads are joined, buffers freed


def test_missing_data_object_surfaces_as_ioexception(tmp_path):
    import os
    d = host.S3ShuffleDispatcher(new_conf(tmp_path))
    write_maps(d, [[100, 100]])
    os.unlink(d.getPath("data", 0, 0))
    it = host.S3BufferedPrefetchIterator(d, 0, [0], 0, 2)
    with pytest.raises(host.IOException, match="File does not exist"):
        drain(it)
    it.close()


def test_empty_iterator(tmp_path):
    d = host.S3ShuffleDispatcher(new_conf(tmp_path))
    it = host.S3BufferedPrefetchIterator(d, 0, [], 0, 2)
    assert not it.hasNext()
    assert it.statistics()["line"] == "Unable to print statistics: / by zero."   # the reference's r == 0 case (:190)
    it.close()


def test_measure_output_stream_counters(tmp_path):
    d = host.S3ShuffleDispatcher(new_conf(tmp_path))
    w = host.S3ShuffleMapOutputWriter(d, 0, 3, 2)
    for r, n in enumerate((70_000, 5)):
        with w.getPartitionWriter(r) as s:
            s.write(b"x" * n)
    with pytest.raises(host.RuntimeException, match="no .data object"):
        w.statistics()
    w.commitAllPartitions(np.zeros(2, dtype=np.int64))
    nbytes, nanos, line = w.statistics()
    assert nbytes == 70_005 and nanos > 0
    assert re.match(r"Statistics: Stage 0\.0 TID 0 -- Writing shuffle_0_3_0\.data 70005 took \d+ ms \(.* MiB/s\)", line)
    w.close()
