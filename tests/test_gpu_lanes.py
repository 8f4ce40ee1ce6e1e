"""Round-2 runtime behaviours of the C ABI on a real device: the two lanes per device (a compress and a decompress call
from different threads in flight together), NUMA placement entry points, and per-block (not per-call) failure for a
block whose slices do not tile it — the reference surfaces that as "Invalid checksum detected" for that block only
(storage/S3ChecksumValidationStream.scala:72-74; storage/S3ShuffleBlockStream.scala:66-69 turns I/O errors into EOF)."""
import threading

import numpy as np
import pytest

from conftest import corpus

pytestmark = pytest.mark.gpu


def test_short_block_fails_alone_with_the_slice_index(capi, oracle):
    parts = [corpus(oracle, "terasort", n, seed=40 + i) for i, n in enumerate([50000, 30000, 9000])]
    streams = [oracle.lz4block_compress(p) for p in parts]
    alg = capi.CHECKSUM_CRC32C
    sums = [oracle.crc32c(s) for s in streams]
    batch = b"".join(streams)                       # one block covering three reduce partitions
    truncated = batch[: len(streams[0]) + 100]      # the fetch came back short, inside the second partition
    slices = [(len(s), c) for s, c in zip(streams, sums)]
    out, st, bad = capi.decompress_batch(capi.CODEC_LZ4BLOCK, [batch, truncated, streams[2]], alg,
                                         [slices, slices, [slices[2]]],
                                         dst_caps=[sum(map(len, parts)), sum(map(len, parts)), len(parts[2])])
    assert st == [0, capi.E_CHECKSUM, 0] and bad[1] == 1
    assert out[0] == b"".join(parts) and out[1] is None and out[2] == parts[2]
    # packed form: the failed block takes no room in the destination arena
    src = np.frombuffer(batch + truncated + streams[2], dtype=np.uint8)
    off = [0, len(batch), len(batch) + len(truncated)]
    ln = [len(batch), len(truncated), len(streams[2])]
    dst = np.zeros(2 * sum(map(len, parts)), dtype=np.uint8)
    r = capi.decompress_packed(capi.CODEC_LZ4BLOCK, src, off, ln, dst, alg, [0, 3, 6, 7],
                               [s[0] for s in slices] * 2 + [slices[2][0]], [s[1] for s in slices] * 2 + [slices[2][1]])
    assert list(r["status"]) == [0, capi.E_CHECKSUM, 0] and r["bad_slice"][1] == 1
    assert r["dst_len"][1] == 0 and r["total"] == len(b"".join(parts)) + len(parts[2])
    o2 = int(r["dst_off"][2])
    assert dst[o2:o2 + len(parts[2])].tobytes() == parts[2]


def test_write_and_read_lanes_run_concurrently_and_stay_correct(capi, oracle):
    """a map-side compress call and a reduce-side decompress call from two threads (the C ABI gives each direction its
    own lane: lock, slots, streams) — results must be what each call produces on its own"""
    parts = [corpus(oracle, "terasort", 300000 + 1000 * i, seed=60 + i) for i in range(24)]
    want_streams = [oracle.lz4block_compress(p, 32768, compressor=1) for p in parts]
    sums = [oracle.crc32c(s) for s in want_streams]
    errors = []

    def writer():
        try:
            for _ in range(6):
                comp, cks, st = capi.compress_batch(capi.CODEC_LZ4BLOCK, parts, 32768, capi.CHECKSUM_CRC32C)
                assert st == [0] * len(parts) and comp == want_streams and cks == sums
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    def reader():
        try:
            for _ in range(6):
                out, st, _ = capi.decompress_batch(capi.CODEC_LZ4BLOCK, want_streams, capi.CHECKSUM_CRC32C,
                                                   [[(len(s), c)] for s, c in zip(want_streams, sums)],
                                                   dst_caps=[len(p) for p in parts])
                assert st == [0] * len(parts) and out == parts
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ths = [threading.Thread(target=writer), threading.Thread(target=reader), threading.Thread(target=reader)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors


def test_numa_entry_points(capi):
    L = capi.load()
    node = capi.bind_thread_to_device(0)          # never an error: a no-op when the topology is unknown
    assert node == -1000 or node >= 0
    assert L.b2s_bind_thread_to_device(99) == capi.E_ARG
    hb = capi.HostBuffer(1 << 20)                 # allocated under the device's node preference
    hb.array[:] = 7
    assert int(hb.array.sum()) == 7 << 20
    hb.free()
