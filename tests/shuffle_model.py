"""Shared helpers for the host-mirror tests: a Kryo-like record model and an ORACLE-side reader of the reference's
on-disk layout (used to check files the product wrote; test infrastructure only)."""
import os

import numpy as np


def encode_pairs(keys, values):
    """(Int, Int) records as zig-zag varint pairs — the byte model of Kryo's default Int serializer
    (SURVEY.md §8d config 1).  Vectorised; returns bytes."""
    k = np.asarray(keys, dtype=np.int64)
    v = np.asarray(values, dtype=np.int64)
    inter = np.empty(k.size * 2, dtype=np.int64)
    inter[0::2], inter[1::2] = k, v
    z = ((inter << 1) ^ (inter >> 63)).astype(np.uint64) & np.uint64(0xFFFFFFFF)
    nbytes = np.ones(z.size, dtype=np.int64)
    for t in (7, 14, 21, 28):
        nbytes += (z >= (1 << t)).astype(np.int64)
    off = np.concatenate(([0], np.cumsum(nbytes)))
    out = np.zeros(int(off[-1]), dtype=np.uint8)
    for j in range(5):
        m = nbytes > j
        byte = ((z[m] >> np.uint64(7 * j)) & np.uint64(0x7F)).astype(np.uint8)
        cont = (nbytes[m] > j + 1).astype(np.uint8) << 7
        out[off[:-1][m] + j] = byte | cont
    return out.tobytes()


def decode_pairs(data):
    b = np.frombuffer(data, dtype=np.uint8)
    if b.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    ends = np.flatnonzero((b & 0x80) == 0)
    starts = np.concatenate(([0], ends[:-1] + 1))
    z = np.zeros(ends.size, dtype=np.uint64)
    for j in range(5):
        idx = starts + j
        m = idx <= ends
        z[m] |= (b[idx[m]].astype(np.uint64) & np.uint64(0x7F)) << np.uint64(7 * j)
    x = (z >> np.uint64(1)).astype(np.int64) ^ -(z & np.uint64(1)).astype(np.int64)
    return x[0::2], x[1::2]


def oracle_read_partition(oracle, dispatcher, shuffle_id, map_ids, reduce_id, alg_name, codec="lz4"):
    """What the unmodified reference's reduce side does, with the oracle's arithmetic: resolve the byte range from
    .index (storage/S3ShuffleBlockIterator.scala:36-43), verify the slice against .checksum
    (storage/S3ChecksumValidationStream.scala:54-86), decompress (storage/S3ShuffleReader.scala:107-109)."""
    out = []
    for m in map_ids:
        ipath = dispatcher.getPath("index", shuffle_id, m)
        if not os.path.exists(ipath):
            continue
        acc = oracle.read_be64(open(ipath, "rb").read())
        a, b = int(acc[reduce_id]), int(acc[reduce_id + 1])
        if a == b:
            continue
        data = open(dispatcher.getPath("data", shuffle_id, m), "rb").read()
        if alg_name:
            ref = oracle.read_be64(open(dispatcher.getPath("checksum", shuffle_id, m), "rb").read())
            alg = {"ADLER32": 1, "CRC32": 2, "CRC32C": 3}[alg_name]
            bad = oracle.validate_slices(alg, data[a:b], acc, ref, reduce_id, reduce_id + 1)
            assert bad == -1, "oracle: invalid checksum for shuffle_%d_%d_%d" % (shuffle_id, m, bad)
        dec = oracle.lz4block_decompress(data[a:b]) if codec == "lz4" else oracle.xerial_decompress(data[a:b])
        out.append((m, dec))
    return out
