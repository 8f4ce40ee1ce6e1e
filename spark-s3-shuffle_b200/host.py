"""ctypes binding of include/b200shuffle_host.h (libb200shuffle_host.so, the C++ host mirror of the reference's
plugin classes on the codec path).  Class and method names follow the reference so the tests read like
src/test/scala/org/apache/spark/shuffle/S3ShuffleManagerTest.scala; nothing here computes.
"""
import ctypes as C

import numpy as np

from . import _build

OK, E_RUNTIME, E_IO, E_SPARK, E_UNSUPPORTED, E_CODEC = 0, -101, -102, -103, -104, -105


class RuntimeException(RuntimeError):
    """java.lang.RuntimeException (preconditions, length mismatch: shuffle/S3ShuffleMapOutputWriter.scala:68-73,96-100)"""


class IOException(IOError):
    """java.io.IOException (closed stream :175-177; "Stream is corrupted")"""


class SparkException(Exception):
    """org.apache.spark.SparkException (storage/S3ChecksumValidationStream.scala:72-74, helper/S3ShuffleHelper.scala:112-114)"""


class UnsupportedOperationException(Exception):
    """java.lang.UnsupportedOperationException (helper/S3ShuffleHelper.scala:100-101)"""


class CodecException(RuntimeError):
    """The C ABI reported a call-level failure (no GPU / CUDA error) — there is no CPU fallback."""


_EXC = {E_RUNTIME: RuntimeException, E_IO: IOException, E_SPARK: SparkException,
        E_UNSUPPORTED: UnsupportedOperationException, E_CODEC: CodecException}

_vp, _i32, _i64, _u32, _u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
PROTOTYPES = [
    ("b2sh_last_error", C.c_char_p, []),
    ("b2sh_dispatcher_create", C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    ("b2sh_dispatcher_destroy", None, [_vp]),
    ("b2sh_dispatcher_get_path", C.c_int, [_vp, C.c_int, _i32, _i64, C.c_char_p, _u32]),
    ("b2sh_dispatcher_remove_shuffle", C.c_int, [_vp, _i32]),
    ("b2sh_helper_checksum_algorithm", C.c_int, [C.c_char_p]),
    ("b2sh_helper_get_partition_lengths", C.c_int, [_vp, _i32, _i64, _vp, _u32, C.POINTER(_u32)]),
    ("b2sh_helper_get_checksums", C.c_int, [_vp, _i32, _i64, _vp, _u32, C.POINTER(_u32)]),
    ("b2sh_writer_create", C.c_int, [_vp, _i32, _i64, _i32, C.POINTER(_vp)]),
    ("b2sh_writer_open_partition", C.c_int, [_vp, _i32]),
    ("b2sh_writer_write", C.c_int, [_vp, _vp, _u64]),
    ("b2sh_writer_close_partition", C.c_int, [_vp]),
    ("b2sh_writer_commit_all_partitions", C.c_int, [_vp, _vp, _vp]),
    ("b2sh_writer_abort", C.c_int, [_vp]),
    ("b2sh_writer_destroy", None, [_vp]),
    ("b2sh_single_spill_transfer", C.c_int, [_vp, _i32, _i64, C.c_char_p, _vp, _vp, _u32, C.c_int]),
    ("b2sh_reader_create", C.c_int, [_vp, _i32, _vp, _u32, _i32, _i32, C.c_int, C.POINTER(_vp)]),
    ("b2sh_reader_read", C.c_int, [_vp, C.POINTER(_u32)]),
    ("b2sh_reader_block", C.c_int,
     [_vp, _u32, C.POINTER(_i64), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_vp), C.POINTER(_u64)]),
    ("b2sh_reader_remote_bytes_read", _u64, [_vp]),
    ("b2sh_reader_destroy", None, [_vp]),
]
SYMBOLS = [p[0] for p in PROTOTYPES]
_lib = None


def load():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build_host())
        for name, res, args in PROTOTYPES:
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _check(rc):
    if rc < 0:
        raise _EXC.get(rc, RuntimeException)(load().b2sh_last_error().decode())
    return rc


class S3ShuffleDispatcher:
    """helper/S3ShuffleDispatcher.scala — conf is a dict of the reference's SparkConf keys."""

    def __init__(self, conf):
        text = "".join("%s=%s\n" % (k, str(v).lower() if isinstance(v, bool) else v) for k, v in conf.items())
        self._h = _vp()
        _check(load().b2sh_dispatcher_create(text.encode(), C.byref(self._h)))

    def getPath(self, kind, shuffleId, mapId):
        """kind: 'data' | 'index' | 'checksum' (helper/S3ShuffleDispatcher.scala:142-143)"""
        buf = C.create_string_buffer(4096)
        _check(load().b2sh_dispatcher_get_path(self._h, {"data": 0, "index": 1, "checksum": 2}[kind], shuffleId, mapId,
                                               buf, 4096))
        return buf.value.decode()

    def removeShuffle(self, shuffleId):
        _check(load().b2sh_dispatcher_remove_shuffle(self._h, shuffleId))

    def close(self):
        if self._h:
            load().b2sh_dispatcher_destroy(self._h)
            self._h = None


class S3ShuffleHelper:
    """helper/S3ShuffleHelper.scala"""

    @staticmethod
    def createChecksumAlgorithm(name):
        return _check(load().b2sh_helper_checksum_algorithm(name.encode()))

    @staticmethod
    def _array(fn, d, shuffleId, mapId):
        cap = 1 << 16
        while True:
            out = np.zeros(cap, dtype=np.int64)
            cnt = _u32(0)
            rc = fn(d._h, shuffleId, mapId, out.ctypes.data, cap, C.byref(cnt))
            if rc == E_RUNTIME and cnt.value > cap:
                cap = cnt.value
                continue
            _check(rc)
            return out[: cnt.value].copy()

    @staticmethod
    def getPartitionLengths(d, shuffleId, mapId):
        """the cumulative offsets stored in .index (:67-81)"""
        return S3ShuffleHelper._array(load().b2sh_helper_get_partition_lengths, d, shuffleId, mapId)

    @staticmethod
    def getChecksums(d, shuffleId, mapId):
        return S3ShuffleHelper._array(load().b2sh_helper_get_checksums, d, shuffleId, mapId)


class S3ShuffleMapOutputWriter:
    """shuffle/S3ShuffleMapOutputWriter.scala — getPartitionWriter(p) returns a stream-like object."""

    class _PartitionStream:
        def __init__(self, w):
            self._w = w

        def write(self, data):
            a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
            _check(load().b2sh_writer_write(self._w._h, a.ctypes.data if a.size else None, a.size))

        def close(self):
            _check(load().b2sh_writer_close_partition(self._w._h))

        def __enter__(self):
            return self

        def __exit__(self, *a):
            self.close()

    def __init__(self, dispatcher, shuffleId, mapId, numPartitions):
        self._h = _vp()
        self.numPartitions = numPartitions
        _check(load().b2sh_writer_create(dispatcher._h, shuffleId, mapId, numPartitions, C.byref(self._h)))

    def getPartitionWriter(self, reducePartitionId):
        _check(load().b2sh_writer_open_partition(self._h, reducePartitionId))
        return self._PartitionStream(self)

    def commitAllPartitions(self, checksums=None):
        """-> partitionLengths (MapOutputCommitMessage).  checksums are only consumed in pass-through mode."""
        out = np.zeros(max(self.numPartitions, 1), dtype=np.int64)
        ck = None
        if checksums is not None:
            ck = np.ascontiguousarray(checksums, dtype=np.int64)
            assert ck.size == self.numPartitions
        _check(load().b2sh_writer_commit_all_partitions(self._h, ck.ctypes.data if ck is not None else None,
                                                        out.ctypes.data))
        return out[: self.numPartitions]

    def abort(self):
        _check(load().b2sh_writer_abort(self._h))

    def close(self):
        if self._h:
            load().b2sh_writer_destroy(self._h)
            self._h = None


class S3SingleSpillShuffleMapOutputWriter:
    """shuffle/S3SingleSpillShuffleMapOutputWriter.scala"""

    def __init__(self, dispatcher, shuffleId, mapId):
        self._d, self.shuffleId, self.mapId = dispatcher, shuffleId, mapId

    def transferMapSpillFile(self, mapSpillFile, partitionLengths, checksums, verifyOnTransfer=False):
        pl = np.ascontiguousarray(partitionLengths, dtype=np.int64)
        ck = np.ascontiguousarray(checksums, dtype=np.int64)
        assert pl.size == ck.size
        _check(load().b2sh_single_spill_transfer(self._d._h, self.shuffleId, self.mapId, str(mapSpillFile).encode(),
                                                 pl.ctypes.data, ck.ctypes.data, pl.size, int(verifyOnTransfer)))


class S3ShuffleReader:
    """storage/S3ShuffleReader.scala — read() yields (blockName-parts, decoded bytes) per non-empty block."""

    def __init__(self, dispatcher, shuffleId, mapIds, startPartition, endPartition, doBatchFetch=False):
        self._h = _vp()
        ids = np.ascontiguousarray(mapIds, dtype=np.int64)
        _check(load().b2sh_reader_create(dispatcher._h, shuffleId, ids.ctypes.data, ids.size, startPartition,
                                         endPartition, int(doBatchFetch), C.byref(self._h)))

    def read(self):
        n = _u32(0)
        _check(load().b2sh_reader_read(self._h, C.byref(n)))
        out = []
        for k in range(n.value):
            m, rs, re, p, ln = _i64(), _i32(), _i32(), _vp(), _u64()
            _check(load().b2sh_reader_block(self._h, k, C.byref(m), C.byref(rs), C.byref(re), C.byref(p), C.byref(ln)))
            data = C.string_at(p.value, ln.value) if ln.value else b""
            out.append(((m.value, rs.value, re.value), data))
        return out

    @property
    def remoteBytesRead(self):
        return load().b2sh_reader_remote_bytes_read(self._h)

    def close(self):
        if self._h:
            load().b2sh_reader_destroy(self._h)
            self._h = None
