"""ctypes binding of include/b200shuffle.h — the same entry points the JNI shim binds.

Nothing here computes: every helper marshals numpy/bytes into plain pointers + sizes and calls libb200shuffle.so.
"""
import ctypes as C
import os

import numpy as np

from . import _build

# ---- constants mirrored from b200shuffle.h ----
CODEC_NONE, CODEC_LZ4BLOCK, CODEC_SNAPPY_XERIAL, CODEC_ZSTD = 0, 1, 2, 3
CHECKSUM_NONE, CHECKSUM_ADLER32, CHECKSUM_CRC32, CHECKSUM_CRC32C = 0, 1, 2, 3
OK, E_CORRUPT, E_CHECKSUM, E_DST_TOO_SMALL, E_UNSUPPORTED, E_ARG, E_CUDA, E_NOT_INIT, E_NOMEM = (
    0, -1, -2, -3, -4, -5, -6, -7, -8)
CODEC_BY_NAME = {"lz4": CODEC_LZ4BLOCK, "snappy": CODEC_SNAPPY_XERIAL, "zstd": CODEC_ZSTD}
CHECKSUM_BY_NAME = {"ADLER32": CHECKSUM_ADLER32, "CRC32": CHECKSUM_CRC32, "CRC32C": CHECKSUM_CRC32C}

_u8p, _u64p, _u32p, _i32p, _vp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
_u32, _u64, _i32 = C.c_uint32, C.c_uint64, C.c_int32


class Timing(C.Structure):
    _fields_ = [
        ("total_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("kernel_ms", C.c_double),
        ("top_kernel_ms", C.c_double), ("h2d_bytes", _u64), ("d2h_bytes", _u64), ("kernel_launches", _u64),
        ("src_bytes", _u64), ("dst_bytes", _u64), ("dominant_ms", C.c_double), ("dominant_launches", _u64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/b200shuffle.h declares: (name, restype, argtypes)
PROTOTYPES = [
    ("b2s_init", C.c_int, [_u32, _u64, _u32]),
    ("b2s_shutdown", None, []),
    ("b2s_device_count", C.c_int, []),
    ("b2s_set_thread_device", C.c_int, [_u32]),
    ("b2s_bind_thread_to_device", C.c_int, [_u32]),
    ("b2s_device_numa_node", C.c_int, [_u32]),
    ("b2s_strerror", C.c_char_p, [_i32]),
    ("b2s_last_error", C.c_char_p, []),
    ("b2s_version", _u32, []),
    ("b2s_host_alloc", C.c_void_p, [_u64]),
    ("b2s_host_free", None, [C.c_void_p]),
    ("b2s_host_register", C.c_int, [C.c_void_p, _u64]),
    ("b2s_host_unregister", C.c_int, [C.c_void_p]),
    ("b2s_compress_bound", _u64, [_u32, _u32, _u64]),
    ("b2s_decompressed_size_batch", C.c_int, [_u32, _u32, _vp, _u64p, _u64p, _i32p]),
    ("b2s_checksum_batch", C.c_int, [_u32, _u32, _vp, _u64p, _u64p]),
    ("b2s_checksum_packed", C.c_int, [_u32, _u32, _u8p, _u64p, _u64p, _u64p]),
    ("b2s_compress_batch", C.c_int, [_u32, _i32, _u32, _u32, _u32, _vp, _u64p, _vp, _u64p, _u64p, _u64p, _i32p]),
    ("b2s_compress_packed", C.c_int,
     [_u32, _i32, _u32, _u32, _u32, _u8p, _u64p, _u64p, _u8p, _u64, _u64p, _u64p, _u64p, _u64p, _i32p]),
    ("b2s_decompress_batch", C.c_int,
     [_u32, _u32, _u32, _vp, _u64p, _u32p, _vp, _vp, _vp, _u64p, _u64p, _i32p, _i32p]),
    ("b2s_decompress_packed", C.c_int,
     [_u32, _u32, _u32, _u8p, _u64p, _u64p, _u32p, _u64p, _u64p, _u8p, _u64, _u64p, _u64p, _u64p, _i32p, _i32p]),
    ("b2s_checksum_dev", C.c_int, [_u32, _u32, _u32, _vp, _u64p, _u64p, _u64p]),
    ("b2s_compress_dev", C.c_int,
     [_u32, _u32, _i32, _u32, _u32, _u32, _vp, _u64p, _u64p, _vp, _u64, _u64p, _u64p, _u64p, _u64p, _i32p]),
    ("b2s_decompress_dev", C.c_int,
     [_u32, _u32, _u32, _u32, _vp, _u64p, _u64p, _u32p, _u64p, _u64p, _vp, _u64, _u64p, _u64p, _u64p, _i32p, _i32p]),
    ("b2s_dev_alloc", C.c_void_p, [_u32, _u64]),
    ("b2s_dev_free", None, [_u32, C.c_void_p]),
    ("b2s_dev_memcpy", C.c_int, [_u32, C.c_void_p, C.c_void_p, _u64, C.c_int]),
    ("b2s_last_timing", C.c_int, [C.POINTER(Timing)]),
    ("b2s_total_kernel_launches", _u64, []),
    ("b2s_mark", C.c_int, [_u32, _u32]),
    ("b2s_marks_elapsed_ms", C.c_int, [_u32, C.POINTER(C.c_double)]),
    ("b2s_gen_terasort_dev", C.c_int, [_u32, C.c_void_p, _u64, _u64, _u64]),
]
SYMBOLS = [p[0] for p in PROTOTYPES]

_lib = None


def load(build_if_needed=True):
    """Loads libb200shuffle.so (building it with nvcc when stale).  Raises if that is impossible — no fallback."""
    global _lib
    if _lib is None:
        path = _build.build() if build_if_needed else _build.LIB
        if not os.path.exists(path):
            raise RuntimeError("libb200shuffle.so is missing and could not be built; there is no CPU fallback")
        L = C.CDLL(path)
        for name, res, args in PROTOTYPES:
            f = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class B2SError(RuntimeError):
    def __init__(self, code, where):
        L = load()
        self.code = code
        msg = L.b2s_strerror(code).decode()
        detail = L.b2s_last_error().decode()
        super().__init__("%s: %s (%d)%s" % (where, msg, code, (": " + detail) if detail else ""))


def _check(rc, where):
    if rc != 0:
        raise B2SError(rc, where)


def init(gpu_mask=0, pinned_bytes_per_gpu=0, streams_per_gpu=0):
    _check(load().b2s_init(gpu_mask, pinned_bytes_per_gpu, streams_per_gpu), "b2s_init")


def bind_thread_to_device(dev=0):
    """pins the calling thread to the CPUs of the device's NUMA node and prefers that node for its allocations"""
    _check(load().b2s_bind_thread_to_device(dev), "b2s_bind_thread_to_device")
    return load().b2s_device_numa_node(dev)


def shutdown():
    load().b2s_shutdown()


def last_timing():
    t = Timing()
    load().b2s_last_timing(C.byref(t))
    return t.as_dict()


def mark(which, dev=0):
    _check(load().b2s_mark(dev, which), "b2s_mark")


def marks_elapsed_ms(dev=0):
    ms = C.c_double(0)
    _check(load().b2s_marks_elapsed_ms(dev, C.byref(ms)), "b2s_marks_elapsed_ms")
    return ms.value


def compress_bound(codec, block_size, n):
    return load().b2s_compress_bound(codec, block_size, n)


# ---------------------------------------------------------------------------------------------------------------
# marshalling helpers
# ---------------------------------------------------------------------------------------------------------------
def _as_u8(b):
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(b, dtype=np.uint8)


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


def _ptr_array(arrs):
    out = np.zeros(max(len(arrs), 1), dtype=np.uint64)
    for i, a in enumerate(arrs):
        out[i] = a.ctypes.data if a.size else 0
    return out


class HostBuffer:
    """Pinned host memory from b2s_host_alloc, viewed as a numpy uint8 array (what the JVM wraps as a direct buffer)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = load().b2s_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise B2SError(E_NOMEM, "b2s_host_alloc")
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.uint8, count=self.nbytes)

    def free(self):
        if self.ptr:
            self.array = None
            load().b2s_host_free(self.ptr)
            self.ptr = None


# ---------------------------------------------------------------------------------------------------------------
# host-pointer API
# ---------------------------------------------------------------------------------------------------------------
def checksum_batch(alg, blocks):
    arrs = [_as_u8(b) for b in blocks]
    n = len(arrs)
    ptrs = _ptr_array(arrs)
    lens = np.array([a.size for a in arrs] or [0], dtype=np.uint64)
    out = np.zeros(max(n, 1), dtype=np.uint64)
    _check(load().b2s_checksum_batch(alg, n, _ptr(ptrs), _ptr(lens), _ptr(out)), "b2s_checksum_batch")
    return [int(v) for v in out[:n]]


def checksum_packed(alg, base, off, length):
    base = _as_u8(base)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    length = np.ascontiguousarray(length, dtype=np.uint64)
    out = np.zeros(max(off.size, 1), dtype=np.uint64)
    _check(load().b2s_checksum_packed(alg, off.size, _ptr(base), _ptr(off), _ptr(length), _ptr(out)),
           "b2s_checksum_packed")
    return out[: off.size]


def compress_batch(codec, blocks, block_size=0, checksum_alg=CHECKSUM_NONE, level=0, dst_caps=None):
    """-> (list of compressed streams (bytes or None on error), checksums, status)"""
    arrs = [_as_u8(b) for b in blocks]
    n = len(arrs)
    lens = np.array([a.size for a in arrs] or [0], dtype=np.uint64)
    caps = np.array(
        (dst_caps if dst_caps is not None else [compress_bound(codec, block_size, a.size) for a in arrs]) or [0],
        dtype=np.uint64)
    outs = [np.empty(int(c), dtype=np.uint8) for c in caps[:n]]
    sp, dp = _ptr_array(arrs), _ptr_array(outs)
    dlen = np.zeros(max(n, 1), dtype=np.uint64)
    cks = np.zeros(max(n, 1), dtype=np.uint64)
    st = np.zeros(max(n, 1), dtype=np.int32)
    _check(load().b2s_compress_batch(codec, level, block_size, checksum_alg, n, _ptr(sp), _ptr(lens), _ptr(dp),
                                     _ptr(caps), _ptr(dlen), _ptr(cks), _ptr(st)), "b2s_compress_batch")
    res = [outs[i][: int(dlen[i])].tobytes() if st[i] == 0 else None for i in range(n)]
    return res, [int(v) for v in cks[:n]], [int(v) for v in st[:n]]


def compress_packed(codec, src, off, length, dst, block_size=0, checksum_alg=CHECKSUM_NONE, level=0):
    """src/dst: uint8 arrays (ideally HostBuffer.array).  -> dict(dst_off, dst_len, total, checksums, status)"""
    src, dst = _as_u8(src), _as_u8(dst)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    length = np.ascontiguousarray(length, dtype=np.uint64)
    n = off.size
    dst_off = np.zeros(max(n, 1), dtype=np.uint64)
    dst_len = np.zeros(max(n, 1), dtype=np.uint64)
    cks = np.zeros(max(n, 1), dtype=np.uint64)
    st = np.zeros(max(n, 1), dtype=np.int32)
    total = _u64(0)
    _check(load().b2s_compress_packed(codec, level, block_size, checksum_alg, n, _ptr(src), _ptr(off), _ptr(length),
                                      _ptr(dst), dst.size, _ptr(dst_off), _ptr(dst_len), C.addressof(total),
                                      _ptr(cks), _ptr(st)), "b2s_compress_packed")
    return dict(dst_off=dst_off[:n], dst_len=dst_len[:n], total=total.value, checksums=cks[:n], status=st[:n])


def decompressed_size_batch(codec, blocks):
    arrs = [_as_u8(b) for b in blocks]
    n = len(arrs)
    sp = _ptr_array(arrs)
    lens = np.array([a.size for a in arrs] or [0], dtype=np.uint64)
    out = np.zeros(max(n, 1), dtype=np.uint64)
    st = np.zeros(max(n, 1), dtype=np.int32)
    _check(load().b2s_decompressed_size_batch(codec, n, _ptr(sp), _ptr(lens), _ptr(out), _ptr(st)),
           "b2s_decompressed_size_batch")
    return [int(v) for v in out[:n]], [int(v) for v in st[:n]]


def decompress_batch(codec, blocks, checksum_alg=CHECKSUM_NONE, slices=None, dst_caps=None):
    """slices: per block a list of (length, checksum) pairs (the block's .index differences / .checksum values).
    -> (list of decoded bytes or None, status, bad_slice)"""
    arrs = [_as_u8(b) for b in blocks]
    n = len(arrs)
    if dst_caps is None:
        sizes, _ = decompressed_size_batch(codec, blocks)
        dst_caps = sizes
    caps = np.array(list(dst_caps) or [0], dtype=np.uint64)
    outs = [np.empty(max(int(c), 1), dtype=np.uint8) for c in caps[:n]]
    sp, dp = _ptr_array(arrs), _ptr_array(outs)
    lens = np.array([a.size for a in arrs] or [0], dtype=np.uint64)
    dlen = np.zeros(max(n, 1), dtype=np.uint64)
    st = np.zeros(max(n, 1), dtype=np.int32)
    bad = np.zeros(max(n, 1), dtype=np.int32)
    ns = sl_ptrs = sc_ptrs = None
    keep = []
    if checksum_alg != CHECKSUM_NONE:
        ns = np.array([len(s) for s in slices] or [0], dtype=np.uint32)
        sl = [np.array([p[0] for p in s] or [0], dtype=np.uint64) for s in slices]
        sc = [np.array([p[1] for p in s] or [0], dtype=np.uint64) for s in slices]
        keep = [sl, sc]
        sl_ptrs = np.array([a.ctypes.data for a in sl] or [0], dtype=np.uint64)
        sc_ptrs = np.array([a.ctypes.data for a in sc] or [0], dtype=np.uint64)
    _check(load().b2s_decompress_batch(codec, checksum_alg, n, _ptr(sp), _ptr(lens), _ptr(ns), _ptr(sl_ptrs),
                                       _ptr(sc_ptrs), _ptr(dp), _ptr(caps), _ptr(dlen), _ptr(st), _ptr(bad)),
           "b2s_decompress_batch")
    del keep
    res = [outs[i][: int(dlen[i])].tobytes() if st[i] == 0 else None for i in range(n)]
    return res, [int(v) for v in st[:n]], [int(v) for v in bad[:n]]


def decompress_packed(codec, src, off, length, dst, checksum_alg=CHECKSUM_NONE, slice_base=None, slice_len=None,
                      slice_checksum=None):
    src, dst = _as_u8(src), _as_u8(dst)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    length = np.ascontiguousarray(length, dtype=np.uint64)
    n = off.size
    sb = sl = sc = None
    if checksum_alg != CHECKSUM_NONE:
        sb = np.ascontiguousarray(slice_base, dtype=np.uint32)
        sl = np.ascontiguousarray(slice_len, dtype=np.uint64)
        sc = np.ascontiguousarray(slice_checksum, dtype=np.uint64)
    dst_off = np.zeros(max(n, 1), dtype=np.uint64)
    dst_len = np.zeros(max(n, 1), dtype=np.uint64)
    st = np.zeros(max(n, 1), dtype=np.int32)
    bad = np.zeros(max(n, 1), dtype=np.int32)
    total = _u64(0)
    _check(load().b2s_decompress_packed(codec, checksum_alg, n, _ptr(src), _ptr(off), _ptr(length), _ptr(sb), _ptr(sl),
                                        _ptr(sc), _ptr(dst), dst.size, _ptr(dst_off), _ptr(dst_len),
                                        C.addressof(total), _ptr(st), _ptr(bad)), "b2s_decompress_packed")
    return dict(dst_off=dst_off[:n], dst_len=dst_len[:n], total=total.value, status=st[:n], bad_slice=bad[:n])


# ---------------------------------------------------------------------------------------------------------------
# device-resident API (bench.py roofline leg); d_* are raw device addresses (ints)
# ---------------------------------------------------------------------------------------------------------------
def dev_alloc(nbytes, dev=0):
    p = load().b2s_dev_alloc(dev, nbytes)
    if not p:
        raise B2SError(E_NOMEM, "b2s_dev_alloc")
    return p


def dev_free(p, dev=0):
    load().b2s_dev_free(dev, p)


def dev_memcpy(dst, src, nbytes, kind, dev=0):
    _check(load().b2s_dev_memcpy(dev, dst, src, nbytes, kind), "b2s_dev_memcpy")


def gen_terasort_dev(d_dst, first_record, n_records, seed=42, dev=0):
    _check(load().b2s_gen_terasort_dev(dev, d_dst, first_record, n_records, seed), "b2s_gen_terasort_dev")


def checksum_dev(alg, d_base, off, length, dev=0):
    off = np.ascontiguousarray(off, dtype=np.uint64)
    length = np.ascontiguousarray(length, dtype=np.uint64)
    out = np.zeros(max(off.size, 1), dtype=np.uint64)
    _check(load().b2s_checksum_dev(dev, alg, off.size, d_base, _ptr(off), _ptr(length), _ptr(out)), "b2s_checksum_dev")
    return out[: off.size]


def compress_dev(codec, d_src, off, length, d_dst, dst_cap, block_size=0, checksum_alg=CHECKSUM_NONE, level=0, dev=0):
    off = np.ascontiguousarray(off, dtype=np.uint64)
    length = np.ascontiguousarray(length, dtype=np.uint64)
    n = off.size
    dst_off = np.zeros(max(n, 1), dtype=np.uint64)
    dst_len = np.zeros(max(n, 1), dtype=np.uint64)
    cks = np.zeros(max(n, 1), dtype=np.uint64)
    st = np.zeros(max(n, 1), dtype=np.int32)
    total = _u64(0)
    _check(load().b2s_compress_dev(dev, codec, level, block_size, checksum_alg, n, d_src, _ptr(off), _ptr(length),
                                   d_dst, dst_cap, _ptr(dst_off), _ptr(dst_len), C.addressof(total), _ptr(cks),
                                   _ptr(st)), "b2s_compress_dev")
    return dict(dst_off=dst_off[:n], dst_len=dst_len[:n], total=total.value, checksums=cks[:n], status=st[:n])


def decompress_dev(codec, d_src, off, length, d_dst, dst_cap, checksum_alg=CHECKSUM_NONE, slice_base=None,
                   slice_len=None, slice_checksum=None, dev=0):
    off = np.ascontiguousarray(off, dtype=np.uint64)
    length = np.ascontiguousarray(length, dtype=np.uint64)
    n = off.size
    sb = sl = sc = None
    if checksum_alg != CHECKSUM_NONE:
        sb = np.ascontiguousarray(slice_base, dtype=np.uint32)
        sl = np.ascontiguousarray(slice_len, dtype=np.uint64)
        sc = np.ascontiguousarray(slice_checksum, dtype=np.uint64)
    dst_off = np.zeros(max(n, 1), dtype=np.uint64)
    dst_len = np.zeros(max(n, 1), dtype=np.uint64)
    st = np.zeros(max(n, 1), dtype=np.int32)
    bad = np.zeros(max(n, 1), dtype=np.int32)
    total = _u64(0)
    _check(load().b2s_decompress_dev(dev, codec, checksum_alg, n, d_src, _ptr(off), _ptr(length), _ptr(sb), _ptr(sl),
                                     _ptr(sc), d_dst, dst_cap, _ptr(dst_off), _ptr(dst_len), C.addressof(total),
                                     _ptr(st), _ptr(bad)), "b2s_decompress_dev")
    return dict(dst_off=dst_off[:n], dst_len=dst_len[:n], total=total.value, status=st[:n], bad_slice=bad[:n])
