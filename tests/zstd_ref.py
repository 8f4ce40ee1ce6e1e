"""libzstd.so.1 (1.5.5 — the library zstd-jni 1.5.5-x bundles) through ctypes: the reference Zstandard implementation
used to pin our decoder/encoder.  Test infrastructure only."""
import ctypes as C

_z = None


def lib():
    global _z
    if _z is None:
        z = C.CDLL("libzstd.so.1")
        z.ZSTD_compressBound.restype = C.c_size_t
        z.ZSTD_compressBound.argtypes = [C.c_size_t]
        z.ZSTD_compress.restype = C.c_size_t
        z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        z.ZSTD_decompress.restype = C.c_size_t
        z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_isError.restype = C.c_uint
        z.ZSTD_isError.argtypes = [C.c_size_t]
        z.ZSTD_createCCtx.restype = C.c_void_p
        z.ZSTD_freeCCtx.argtypes = [C.c_void_p]
        z.ZSTD_CCtx_setParameter.restype = C.c_size_t
        z.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        z.ZSTD_compressStream2.restype = C.c_size_t
        z.ZSTD_compressStream2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        z.ZSTD_createDCtx.restype = C.c_void_p
        z.ZSTD_freeDCtx.argtypes = [C.c_void_p]
        z.ZSTD_decompressStream.restype = C.c_size_t
        z.ZSTD_decompressStream.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _z = z
    return _z


class _Buf(C.Structure):
    _fields_ = [("p", C.c_void_p), ("size", C.c_size_t), ("pos", C.c_size_t)]


def compress(data, level=3):
    z = lib()
    cap = z.ZSTD_compressBound(len(data))
    buf = C.create_string_buffer(cap)
    n = z.ZSTD_compress(buf, cap, data, len(data), level)
    assert not z.ZSTD_isError(n)
    return buf.raw[:n]


def compress_stream(data, level=1, chunk=32768, flush_every=0):
    """What zstd-jni's ZstdOutputStreamNoFinalizer does under Spark's BufferedOutputStream(32 KiB): a streaming frame —
    no Frame_Content_Size, window descriptor present — fed in 32 KiB writes, ended with ZSTD_e_end."""
    z = lib()
    c = z.ZSTD_createCCtx()
    z.ZSTD_CCtx_setParameter(c, 100, level)  # ZSTD_c_compressionLevel
    out = b""
    ob = C.create_string_buffer(1 << 17)
    src = C.create_string_buffer(data, len(data)) if data else C.create_string_buffer(1)
    base = C.addressof(src)
    pos = 0
    k = 0
    while True:
        n = min(chunk, len(data) - pos)
        last = pos + n >= len(data)
        ib = _Buf(base + pos, n, 0)
        mode = 2 if last else (1 if flush_every and (k + 1) % flush_every == 0 else 0)  # end / flush / continue
        while True:
            o = _Buf(C.addressof(ob), len(ob), 0)
            rem = z.ZSTD_compressStream2(c, C.byref(o), C.byref(ib), mode)
            assert not z.ZSTD_isError(rem)
            out += ob.raw[:o.pos]
            if (mode == 0 and ib.pos == ib.size) or (mode != 0 and rem == 0):
                break
        pos += n
        k += 1
        if last:
            break
    z.ZSTD_freeCCtx(c)
    return out


def decompress(data, cap=None):
    """streaming decode of concatenated frames; raises IOError on malformed input"""
    z = lib()
    d = z.ZSTD_createDCtx()
    out = b""
    ob = C.create_string_buffer(1 << 17)
    src = C.create_string_buffer(data, len(data)) if data else C.create_string_buffer(1)
    ib = _Buf(C.addressof(src), len(data), 0)
    ret = 0
    try:
        while ib.pos < ib.size:
            o = _Buf(C.addressof(ob), len(ob), 0)
            ret = z.ZSTD_decompressStream(d, C.byref(o), C.byref(ib))
            if z.ZSTD_isError(ret):
                raise IOError("zstd: corrupt input")
            out += ob.raw[:o.pos]
        while ret != 0 and not z.ZSTD_isError(ret):  # drain
            o = _Buf(C.addressof(ob), len(ob), 0)
            r2 = z.ZSTD_decompressStream(d, C.byref(o), C.byref(ib))
            if z.ZSTD_isError(r2):
                raise IOError("zstd: corrupt input")
            out += ob.raw[:o.pos]
            if o.pos == 0:
                if r2 != 0:
                    raise IOError("zstd: truncated frame")
                break
            ret = r2
    finally:
        z.ZSTD_freeDCtx(d)
    return out
