"""CPU check of spark-s3-shuffle_b200/csrc/lz4_parse_core.h — the lane-local walk of the sub-chunk parallel parse
(B2S_LZ4_PIPE=4) — against the oracle's executable specification of that generation (orc_lz4_compress_block_win_sub,
orc_snappy_compress_raw_win_sub via xerial framing, compressor=2).

The header is compiled by g++ into tests/native/lz4_parse_host.cpp together with a plain restatement of what the match
kernel hands it (off[] + the "exactly 4" flag + the window masks) and of the kernel's stitch; the records are turned into
LZ4 block bytes here and must equal the oracle's bytes.  No GPU involved; tests/test_gpu_pipe4.py then demands the same
bytes from the kernels themselves.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "lz4_parse_host.cpp")
HDR = os.path.join(ROOT, "spark-s3-shuffle_b200", "csrc", "lz4_parse_core.h")


@pytest.fixture(scope="module")
def ph(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("parse_core") / "libparse_host.so")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Werror", "-o", so, SRC])
    L = C.CDLL(so)
    L.ph_parse4.restype = C.c_int
    L.ph_parse4.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_void_p]
    return L


def lz4_bytes_from_records(data, recs, offs):
    """lz4_emit_kernel's layout: token, literal length bytes, literals, LE16 offset, match length bytes"""
    out = bytearray()
    for (x, y), off in zip(recs.tolist(), offs.tolist()):
        anchor, lit, ml, op = x & 0xFFFF, x >> 16, y & 0xFFFF, y >> 16
        assert op == len(out), "record's output offset is not the running sum"
        mlc = ml - 4
        out.append((min(lit, 15) << 4) | (min(mlc, 15) if ml else 0))
        if lit >= 15:
            r = lit - 15
            while r >= 255:
                out.append(255)
                r -= 255
            out.append(r)
        out += data[anchor:anchor + lit]
        if ml:
            out += bytes([off & 0xFF, off >> 8])
            if mlc >= 15:
                r = mlc - 15
                while r >= 255:
                    out.append(255)
                    r -= 255
                out.append(r)
    return bytes(out)


def corpora():
    rng = np.random.default_rng(7)
    tera = oracle.gen_terasort(3, 700).tobytes()
    words = [b"shuffle", b"block", b"spark", b"partition", b"index", b"checksum", b" ", b"\n", b"s3a://bucket/", b"0000"]
    text = b"".join(words[i] for i in rng.integers(0, len(words), 9000))
    rows = b"".join(int(v).to_bytes(8, "little") + b"\x00" * 8 + int(v % 97).to_bytes(8, "little")
                    for v in rng.zipf(1.3, 1500))
    return {
        "terasort": tera, "zeros": bytes(40000), "random": rng.integers(0, 256, 33000, dtype=np.uint8).tobytes(),
        "text": text, "rows": rows, "period3": (b"abc" * 12000), "period7": (b"0123456" * 5000),
        "runs": b"".join(bytes([i & 255]) * (1 + (i * 7) % 40) for i in range(2500)),
        "ab_long": b"A" * 300 + rng.integers(0, 256, 100, dtype=np.uint8).tobytes() + b"A" * 5000 + b"tail-bytes-here",
    }


# ------------------------------------------------------------------------------------------------------------------
# generation 4: sub-chunk parallel parse (walk_subchunk per lane + the kernel's stitch, restated in the harness)
def parse4(L, codec, data, block_size, sb=0, hash_log=12):
    a = np.frombuffer(bytes(data), dtype=np.uint8) if len(data) else np.zeros(0, dtype=np.uint8)
    n = a.size
    rec = np.zeros(3 * (n // 4 + 34) + 8, dtype=np.uint32)
    nseq, csize, size = C.c_uint32(), C.c_uint32(), C.c_uint64()
    buf = np.ascontiguousarray(a)
    rc = L.ph_parse4(codec, buf.ctypes.data if n else None, n, sb, hash_log, block_size, C.byref(nseq), C.byref(csize),
                     C.byref(size), rec.ctypes.data)
    assert rc == 0
    k = nseq.value
    return rec[: 2 * k].reshape(k, 2), rec[2 * (n // 4 + 34): 2 * (n // 4 + 34) + k], csize.value, size.value


@pytest.mark.parametrize("name", sorted(corpora().keys()))
def test_subchunk_parse_equals_the_specification(ph, name):
    """32 lanes x walk_subchunk + stitch == orc_lz4_compress_block_win_sub with the kernel's sub-chunk size"""
    data = corpora()[name]
    for bs, n in ((32768, 32768), (32768, 32767), (32768, 20001), (32768, 1025), (32768, 1024), (32768, 1023),
                  (32768, 40), (32768, 13), (32768, 12), (32768, 0), (4096, 4096), (4096, 777), (64, 64), (64, 33),
                  (65536, 65536), (65536, 40000)):
        blk = (data * (1 + n // max(len(data), 1)))[:n]
        sub = oracle.lz4_subchunk(bs)
        want = oracle.lz4_compress_block(blk, win=True, cap=max(n - 1, 0), sub=sub) if n else None
        for sb in (0, 3):
            recs, offs, csize, size = parse4(ph, 0, blk, bs, sb)
            if want is None:
                assert csize == (n | 0x80000000) and size == 21 + n, (name, bs, n, sb)
                continue
            got = lz4_bytes_from_records(blk, recs, offs)
            assert got == want, (name, bs, n, sb)
            assert csize == len(want) and size == 21 + len(want)


def test_subchunking_costs_little_ratio():
    tera = oracle.gen_terasort(11, 3000).tobytes()[:10 * 32768]
    whole = sum(len(oracle.lz4_compress_block(tera[i:i + 32768], win=True)) for i in range(0, len(tera), 32768))
    cut = sum(len(oracle.lz4_compress_block(tera[i:i + 32768], win=True, sub=1024)) for i in range(0, len(tera), 32768))
    assert whole <= cut <= whole * 1.01


def test_subchunk_parse_snappy_and_zstd_grammars(ph):
    """generation 4 with the Snappy / Zstandard size rules: the Snappy element stream of the specification
    (xerial_compress(compressor=2): BE32 length + raw block per chunk) has exactly csize bytes, and all three grammars
    select the same matches"""
    c = corpora()
    for name in ("terasort", "text", "zeros", "runs", "rows"):
        blk = c[name][:32768]
        r0, o0, _, _ = parse4(ph, 0, blk, 32768, 1)
        r1, o1, cs1, sz1 = parse4(ph, 1, blk, 32768, 1)
        r2, o2, _, _ = parse4(ph, 2, blk, 32768, 1)
        x = oracle.xerial_compress(blk, 32768, compressor=2)
        assert len(x) == 16 + 4 + cs1 and sz1 == 4 + cs1, name
        ms = [[(int(x_ & 0xFFFF) + int(x_ >> 16), int(y & 0xFFFF), int(o)) for (x_, y), o in zip(r.tolist(), o.tolist())
               if y & 0xFFFF] for r, o in ((r0, o0), (r1, o1), (r2, o2))]
        assert ms[0] == ms[1] == ms[2], name
        lits = 0
        for x_, y in r2.tolist():
            assert (y >> 16) == lits
            lits += x_ >> 16
        assert lits + sum(m[1] for m in ms[2]) == len(blk)
