#!/usr/bin/env python
"""Ratio of the GPU Zstandard encoder's CPU model (tests/native/zstd_core_host.cpp: the same zstd_enc_core.h functions
the kernels run) against libzstd levels 1 and 3 on the test corpora — no GPU needed.  Every model frame is decoded by
libzstd.  Usage: python tools/zstd_ratio_probe.py [-DkLLLog=.. style defines are edited in zstd_enc_core.h]"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import zstd_ref  # noqa: E402
from conftest import KINDS, corpus  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    out = os.path.join(tempfile.mkdtemp(), "libzc.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out] + sys.argv[1:] +
                          [os.path.join(ROOT, "tests", "native", "zstd_core_host.cpp")])
    L = C.CDLL(out)
    L.zc_compress_model_hlog.restype = C.c_longlong
    L.zc_compress_model_hlog.argtypes = [C.c_char_p, C.c_ulonglong, C.c_uint, C.c_char_p, C.c_ulonglong, C.c_int]
    n = 671112
    print("%-10s %8s %8s %8s | %8s %8s" % ("corpus", "hlog11", "hlog12", "hlog13", "zstd-1", "zstd-3"))
    for kind in KINDS:
        d = corpus(oracle, kind, n, seed=1)
        row = []
        for hlog in (11, 12, 13):
            cap = n + n // 64 + 1024
            buf = C.create_string_buffer(cap)
            c = L.zc_compress_model_hlog(d, n, 32768, buf, cap, hlog)
            assert c > 0 and zstd_ref.decompress(buf.raw[:c]) == d
            row.append(c / n)
        ref = [len(zstd_ref.compress_stream(d, lvl)) / n for lvl in (1, 3)]
        print("%-10s %8.4f %8.4f %8.4f | %8.4f %8.4f" % (kind, *row, *ref))


if __name__ == "__main__":
    main()
