"""B2S_LZ4_PIPE=4 (match2 + sub-chunk parallel parse) on a real device.  The pipeline generation is read once at b2s_init,
so this runs in a subprocess; the bytes must equal the sub-chunk specification (compressor 2 in oracle/) and decode back
through liblz4's rules (the oracle decoder) and through the GPU decoder."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import spark_s3_shuffle_b200 as pkg
from oracle import oracle
from conftest import corpus
c = pkg.capi
c.init(1)
for bs in (32768, 4096, 65536):
    parts = [corpus(oracle, kind, n, seed=i) for i, (kind, n) in enumerate(
        [("terasort", 300000), ("text", 100000), ("zeros", 70000), ("random", 40000), ("terasort", 13), ("text", 0),
         ("runs", 33000)])]
    comp, cks, st = c.compress_batch(c.CODEC_LZ4BLOCK, parts, bs, c.CHECKSUM_CRC32C)
    assert st == [0] * len(parts), st
    for p, s, k in zip(parts, comp, cks):
        assert s == oracle.lz4block_compress(p, bs, compressor=2), "pipe 4 output differs from the sub-chunk specification"
        assert oracle.lz4block_decompress(s) == p and k == oracle.crc32c(s)
    out, st, _ = c.decompress_batch(c.CODEC_LZ4BLOCK, comp)
    assert st == [0] * len(parts) and out == parts
parts = [corpus(oracle, "terasort", 200000, seed=3), corpus(oracle, "ints", 90000, seed=4)]
comp, _, st = c.compress_batch(c.CODEC_SNAPPY_XERIAL, parts, 32768)
assert st == [0, 0] and comp == [oracle.xerial_compress(p, 32768, compressor=2) for p in parts]
comp, _, st = c.compress_batch(c.CODEC_ZSTD, parts, 32768)
out, st2, _ = c.decompress_batch(c.CODEC_ZSTD, comp)
assert st == [0, 0] and st2 == [0, 0] and out == parts
print("PIPE4-OK")
'''


def test_pipe4_equals_the_subchunk_specification():
    env = dict(os.environ, B2S_LZ4_PIPE="4")
    r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, os.path.join(ROOT, "tests"))], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "PIPE4-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
