"""Generates tests/golden/vectors.json — golden vectors for the codec boundary.

The reference's own tests hold no vectors for this path, so these are produced HERE from the native
libraries its JVM dependencies wrap, never from our own code:
  * liblz4.so.1 (1.9.4)  LZ4_compress_default  -> payloads of LZ4Block streams (framing assembled from the lz4-java
    wire layout: magic | token | compressedLen | originalLen | XXH32(seed 0x9747b28c) & 0x0FFFFFFF)
  * xxhash (python binding of the reference xxHash C code) for the block checksums
  * zlib crc32/adler32, and the standard CRC-32C check values
  * pyarrow's bundled snappy for raw snappy blocks wrapped in the xerial SnappyOutputStream framing
Run from the repo root:  python tests/golden/make_golden.py
"""
import ctypes as C
import json
import os
import zlib

import numpy as np
import pyarrow as pa
import xxhash

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 0x9747B28C


def inputs():
    rng = np.random.default_rng(20260922)
    words = [b"shuffle", b"spark", b"partition", b" ", b"block", b"\n", b"reduce", b"0123456789"]
    text = b"".join(words[i] for i in rng.integers(0, len(words), 9000))
    return {
        "empty": b"",
        "one": b"a",
        "abc": b"abc",
        "twelve": b"abcdefghijkl",
        "thirteen": b"abcdefghijklm",
        "zeros_100k": bytes(100000),
        "text_40k": text[:40000],
        "random_5k": rng.integers(0, 256, 5000, dtype=np.uint8).tobytes(),
        "period_70k": (rng.integers(0, 256, 37, dtype=np.uint8).tobytes() * 2000)[:70000],
        "ints_33k": (np.arange(9000, dtype=np.uint32) * 7).tobytes()[:33000],
    }


def lz4block_stream(L, data, block_size=32768):
    level = max(0, (block_size - 1).bit_length() - 10)
    out = b""
    for o in range(0, len(data), block_size):
        chunk = data[o:o + block_size]
        cap = L.LZ4_compressBound(len(chunk))
        buf = C.create_string_buffer(cap)
        c = L.LZ4_compress_default(chunk, buf, len(chunk), cap)
        method, payload = (0x20, buf.raw[:c]) if c < len(chunk) else (0x10, chunk)
        chk = xxhash.xxh32(chunk, seed=SEED).intdigest() & 0x0FFFFFFF
        out += (b"LZ4Block" + bytes([method | level]) + len(payload).to_bytes(4, "little")
                + len(chunk).to_bytes(4, "little") + chk.to_bytes(4, "little") + payload)
    return out + b"LZ4Block" + bytes([0x10 | level]) + bytes(12)


def xerial_stream(data, block_size=32768):
    codec = pa.Codec("snappy")
    out = bytes([0x82]) + b"SNAPPY\x00" + (1).to_bytes(4, "big") + (1).to_bytes(4, "big")
    for o in range(0, len(data), block_size):
        c = codec.compress(data[o:o + block_size]).to_pybytes()
        out += len(c).to_bytes(4, "big") + c
    return out


def main():
    L = C.CDLL("liblz4.so.1")
    vec = {"_generator": "tests/golden/make_golden.py", "_liblz4": L.LZ4_versionNumber(), "cases": {}}
    for name, data in inputs().items():
        s = lz4block_stream(L, data)
        x = xerial_stream(data)
        vec["cases"][name] = {
            "input_hex": data.hex() if len(data) <= 64 else None,
            "input_len": len(data),
            "crc32": zlib.crc32(data), "adler32": zlib.adler32(data),
            "xxh32_seed9747b28c": xxhash.xxh32(data, seed=SEED).intdigest(),
            "lz4block_stream_hex": s.hex(),
            "lz4block_stream_crc32": zlib.crc32(s), "lz4block_stream_adler32": zlib.adler32(s),
            "xerial_stream_hex": x.hex(),
        }
    vec["kat"] = {
        "crc32_123456789": 0xCBF43926, "crc32c_123456789": 0xE3069283, "adler32_123456789": 0x091E01DE,
        "crc32c_32_zero_bytes": 0x8A9136AA, "crc32c_32_ff_bytes": 0x62A8AB43,       # RFC 3720 B.4
        "crc32c_32_incrementing": 0x46DD794E, "crc32c_32_decrementing": 0x113FDB5C,   # RFC 3720 B.4
        "xxh32_abc_seed9747b28c": 0x4D4CB222, "xxh32_empty_seed0": 0x02CC5D05,
        "lz4block_empty_stream_hex": "4c5a34426c6f636b15" + "00" * 12,
        "index_for_lengths_3_0_5_hex": "0000000000000000" "0000000000000003" "0000000000000003" "0000000000000008",
        "xerial_header_hex": "82534e41505059000000000100000001",
    }
    with open(os.path.join(HERE, "vectors.json"), "w") as f:
        json.dump(vec, f, indent=0, sort_keys=True)
    print("wrote", os.path.join(HERE, "vectors.json"))


if __name__ == "__main__":
    main()
