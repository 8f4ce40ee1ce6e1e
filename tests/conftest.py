"""pytest configuration: registers the `gpu` marker, puts the repo root on sys.path, shared fixtures.

`-m "not gpu"` runs here (no GPU): the oracle against golden vectors / native libraries, the host logic and the
C-ABI load/export checks.  `-m gpu` runs on the B200 box: parity of the CUDA path (through the C ABI) vs the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.lib()
    return o


@pytest.fixture(scope="session")
def capi():
    """Initialised C-ABI binding; only gpu-marked tests use it."""
    import spark_s3_shuffle_b200 as pkg

    c = pkg.capi
    c.init(0)
    yield c
    c.shutdown()


def corpus(oracle, kind, n, seed=0):
    """Deterministic byte corpora covering the shapes the codecs meet."""
    rng = np.random.default_rng(seed)
    if kind == "terasort":
        recs = (n + 103) // 104
        return oracle.gen_terasort(seed * 1000, recs).tobytes()[:n]
    if kind == "zeros":
        return bytes(n)
    if kind == "random":
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    if kind == "text":
        words = [b"shuffle", b"spark", b"partition", b"block", b"reduce", b"map", b"s3", b"the", b"of", b"and",
                 b"checksum", b"stream", b"index", b"0123456789", b"\n", b" ", b" ", b" "]
        idx = rng.integers(0, len(words), n // 3 + 8)
        return b"".join(words[i] for i in idx)[:n]
    if kind == "runs":
        out = bytearray()
        while len(out) < n:
            out += bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 700))
        return bytes(out[:n])
    if kind == "period":
        p = rng.integers(0, 256, 37, dtype=np.uint8).tobytes()
        return (p * (n // 37 + 1))[:n]
    if kind == "ints":
        return (np.arange(n // 4 + 1, dtype=np.uint32) * 7 + seed).tobytes()[:n]
    raise ValueError(kind)


KINDS = ["terasort", "zeros", "random", "text", "runs", "period", "ints"]
