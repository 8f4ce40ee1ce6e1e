"""spark-s3-shuffle_b200 — B200-native shuffle-block codec path for IBM/spark-s3-shuffle.

The product is ``libb200shuffle.so`` (hand-written sm_100a CUDA behind the flat C ABI in ``include/b200shuffle.h``).
This package is the thin Python face used by tests and ``bench.py``: it builds/loads the library with ``ctypes`` and
offers numpy-friendly wrappers that call *through the C ABI* exactly as the JNI shim does.

There is deliberately no CPU fallback: if the library cannot be built or loaded, importing ``capi`` raises, and every
compute entry point returns ``B2S_E_CUDA`` when no GPU is usable.
"""
import importlib
import os

__all__ = ["capi", "host", "ranks", "lib_path", "build"]
_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "libb200shuffle.so")


def build(force=False, verbose=False):
    from . import _build as _b

    return _b.build(force=force, verbose=verbose)


def __getattr__(name):
    if name in ("capi", "host", "ranks", "_build"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
