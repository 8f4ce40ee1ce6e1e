"""In-process multi-GPU: b2s_*_batch shards streams round-robin over every device selected by b2s_init (stream i ->
device i mod D), and packed calls follow the calling thread's device.  Needs >= 2 GPUs (gpurun --gpus 2); skipped
otherwise.  Results must be identical to the single-device ones."""
import numpy as np
import pytest

from conftest import corpus

pytestmark = pytest.mark.gpu


def test_batch_calls_shard_over_all_devices_with_identical_results(capi, oracle):
    L = capi.load()
    if L.b2s_device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    parts = [corpus(oracle, k, 20000 + 777 * i, seed=i) for i, k in enumerate(["terasort", "text", "runs", "zeros", "random"] * 5)]
    for codec in (capi.CODEC_LZ4BLOCK, capi.CODEC_SNAPPY_XERIAL, capi.CODEC_ZSTD):
        comp, cks, st = capi.compress_batch(codec, parts, 32768, capi.CHECKSUM_CRC32C)
        assert st == [0] * len(parts)
        if codec == capi.CODEC_LZ4BLOCK:
            assert comp == [oracle.lz4block_compress(p, 32768, compressor=1) for p in parts]
        assert cks == [oracle.crc32c(c) for c in comp]
        slices = [[(len(c), k)] for c, k in zip(comp, cks)]
        out, st, _ = capi.decompress_batch(codec, comp, capi.CHECKSUM_CRC32C, slices)
        assert st == [0] * len(parts) and out == parts
    assert capi.checksum_batch(capi.CHECKSUM_ADLER32, parts) == [oracle.adler32(p) for p in parts]
    assert capi.last_timing()["kernel_launches"] > 0


def test_thread_device_selects_the_gpu_of_packed_calls(capi, oracle):
    L = capi.load()
    if L.b2s_device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    p = corpus(oracle, "terasort", 300000, 3)
    src = np.frombuffer(p, dtype=np.uint8)
    off, ln = np.array([0], dtype=np.uint64), np.array([len(p)], dtype=np.uint64)
    want = oracle.lz4block_compress(p, 32768, compressor=1)
    for dev in (1, 0):
        assert L.b2s_set_thread_device(dev) == 0
        dst = np.empty(capi.compress_bound(capi.CODEC_LZ4BLOCK, 32768, len(p)), dtype=np.uint8)
        w = capi.compress_packed(capi.CODEC_LZ4BLOCK, src, off, ln, dst, 32768)
        assert dst[: w["total"]].tobytes() == want
    assert L.b2s_set_thread_device(99) == capi.E_ARG
