"""N>1 path on CPU: world_size 2 over gloo (SURVEY.md §8e — blocks shard i mod N, no data-path collective)."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(args, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port())] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_round_robin_sharding_and_metadata_gather_world2():
    r = _torchrun([os.path.join(ROOT, "tests", "multirank_worker.py")])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["ok"] and out["same_index"]
    assert out["max"] == 2.0 and out["sum"] == 11.0
    assert out["index_len"] == 12 * 8


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    r = _torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                   "--blocks", "40", "--cpu-sample-bytes", "30000000"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["n_gpus"] == 2 and out["value"] > 0
    assert out["e2e"]["h2d_bytes_per_step"] == 0 and out["cpu_baseline"]["kind"] == "port"


import pytest  # noqa: E402


@pytest.mark.parametrize("codec", ["lz4", "snappy", "zstd"])
def test_reference_arm_contract_for_every_codec(codec):
    """`bench.py --impl reference --codec X`: one JSON line with the driver's keys, timed on the CPU arithmetic of that
    codec (liblz4 / restated snappy / libzstd)"""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--codec", codec, "--steps", "1",
                        "--warmup", "1", "--blocks", "30", "--cpu-sample-bytes", "20000000"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in out, k
    assert out["impl"] == "reference" and out["value"] > 0 and out["unit"] == "GB/s" and out["dtype"] == "u8"
    assert out["cpu_baseline"]["cores"] >= 1 and 0.1 < out["cpu_baseline"]["compressed_ratio"] < 0.9
    assert {"lz4": "liblz4", "snappy": "Snappy", "zstd": "libzstd"}[codec] in out["cpu_baseline"]["note"]
