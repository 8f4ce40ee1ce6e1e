"""The group-commit queue of the host library (spark-s3-shuffle_b200/host/coalesce.h) on its own: compiled into a native
harness whose two C-ABI batch calls are stand-ins defined in the harness (tests/native/coalesce_host.cpp), so merging,
per-request hand-back, leadership hand-over and error propagation are checked without a GPU.  (The same harness is
clean under -fsanitize=thread.)  The real thing — the queue over the CUDA path — is tests/test_gpu_host_streams.py."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("coalesce") / "libcoalesce.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", out,
                           os.path.join(ROOT, "tests", "native", "coalesce_host.cpp")])
    L = C.CDLL(out)
    L.coalesce_run.restype = C.c_int
    L.coalesce_run.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    return L


def run(L, threads, reps, fail_one=0):
    st = (C.c_uint64 * 6)()
    bad = L.coalesce_run(threads, reps, fail_one, st)
    return bad, dict(zip(("calls", "batches", "maxMerged", "streams", "backend_calls", "failed"), list(st)))


def test_single_thread_runs_every_call_at_once(harness):
    bad, st = run(harness, 1, 10)
    assert bad == 0
    assert st["calls"] == 20 and st["batches"] == 20 and st["maxMerged"] == 1 and st["backend_calls"] == 20


def test_concurrent_callers_are_merged_and_get_their_own_results(harness):
    bad, st = run(harness, 8, 25)
    assert bad == 0, "a caller received another caller's streams, lengths, checksums or status"
    assert st["calls"] == 8 * 25 * 2 and st["streams"] == st["calls"] * 5
    assert st["backend_calls"] == st["batches"] < st["calls"] and 2 <= st["maxMerged"] <= 8


def test_a_call_level_failure_of_a_merged_batch_is_isolated_per_request(harness):
    """one failing backend call: a merged round is re-run request by request (the stand-in fails only once, so every
    task then succeeds); only a request that failed on its own keeps the error — nobody pays for a neighbour's block"""
    bad, st = run(harness, 6, 10, fail_one=1)
    assert bad == 0
    assert st["failed"] <= 1
    assert st["calls"] == 6 * 10 * 2 - st["failed"]   # a failed compress skips its read-back
