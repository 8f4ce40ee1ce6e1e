#!/usr/bin/env python
"""bench.py — shuffle write+read GB/s (compress + CRC / verify + decompress) on BASELINE.json config[1]:
terasort-shaped data, 10 GiB, 80 maps x 200 partitions = 16,000 shuffle blocks, LZ4Block 32 KiB + CRC32C, 1 x B200
(per GPU; with --gpus N every rank processes its own 10 GiB => weak scaling, no data-path collective).

One "step" = one write pass (b2s_compress_*: XXH32 + LZ4 + framing + CRC32C over the compressed streams) plus one read
pass (b2s_decompress_*: CRC32C verify + LZ4 decode + XXH32 verify) over the whole dataset.
  value : uncompressed bytes / (t_write + t_read), inputs and outputs resident in HBM (device API)
  e2e   : the same through the host-pointer C ABI a JVM would call (pinned host buffers; H2D and D2H inside the timed region)
  roofline : dominant kernel (lz4_compress_kernel) algorithmic bytes (1+r)*U per launch / CUDA-event duration vs measured HBM peak
  cpu_baseline : the reference's CPU arithmetic (liblz4 LZ4_compress_default/LZ4_decompress_fast + LZ4Block framing +
                 XXH32 + CRC32C, oracle/), all host cores, bounded sample of the same data
`--impl reference` times that CPU path alone on the same config (the reference itself is Scala on a JVM, absent here).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RECORD = 104
RECORDS_PER_BLOCK = 6453            # 671,112 B per shuffle block
N_BLOCKS_FULL = 16000               # 80 maps x 200 reduce partitions  => 10.0003 GiB
LZ4_BLOCK = 32768
# dram__bytes_read.sum + dram__bytes_write.sum of one lz4_match_kernel launch over a 1.25 GiB chunk (ncu --set full,
# profiles/r1c_*), scaled to the bench's 1 GiB chunks at run time; None until a capture of the current kernel exists
NCU_TRAFFIC_PER_LAUNCH = 6089710000  # 1.948 GB read + 4.141 GB written, full 32,768-block launch (profiles/r1p_multikernel_lz4.md)
METRIC = "shuffle write+read GB/s (compress+CRC) at 1/2/4/8 B200 vs JVM-LZ4 CPU baseline"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


CPU_CODEC = "lz4"  # set from --codec
CPU_NOTES = {
    "lz4": "oracle framing/XXH32/CRC32C + liblz4.so.1 LZ4_compress_default/LZ4_decompress_fast (the native routines "
           "lz4-java's JNI path calls); no JVM stream wrappers / JNI => upper bound on the reference",
    "snappy": "oracle xerial framing + restated raw-Snappy compressor/decompressor + CRC32C (snappy-java's native side is "
              "absent here); no JVM stream wrappers / JNI",
    "zstd": "libzstd.so.1 ZSTD_compress level 3 / ZSTD_decompress (what zstd-jni executes; BASELINE config 5's level) + "
            "CRC32C; no JVM stream wrappers / JNI => upper bound on the reference",
}


def cpu_arm(oracle, sample, block_bytes, threads, repeats=1):
    """reference CPU arithmetic on `sample` (numpy uint8): returns (GB/s write+read, detail)"""
    best = None
    for _ in range(repeats):
        r = oracle.baseline_run_codec(CPU_CODEC, sample, block_bytes, LZ4_BLOCK, oracle.CRC32C, threads=threads, level=3)
        if r["rc"] != 0:
            raise RuntimeError("CPU baseline reported %d errors" % r["errors"])
        t = r["write_s"] + r["read_s"]
        if best is None or t < best[0]:
            best = (t, r)
    t, r = best
    return r["bytes"] / t / 1e9, r


def main():
    global RECORDS_PER_BLOCK, CPU_CODEC
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blocks", type=int, default=int(os.environ.get("B2S_BENCH_BLOCKS", N_BLOCKS_FULL)),
                    help="shuffle blocks per GPU (16000 = the 10 GiB config; smaller only for quick checks)")
    ap.add_argument("--records-per-block", type=int, default=RECORDS_PER_BLOCK,
                    help="104-byte records per shuffle block: 6453 = config 2 (default), 630 = config 3's ~64 KiB blocks")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample-blocks", type=int, default=6000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--codec", default="lz4", choices=["lz4", "snappy", "zstd"],
                    help="lz4 = the BASELINE configuration; snappy / zstd report the other codecs on the same data")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = args.steps, max(args.warmup, 3)
    n = args.blocks
    RECORDS_PER_BLOCK = args.records_per_block
    CPU_CODEC = args.codec
    block_bytes = RECORDS_PER_BLOCK * RECORD
    total = n * block_bytes
    workload = "terasort %.2f GiB/GPU, %d shuffle blocks x %d B (%s), LZ4Block 32 KiB + CRC32C" % (
        total / 2**30, n, block_bytes,
        "80 maps x 200 partitions" if RECORDS_PER_BLOCK == 6453 else "config-3 shape: 2 full LZ4 blocks + end mark per block")
    codec_desc = {"lz4": "lz4 (LZ4Block, blockSize 32 KiB)", "snappy": "snappy (xerial framing, blockSize 32 KiB)",
                  "zstd": "zstd (frames of 32 KiB blocks; raw literals + predefined-FSE sequences)"}[args.codec]
    if args.codec != "lz4":
        workload = workload.replace("LZ4Block 32 KiB", codec_desc)
    config = {"workload": workload, "codec": codec_desc, "checksum": "CRC32C over compressed bytes",
              "blocks_per_gpu": n, "block_bytes": block_bytes, "l2_policy": "inputs (>= 6 GiB per pass) far larger than the 126 MB L2",
              "sharding": "block i -> GPU i mod N (each rank owns its blocks; no collective on the data path)"}

    from oracle import oracle  # CPU baseline legs only (never on the GPU product path)

    # ------------------------------------------------------------------ reference arm: CPU only
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        sample_blocks = min(args.cpu_sample_blocks, n)
        sample = oracle.gen_terasort(0, sample_blocks * RECORDS_PER_BLOCK)
        for _ in range(max(1, min(warmup, 1))):
            cpu_arm(oracle, sample[: 200 * block_bytes], block_bytes, threads)
        vals, t0 = [], time.perf_counter()
        for _ in range(steps):
            v, detail = cpu_arm(oracle, sample, block_bytes, threads)
            vals.append(v)
        wall = time.perf_counter() - t0
        value = sample.size * steps / sum(sample.size / (v * 1e9) for v in vals) / 1e9
        sample_desc = "%d of %d shuffle blocks (%.2f GiB) per step, same generator/seed" % (sample_blocks, n, sample.size / 2**30)
        line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": round(wall / steps * 1e3, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": round(value, 3), "unit": "GB/s", "cores": threads,
                                 "kind": "port", "sample": sample_desc,
                                 "note": CPU_NOTES[args.codec],
                                 "compressed_ratio": round(detail["compressed_bytes"] / detail["bytes"], 4)},
                "e2e": {"value": round(value, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch

    torch.cuda.set_device(local_rank)
    import spark_s3_shuffle_b200 as pkg

    rk = pkg.ranks.Ranks(backend="nccl", device=torch.device("cuda", local_rank))  # barrier + max-over-ranks only
    c = pkg.capi
    c.init(1 << local_rank)
    L = c.load()
    barrier, max_over_ranks = rk.barrier, rk.max_over_ranks

    CODEC = {"lz4": c.CODEC_LZ4BLOCK, "snappy": c.CODEC_SNAPPY_XERIAL, "zstd": c.CODEC_ZSTD}[args.codec]
    cmp_cap = int(c.compress_bound(CODEC, LZ4_BLOCK, block_bytes)) * n
    d_src, d_cmp, d_out = c.dev_alloc(total), c.dev_alloc(cmp_cap), c.dev_alloc(total)
    c.gen_terasort_dev(d_src, rank * n * RECORDS_PER_BLOCK, n * RECORDS_PER_BLOCK, 42)
    off = np.arange(n, dtype=np.uint64) * block_bytes
    ln = np.full(n, block_bytes, dtype=np.uint64)
    sb = np.arange(n + 1, dtype=np.uint32)

    kt = {"write_kernel_ms": [], "read_kernel_ms": [], "compress_ms": [], "decompress_ms": [], "match_ms": [],
          "match_launches": []}

    def step_device(record):
        w = c.compress_dev(CODEC, d_src, off, ln, d_cmp, cmp_cap, LZ4_BLOCK, c.CHECKSUM_CRC32C)
        tw = c.last_timing()
        r = c.decompress_dev(CODEC, d_cmp, w["dst_off"], w["dst_len"], d_out, total, c.CHECKSUM_CRC32C, sb,
                             w["dst_len"], w["checksums"])
        tr = c.last_timing()
        if record:
            kt["write_kernel_ms"].append(tw["kernel_ms"]); kt["compress_ms"].append(tw["top_kernel_ms"])
            kt["match_ms"].append(tw["dominant_ms"]); kt["match_launches"].append(tw["dominant_launches"])
            kt["read_kernel_ms"].append(tr["kernel_ms"]); kt["decompress_ms"].append(tr["top_kernel_ms"])
        return w, r

    want = c.checksum_dev(c.CHECKSUM_CRC32C, d_src, off, ln)
    for _ in range(warmup):
        w, r = step_device(False)
    assert not w["status"].any() and not r["status"].any() and r["total"] == total, "device round trip failed"
    got = c.checksum_dev(c.CHECKSUM_CRC32C, d_out, off, ln)
    assert (got == want).all(), "decode(encode(x)) != x"
    ratio = w["total"] / total

    sampler = ClockSampler(local_rank)
    launches0 = L.b2s_total_kernel_launches()
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    c.mark(0)                      # CUDA events on the library's own stream (every call is synchronous)
    for _ in range(steps):
        step_device(True)
    c.mark(1)
    dev_elapsed = c.marks_elapsed_ms() * 1e-3
    barrier()
    wall = time.perf_counter() - t0
    elapsed = max_over_ranks(max(dev_elapsed, 0.0) if dev_elapsed > 0.5 * wall else wall)
    clocks = sampler.stop()
    launches = L.b2s_total_kernel_launches() - launches0
    ms_per_step = elapsed / steps * 1e3
    value = world * total / (elapsed / steps) / 1e9

    comp_ms = statistics.mean(kt["compress_ms"])
    match_ms = statistics.mean(kt["match_ms"])
    match_launches = max(1, int(statistics.mean(kt["match_launches"])))
    peak, peak_src = measured_peak()
    # dominant kernel = lz4_match_kernel (phase A of the compressor; largest share in profiles/*launches*.csv).
    # achieved = SURVEY §8(d)'s write-step figure (1+r) x the input bytes one launch processes / its launch duration.
    step_alg = (1.0 + ratio) * total                       # compress step reads U, writes C (= decompress step mirrored)
    alg_bytes = step_alg / match_launches
    achieved = alg_bytes / (match_ms / match_launches * 1e-3) / 1e9
    dec_ms = statistics.mean(kt["decompress_ms"])
    roofline = {"bound": "hbm", "kernel": "lz4_match_kernel<12>", "achieved": round(achieved, 2), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 5), "traffic": NCU_TRAFFIC_PER_LAUNCH,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(alg_bytes), "avg_launch_ms": round(match_ms / match_launches, 4),
                "launches_per_step": match_launches,
                "compress_step": {"kernels": "lz4_match + lz4_parse + scan + lz4_emit (what one fused kernel would do)",
                                  "algorithmic_bytes": int(step_alg), "ms": round(comp_ms, 3),
                                  "achieved": round(step_alg / (comp_ms * 1e-3) / 1e9, 2),
                                  "frac": round(step_alg / (comp_ms * 1e-3) / 1e9 / peak, 5)},
                "decompress_step": {"kernels": "lz4_tokens + lz4_copy", "algorithmic_bytes": int(step_alg),
                                    "ms": round(dec_ms, 3), "achieved": round(step_alg / (dec_ms * 1e-3) / 1e9, 2),
                                    "frac": round(step_alg / (dec_ms * 1e-3) / 1e9 / peak, 5)},
                "note": "LZ77 coding is issue/latency-bound byte-stream work; traffic (ncu dram bytes per launch, "
                        "profiles/) exceeds the algorithmic bytes by the per-position match table handed to the parse kernel"}
    kernels = {"compress_ms": round(comp_ms, 3), "match_ms": round(match_ms, 3), "decompress_ms": round(statistics.mean(kt["decompress_ms"]), 3),
               "write_pass_kernels_ms": round(statistics.mean(kt["write_kernel_ms"]), 3),
               "read_pass_kernels_ms": round(statistics.mean(kt["read_kernel_ms"]), 3),
               "compress_GBps_uncompressed": round(total / comp_ms / 1e6, 2),
               "decompress_GBps_uncompressed": round(total / statistics.mean(kt["decompress_ms"]) / 1e6, 2)}

    # ------------------------------------------------------------------ e2e through the host-pointer C ABI
    e2e = None
    if not args.no_e2e:
        # pinned host arenas: 3 x ~10 GiB per rank.  If any rank cannot get them (host RAM at N=8), every rank falls
        # back together to a quarter of the blocks — and the line says so — instead of one rank dying at a barrier.
        def alloc(nblk):
            bufs = []
            try:
                for nbytes in (nblk * block_bytes, int(c.compress_bound(CODEC, LZ4_BLOCK, block_bytes)) * nblk,
                               nblk * block_bytes):
                    bufs.append(c.HostBuffer(nbytes))
                return bufs
            except Exception:
                for hb in bufs:
                    hb.free()
                return None

        n_e = n
        bufs = alloc(n_e)
        if max_over_ranks(0.0 if bufs is not None else 1.0) > 0:
            if bufs is not None:
                for hb in bufs:
                    hb.free()
            n_e = max(1, n // 4)
            bufs = alloc(n_e)
        if max_over_ranks(0.0 if bufs is not None else 1.0) > 0:
            if bufs is not None:
                for hb in bufs:
                    hb.free()
            e2e = {"value": None, "unit": "GB/s", "error": "pinned host memory for the end-to-end leg could not be allocated"}
        else:
            h_src, h_cmp, h_out = bufs
            total_e = n_e * block_bytes
            off_e, ln_e, sb_e = off[:n_e], ln[:n_e], sb[: n_e + 1]
            c.dev_memcpy(h_src.ptr, d_src, total_e, 2)
            for p in (d_src, d_cmp, d_out):
                c.dev_free(p)
            d_src = d_cmp = d_out = None

            def step_host():
                w = c.compress_packed(CODEC, h_src.array, off_e, ln_e, h_cmp.array, LZ4_BLOCK, c.CHECKSUM_CRC32C)
                tw = c.last_timing()
                r = c.decompress_packed(CODEC, h_cmp.array, w["dst_off"], w["dst_len"], h_out.array,
                                        c.CHECKSUM_CRC32C, sb_e, w["dst_len"], w["checksums"])
                tr = c.last_timing()
                return w, r, tw, tr

            w, r, tw, tr = step_host()   # warm-up (allocates slot buffers)
            assert not w["status"].any() and not r["status"].any() and r["total"] == total_e
            for i in (0, n_e // 2, n_e - 1):
                a = h_src.array[i * block_bytes:(i + 1) * block_bytes]
                b = h_out.array[int(r["dst_off"][i]):int(r["dst_off"][i]) + block_bytes]
                assert np.array_equal(a, b), "e2e round trip mismatch"
            step_host()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                w, r, tw, tr = step_host()
            barrier()
            e_elapsed = max_over_ranks(time.perf_counter() - t0)
            e_val = world * total_e / (e_elapsed / args.e2e_steps) / 1e9
            e2e = {"value": round(e_val, 3), "unit": "GB/s",
                   "h2d_bytes_per_step": int(tw["h2d_bytes"] + tr["h2d_bytes"]),
                   "d2h_bytes_per_step": int(tw["d2h_bytes"] + tr["d2h_bytes"]),
                   "ms_per_step": round(e_elapsed / args.e2e_steps * 1e3, 2), "steps": args.e2e_steps,
                   "blocks_per_gpu": n_e,
                   "write_ms": round(tw["total_ms"], 2), "read_ms": round(tr["total_ms"], 2),
                   "write_sums_ms": {k: round(tw[k], 1) for k in ("h2d_ms", "kernel_ms", "d2h_ms")},
                   "read_sums_ms": {k: round(tr[k], 1) for k in ("h2d_ms", "kernel_ms", "d2h_ms")},
                   "api": "b2s_compress_packed + b2s_decompress_packed on pinned host arenas (b2s_host_alloc)"}
            if n_e != n:
                e2e["note"] = "pinned host memory did not allow the full %d blocks per rank; measured on %d" % (n, n_e)
            if world == 1 and n_e >= 200:
                # what ONE Spark task hands over: a map task's 200 partitions (commitAllPartitions), a reduce task's
                # 80 map outputs (S3ShuffleReader.read) — latency of a single synchronous C-ABI call, median of 5
                tws, trs = [], []
                for _ in range(6):
                    t = time.perf_counter()
                    w1 = c.compress_packed(CODEC, h_src.array, off_e[:200], ln_e[:200], h_cmp.array, LZ4_BLOCK,
                                           c.CHECKSUM_CRC32C)
                    tws.append(time.perf_counter() - t)
                    t = time.perf_counter()
                    r1 = c.decompress_packed(CODEC, h_cmp.array, w1["dst_off"][:80], w1["dst_len"][:80], h_out.array,
                                             c.CHECKSUM_CRC32C, sb_e[:81], w1["dst_len"][:80], w1["checksums"][:80])
                    trs.append(time.perf_counter() - t)
                    assert not w1["status"].any() and not r1["status"].any()
                tw1, tr1 = statistics.median(tws[1:]), statistics.median(trs[1:])
                e2e["task_sized_calls"] = {
                    "map_task": {"blocks": 200, "uncompressed_MB": round(200 * block_bytes / 1e6, 1),
                                 "ms": round(tw1 * 1e3, 2), "GBps": round(200 * block_bytes / tw1 / 1e9, 2)},
                    "reduce_task": {"blocks": 80, "uncompressed_MB": round(80 * block_bytes / 1e6, 1),
                                    "ms": round(tr1 * 1e3, 2), "GBps": round(80 * block_bytes / tr1 / 1e9, 2)}}

    # ------------------------------------------------------------------ CPU baseline beside it (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        sample_blocks = min(args.cpu_sample_blocks, n)
        sample = oracle.gen_terasort(0, sample_blocks * RECORDS_PER_BLOCK)
        cpu_arm(oracle, sample[: 200 * block_bytes], block_bytes, threads)
        v, detail = cpu_arm(oracle, sample, block_bytes, threads, repeats=2)
        v1, _ = cpu_arm(oracle, sample[: 64 * block_bytes], block_bytes, 1)
        cpu = {"value": round(v, 3), "unit": "GB/s", "cores": threads, "kind": "port",
               "sample": "%d of %d shuffle blocks (%.2f GiB), same generator/seed, best of 2" % (
                   sample_blocks, n, sample.size / 2**30),
               "single_core_GBps": round(v1, 3),
               "note": CPU_NOTES[args.codec]}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "compressed_ratio": round(ratio, 4), "gpu_launches": int(launches), "clocks": clocks,
                "roofline": roofline, "kernels": kernels}
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    rk.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
