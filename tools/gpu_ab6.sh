# sixth A/B round: deferred long-match resolution in the (generation 1) match kernel; decode chunk sizes
run() { # name env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > gpurun_out/ab6_$name.json 2> gpurun_out/ab6_$name.err
  echo "== $name"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab6_$name.json")); print(d["value"], d["ms_per_step"], d["compressed_ratio"], d["kernels"])
except Exception as e: print("FAILED", e)
PY
  tail -2 gpurun_out/ab6_$name.err
}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run defer1
run defer0 B2S_LZ4_MATCH_DEFER=0
run defer1_d32k B2S_LZ4D_CHUNK_BLOCKS=32768
run defer1_d16k B2S_LZ4D_CHUNK_BLOCKS=16384
run defer1_c64k B2S_LZ4_CHUNK_BLOCKS=65536
B2S_BENCH_BLOCKS=3200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lz4_match_kernel" -s 3 -c 1 -o gpurun_out/r2f_match_defer -f python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_r2f.log 2>&1
tail -2 gpurun_out/ncu_r2f.log
