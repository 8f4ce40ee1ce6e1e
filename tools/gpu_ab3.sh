# third A/B round: sub-chunk parallel parse (pipe 4, default)
run() { # name env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > gpurun_out/ab3_$name.json 2> gpurun_out/ab3_$name.err
  echo "== $name"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab3_$name.json")); print(d["value"], d["ms_per_step"], d["compressed_ratio"], d["kernels"])
except Exception as e: print("FAILED", e)
PY
  tail -2 gpurun_out/ab3_$name.err
}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
run p4_c32k B2S_LZ4_CHUNK_BLOCKS=32768
run p4_c64k B2S_LZ4_CHUNK_BLOCKS=65536
run p4_c8k B2S_LZ4_CHUNK_BLOCKS=8192
run p1_c32k B2S_LZ4_PIPE=1
B2S_BENCH_BLOCKS=3200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lz4_parse4|lz4_emit" -s 6 -c 2 -o gpurun_out/r2d_parse4 -f python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_r2d.log 2>&1
tail -2 gpurun_out/ncu_r2d.log
B2S_BENCH_BLOCKS=3200 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2d.csv python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/bench_under_ncu_r2d.log 2>&1
timeout 600 python bench.py --steps 3 --no-cpu > gpurun_out/bench_e2e_r2d.json 2> gpurun_out/bench_e2e_r2d.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_e2e_r2d.json")); print(d["value"], json.dumps(d["e2e"])[:1800])
PY
tail -3 gpurun_out/bench_e2e_r2d.err
