# fifth A/B round: parse4 with cooperative (coalesced) match measurement
run() { # name env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > gpurun_out/ab5_$name.json 2> gpurun_out/ab5_$name.err
  echo "== $name"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab5_$name.json")); print(d["value"], d["ms_per_step"], d["compressed_ratio"], d["kernels"])
except Exception as e: print("FAILED", e)
PY
  tail -2 gpurun_out/ab5_$name.err
}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run p4coop
run p4coop_serial B2S_OVERLAP=0
run p4coop_c64k B2S_LZ4_CHUNK_BLOCKS=65536
B2S_BENCH_BLOCKS=3200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lz4_parse4" -s 3 -c 1 -o gpurun_out/r2e_parse4 -f python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_r2e.log 2>&1
tail -2 gpurun_out/ncu_r2e.log
e2e() { # name env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --no-cpu --e2e-steps 3 > gpurun_out/ab5_$name.json 2> gpurun_out/ab5_$name.err
  echo "== $name"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab5_$name.json")); e=d["e2e"]; print("e2e", e["value"], "serial", e["serial"]["value"], "write", e["write_ms"], "read", e["read_ms"], e["write_sums_ms"], e["read_sums_ms"], e.get("task_sized_calls"))
except Exception as e: print("FAILED", e)
PY
  tail -2 gpurun_out/ab5_$name.err
}
e2e prio_lag2
e2e prio_lag1 B2S_READ_LAG=1
e2e noprio_lag2 B2S_READ_PRIORITY=0
e2e prio_lag3_slots8 B2S_READ_LAG=3 B2S_SLOTS=8
