# final 1-GPU validation: tests, default bench line (with e2e + cpu arm), reference arm, CRC A/B, launch list, ncu of the
# dominant kernel and of the CRC kernel
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
run() { # name env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > gpurun_out/fin_$name.json 2> gpurun_out/fin_$name.err
  echo "== $name"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/fin_$name.json")); print(d["value"], d["ms_per_step"], d["kernels"])
except Exception as e: print("FAILED", e)
PY
  tail -2 gpurun_out/fin_$name.err
}
run default
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 1500 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -c 600 gpurun_out/bench_final_ref.json
B2S_BENCH_BLOCKS=3200 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2z.csv python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/bench_under_ncu_r2z.log 2>&1
B2S_BENCH_BLOCKS=3200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lz4_match_kernel|crc_items" -s 6 -c 3 -o gpurun_out/r2z_top -f python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_r2z.log 2>&1
tail -2 gpurun_out/ncu_r2z.log
