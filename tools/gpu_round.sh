python bench.py > gpurun_out/bench_r1z.json 2> gpurun_out/bench_r1z.err; tail -2 gpurun_out/bench_r1z.err
python bench.py --impl reference > gpurun_out/bench_ref_r1z.json 2>> gpurun_out/bench_r1z.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1z.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu_r1z.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lz4_match|lz4_copy" -c 2 -o gpurun_out/top_r1z python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_top_r1z.log 2>&1
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_r1z.json')); r=json.load(open('gpurun_out/bench_ref_r1z.json'))
print("value",b["value"],"e2e",b["e2e"]["value"],"cpu",b["cpu_baseline"]["value"],"ref arm",r["value"],"kernels",b["kernels"])
PY
