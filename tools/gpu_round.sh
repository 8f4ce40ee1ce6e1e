python bench.py > gpurun_out/bench_r1q.json 2> gpurun_out/bench_r1q.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_r1q.json'))
print("LZ4 value",b["value"],"ms/step",b["ms_per_step"],"kernels",b["kernels"]); print("e2e",b.get("e2e")); print("cpu",b.get("cpu_baseline")); print("roofline",b["roofline"]["frac"], b["roofline"]["compress_step"]["frac"], b["roofline"]["decompress_step"]["frac"])
PY
tail -2 gpurun_out/bench_r1q.err
python bench.py --codec snappy --steps 2 --warmup 3 --no-cpu 2>&1 | python -c "import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SNAPPY value',b['value'],'ratio',b['compressed_ratio'],b['kernels']); print(b.get('e2e'))"
timeout 600 python bench.py --codec zstd --blocks 2000 --steps 1 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ZSTD value',b['value'],'ratio',b['compressed_ratio'],b['kernels'])"
