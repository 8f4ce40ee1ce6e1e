python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_r1n.json'))
print("value",b["value"],"kernels",b["kernels"]); print("e2e",b.get("e2e"))
PY
tail -3 gpurun_out/bench_r1n.err
for mb in 64 128 512; do echo "host chunk $mb"; B2S_HOST_CHUNK_MB=$mb python bench.py --steps 1 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print(b['e2e'])"; done
