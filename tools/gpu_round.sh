python -m pytest tests -m gpu -q -k "lz4 or parity or golden or host" 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print(b['value'], b['kernels'])"
ncu --set full --clock-control none -k regex:"lz4_tokens" -c 1 -o gpurun_out/tok_r1r python bench.py --blocks 4000 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_tok_r1r.log 2>&1
