python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 900 python bench.py --codec zstd --steps 1 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ZSTD 16000', b['value'],'ratio',b['compressed_ratio'],b['kernels'])"
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('LZ4', b['value'], b['kernels'])"
