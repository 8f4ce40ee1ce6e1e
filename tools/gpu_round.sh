python -m pytest tests/test_gpu_zstd.py -m gpu -q 2>&1 | tail -8
timeout 600 python bench.py --codec zstd --blocks 2000 --steps 1 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ZSTD 2000', b['value'],'ratio',b['compressed_ratio'],b['kernels'])"
timeout 900 python bench.py --codec zstd --steps 1 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ZSTD 16000', b['value'],'ratio',b['compressed_ratio'],b['kernels'])"
