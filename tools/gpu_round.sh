python -m pytest tests -m gpu -q -k "lz4 or parity or golden" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('LZ4', b['value'], b['kernels'])"
