python bench.py --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('NSLOT6', b['value']); print(b['e2e'])"
B2S_HOST_CHUNK_MB=128 python bench.py --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('NSLOT6 chunk128', b['e2e'])"
