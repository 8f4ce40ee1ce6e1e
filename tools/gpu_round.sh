for c in 16384 32768 65536 131072 400000; do echo "chunk $c"; B2S_LZ4_CHUNK_BLOCKS=$c python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -c 330; echo; done
