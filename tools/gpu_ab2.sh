# second A/B round: event-driven parse (pipe 3) and the TMA-ring token walk (tokens 2), tests under both
run() { # name env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > gpurun_out/ab2_$name.json 2> gpurun_out/ab2_$name.err
  echo "== $name"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab2_$name.json")); print(d["value"], d["ms_per_step"], d["kernels"])
except Exception as e: print("FAILED", e)
PY
  tail -2 gpurun_out/ab2_$name.err
}
B2S_LZ4_PIPE=3 B2S_LZ4D_TOKENS=2 B2S_LZ4_MATCH_DEPTH=2 B2S_LZ4D_COPYGROUP=4 timeout 900 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -22
run p2_c64k B2S_LZ4_PIPE=2 B2S_LZ4_CHUNK_BLOCKS=65536
run p3_c32k B2S_LZ4_PIPE=3 B2S_LZ4_CHUNK_BLOCKS=32768
run p3_c64k B2S_LZ4_PIPE=3 B2S_LZ4_CHUNK_BLOCKS=65536
run p3_c160k B2S_LZ4_PIPE=3 B2S_LZ4_CHUNK_BLOCKS=163840
run p3_c160k_depth2 B2S_LZ4_PIPE=3 B2S_LZ4_CHUNK_BLOCKS=163840 B2S_LZ4_MATCH_DEPTH=2
run p3_c160k_cg4 B2S_LZ4_PIPE=3 B2S_LZ4_CHUNK_BLOCKS=163840 B2S_LZ4D_COPYGROUP=4
run p3_c160k_tok2 B2S_LZ4_PIPE=3 B2S_LZ4_CHUNK_BLOCKS=163840 B2S_LZ4D_TOKENS=2
run p3_c160k_tok2_d160k B2S_LZ4_PIPE=3 B2S_LZ4_CHUNK_BLOCKS=163840 B2S_LZ4D_TOKENS=2 B2S_LZ4D_CHUNK_BLOCKS=163840
run p3_c160k_tok1_d160k B2S_LZ4_PIPE=3 B2S_LZ4_CHUNK_BLOCKS=163840 B2S_LZ4D_TOKENS=1 B2S_LZ4D_CHUNK_BLOCKS=163840
B2S_LZ4_PIPE=3 B2S_LZ4D_TOKENS=2 B2S_BENCH_BLOCKS=3200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lz4_parse2|lz4_tokens_tma|lz4_match2" -s 9 -c 3 -o gpurun_out/r2c_kernels -f python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_r2c.log 2>&1
tail -2 gpurun_out/ncu_r2c.log
B2S_LZ4_PIPE=3 B2S_LZ4D_TOKENS=1 B2S_BENCH_BLOCKS=3200 timeout 900 ncu --set full --clock-control none -k regex:"lz4_tokens_kernel" -s 2 -c 1 -o gpurun_out/r2c_tokens_v1 -f python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_r2c1.log 2>&1
