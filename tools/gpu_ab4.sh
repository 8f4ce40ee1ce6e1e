# fourth A/B round: pipe 4 without the two-stream overlap; e2e copy interleaving knobs
run() { # name env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > gpurun_out/ab4_$name.json 2> gpurun_out/ab4_$name.err
  echo "== $name"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab4_$name.json")); print(d["value"], d["ms_per_step"], d["compressed_ratio"], d["kernels"])
except Exception as e: print("FAILED", e)
PY
  tail -2 gpurun_out/ab4_$name.err
}
e2e() { # name env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --no-cpu --e2e-steps 3 > gpurun_out/ab4_$name.json 2> gpurun_out/ab4_$name.err
  echo "== $name"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab4_$name.json")); e=d["e2e"]; print("e2e", e["value"], "serial", e["serial"]["value"], "write", e["write_ms"], "read", e["read_ms"], e["write_sums_ms"], e["read_sums_ms"], e.get("task_sized_calls"))
except Exception as e: print("FAILED", e)
PY
  tail -2 gpurun_out/ab4_$name.err
}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run p4_serial B2S_OVERLAP=0
run p4_serial_c64k B2S_OVERLAP=0 B2S_LZ4_CHUNK_BLOCKS=65536
run p1_serial B2S_OVERLAP=0 B2S_LZ4_PIPE=1
env B2S_OVERLAP=0 timeout 600 python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 --codec snappy > gpurun_out/ab4_snappy.json 2>&1; tail -c 400 gpurun_out/ab4_snappy.json
e2e base
e2e piece16 B2S_COPY_PIECE_MB=16
e2e piece16_p1 B2S_COPY_PIECE_MB=16 B2S_LZ4_PIPE=1
e2e chunk64 B2S_HOST_CHUNK_MB=64
e2e chunk64_slots8 B2S_HOST_CHUNK_MB=64 B2S_SLOTS=8
