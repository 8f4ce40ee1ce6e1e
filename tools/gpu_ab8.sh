e2e() { # name env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --no-cpu --e2e-steps 3 > gpurun_out/ab8_$name.json 2> gpurun_out/ab8_$name.err
  echo "== $name"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab8_$name.json")); e=d["e2e"]; print("e2e", e["value"], "serial", e["serial"]["value"], "write", e["write_ms"], "read", e["read_ms"], e["write_sums_ms"], e["read_sums_ms"], e.get("task_sized_calls"))
except Exception as e: print("FAILED", e)
PY
  tail -2 gpurun_out/ab8_$name.err
}
e2e chunk512_s4 B2S_HOST_CHUNK_MB=512 B2S_SLOTS=4
e2e chunk1024_s3 B2S_HOST_CHUNK_MB=1024 B2S_SLOTS=3
e2e chunk512_s6 B2S_HOST_CHUNK_MB=512
