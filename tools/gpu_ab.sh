# A/B of the compress pipeline generations x chunk sizes on the GPU box (one gpurun call)
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for P in 1 2; do for CH in 32768 65536 163840; do
  B2S_LZ4_PIPE=$P B2S_LZ4_CHUNK_BLOCKS=$CH timeout 600 python bench.py --no-e2e --no-cpu --steps 4 --warmup 3 > gpurun_out/ab_p${P}_c$CH.json 2> gpurun_out/ab_p${P}_c$CH.err
  echo "pipe=$P chunk=$CH"; python - <<PY
import json
d=json.load(open("gpurun_out/ab_p${P}_c$CH.json"))
print(d["value"], d["ms_per_step"], d["kernels"])
PY
  tail -2 gpurun_out/ab_p${P}_c$CH.err
done; done
B2S_LZ4_CHUNK_BLOCKS=65536 B2S_BENCH_BLOCKS=6400 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lz4_parse2" -s 3 -c 1 -o gpurun_out/r2b_parse -f python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_r2b.log 2>&1
tail -2 gpurun_out/ncu_r2b.log
timeout 600 python bench.py --steps 3 --no-cpu > gpurun_out/bench_e2e_r2b.json 2> gpurun_out/bench_e2e_r2b.err; tail -c 2500 gpurun_out/bench_e2e_r2b.json; tail -3 gpurun_out/bench_e2e_r2b.err
