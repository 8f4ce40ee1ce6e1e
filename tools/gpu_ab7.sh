timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/duplex_probe.py 2>&1 | tail -2
B2S_LZ4D_TOKENS=2 B2S_BENCH_BLOCKS=3200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lz4_tokens_tma" -s 2 -c 1 -o gpurun_out/r2g_tokens_tma -f python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_r2g.log 2>&1
tail -2 gpurun_out/ncu_r2g.log
B2S_BENCH_BLOCKS=3200 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2g.csv python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/bench_under_ncu_r2g.log 2>&1
