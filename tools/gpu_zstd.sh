# Zstandard: parity suite, the --codec zstd bench line, the block-size sweep (config 5), launch list
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --codec zstd --e2e-steps 3 > gpurun_out/bench_zstd.json 2> gpurun_out/bench_zstd.err; tail -c 1200 gpurun_out/bench_zstd.json; tail -3 gpurun_out/bench_zstd.err
timeout 900 python bench.py --config 5 > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; head -c 300 gpurun_out/bench_config5.json; tail -2 gpurun_out/bench_config5.err
B2S_BENCH_BLOCKS=3200 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_zstd.csv python bench.py --codec zstd --no-e2e --no-cpu --steps 1 --warmup 2 > gpurun_out/bench_under_ncu_zstd.log 2>&1
