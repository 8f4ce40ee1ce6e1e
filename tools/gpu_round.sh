timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --codec snappy --steps 3 > gpurun_out/bench_snappy8.json 2> gpurun_out/bench_snappy8.err; tail -c 400 gpurun_out/bench_snappy8.json
timeout 900 python bench.py --codec zstd --steps 3 > gpurun_out/bench_zstd8.json 2> gpurun_out/bench_zstd8.err; tail -c 400 gpurun_out/bench_zstd8.json
