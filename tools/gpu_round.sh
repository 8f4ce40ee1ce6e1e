timeout 1200 python -m pytest tests/test_gpu_zstd.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --codec zstd --no-cpu --steps 3 > gpurun_out/bench_zstd7.json 2> gpurun_out/bench_zstd7.err; tail -c 300 gpurun_out/bench_zstd7.json
timeout 900 python tools/zstd_sweep.py > gpurun_out/zstd_sweep7.json 2> gpurun_out/zstd_sweep7.err; tail -3 gpurun_out/zstd_sweep7.err
