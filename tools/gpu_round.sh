timeout 1200 python -m pytest tests/test_gpu_zstd.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/zstd_sweep.py > gpurun_out/zstd_sweep5.json 2> gpurun_out/zstd_sweep5.err; tail -3 gpurun_out/zstd_sweep5.err
timeout 600 python bench.py --codec zstd --no-cpu --no-e2e --steps 3 > gpurun_out/bench_zstd5.json 2> gpurun_out/bench_zstd5.err; tail -c 300 gpurun_out/bench_zstd5.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_zstd.py -m gpu -x -q > gpurun_out/memcheck_zstd.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_zstd.log
