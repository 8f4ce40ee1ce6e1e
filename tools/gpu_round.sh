python -m pytest tests/test_gpu_host_streams.py tests/test_gpu_host_shuffle.py -m gpu -x -q 2>&1 | tail -15
