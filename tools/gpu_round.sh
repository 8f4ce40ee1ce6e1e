python -m pytest tests/test_gpu_configs.py -m gpu -q 2>&1 | tail -6
