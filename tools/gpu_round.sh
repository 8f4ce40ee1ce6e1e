# what a full check on the GPU box runs (the driver's round-end steps, in one place)
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
