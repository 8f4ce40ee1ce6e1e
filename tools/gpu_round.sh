timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 700 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; head -c 300 gpurun_out/bench_final_ref.json
