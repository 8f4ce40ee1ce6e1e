timeout 600 python tools/coalesce_bench.py > gpurun_out/coalesce.json 2> gpurun_out/coalesce.err; tail -3 gpurun_out/coalesce.err; cat gpurun_out/coalesce.json | tr -d '\n' | cut -c1-1600
