# final 1-GPU validation of a round: tests, smoke, the default bench line (with e2e + cpu arm), the reference arm, the
# other driver-visible configs, the launch list and an ncu capture of the dominant kernels.  Everything lands in
# gpurun_out/ (copied into profiles/ by hand afterwards).
tag=${1:-r2z}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 1800 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -c 600 gpurun_out/bench_final_ref.json
timeout 900 python bench.py --config 4 --e2e-steps 3 > gpurun_out/bench_final_config4.json 2> gpurun_out/bench_final_config4.err; tail -c 700 gpurun_out/bench_final_config4.json; tail -2 gpurun_out/bench_final_config4.err
timeout 900 python bench.py --config 5 > gpurun_out/bench_final_config5.json 2> gpurun_out/bench_final_config5.err; head -c 600 gpurun_out/bench_final_config5.json; tail -2 gpurun_out/bench_final_config5.err
B2S_BENCH_BLOCKS=3200 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/bench_under_ncu_$tag.log 2>&1
B2S_BENCH_BLOCKS=3200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lz4_tokens_kernel|lz4_copy_kernel|lz4_parse_kernel|lz4_emit_kernel" -s 8 -c 4 -o gpurun_out/${tag}_rest -f python bench.py --no-e2e --no-cpu --steps 1 --warmup 3 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
