compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests -m gpu -q -x -k "not full_size and not many_shuffle and not combineByKey" 2>&1 | tail -8 > gpurun_out/sanitizer_r1s.log; tail -8 gpurun_out/sanitizer_r1s.log
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
