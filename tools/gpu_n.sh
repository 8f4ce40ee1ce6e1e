# usage: bash tools/gpu_n.sh N [extra bench args]   — one torchrun bench line at N ranks into gpurun_out/bench_nN.json
N=$1; shift
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
fi
echo "rc=$?"; tail -c 3500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
