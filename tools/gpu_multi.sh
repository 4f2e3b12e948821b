#!/bin/bash
# usage: tools/gpu_multi.sh N  — the driver's multi-GPU launch line for bench.py (both arms) at N ranks
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 1000 --warmup 20 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
