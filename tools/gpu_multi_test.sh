#!/bin/bash
# 2-rank GPU tests (peer-memory exchange + NCCL) and the N-rank bench line with both exchanges
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_gpu_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_multi.log
tail -15 gpurun_out/pytest_gpu_multi.log
for X in p2p nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 1000 --warmup 20 --exchange $X > gpurun_out/bench_n${N}_$X.json 2> gpurun_out/bench_n${N}_$X.err
cat gpurun_out/bench_n${N}_$X.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n${N}_$X.err | tail -5
done
