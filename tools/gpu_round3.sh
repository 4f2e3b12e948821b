#!/bin/bash
# 2-GPU box: full GPU test suite (incl. the 2-rank tests), N=2 bench (fused exchange), secondary benches on GPU 0
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1000 --warmup 20 2> gpurun_out/bench_n2.err | grep "^{" > gpurun_out/bench_n2.json
cat gpurun_out/bench_n2.json | cut -c1-700
timeout 200 python tools/bench_coarse.py > gpurun_out/bench_coarse.json 2> gpurun_out/bench_coarse.err; cat gpurun_out/bench_coarse.json; tail -3 gpurun_out/bench_coarse.err
timeout 200 python tools/bench_coarse.py --levels 5 > gpurun_out/bench_coarse5.json 2>> gpurun_out/bench_coarse.err; cat gpurun_out/bench_coarse5.json
timeout 300 python tools/bench_stream.py > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err; cat gpurun_out/bench_stream.json; tail -5 gpurun_out/bench_stream.err
