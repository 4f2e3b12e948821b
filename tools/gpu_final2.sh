#!/bin/bash
# 2 GPUs, final code: whole GPU suite + bench at N = 2 (weak + strong + config 4 with the per-window chunk shape)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/f2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f2_pytest_gpu.log
grep -n "passed\|failed\|^FAILED\|^ERROR\|pytest rc" gpurun_out/f2_pytest_gpu.log | cut -c1-300 | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 500 --warmup 20 > gpurun_out/f2_bench_n2.json 2> gpurun_out/f2_bench_n2.err; cut -c1-330 gpurun_out/f2_bench_n2.json; tail -2 gpurun_out/f2_bench_n2.err | cut -c1-200
