#!/bin/bash
# 2 GPUs: the whole GPU suite (incl. sharded tests, initialiser kernel, drop-in on the CUDA library) + bench at N=2 (reordered exchange)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2n_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_gpu.log
grep -n "passed\|failed\|^FAILED\|^ERROR\|pytest rc" gpurun_out/r2n_pytest_gpu.log | cut -c1-300 | tail -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 20 --no-extras > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err; cut -c1-700 gpurun_out/r2n_bench_n2.json; tail -3 gpurun_out/r2n_bench_n2.err | cut -c1-300
