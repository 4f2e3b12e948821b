#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ba.py tests/test_gpu_multi.py tests/test_gpu_marg.py tests/test_golden.py -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
grep -n "passed\|failed\|^FAILED\|^ERROR\|pytest rc" gpurun_out/r2q_pytest.log | cut -c1-300 | tail -12
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err; cut -c1-330 gpurun_out/r2q_bench_n1.json; tail -3 gpurun_out/r2q_bench_n1.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 20 --no-extras > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err; cut -c1-330 gpurun_out/r2q_bench_n2.json; tail -3 gpurun_out/r2q_bench_n2.err | cut -c1-300
