#!/bin/bash
# 2 GPUs: sharded tests + bench at N=2 (weak headline + strong + config 4 side records)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2j_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_ba.py -m gpu -q -k "sharded or two_devices" > gpurun_out/r2j_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_multi.log
tail -25 gpurun_out/r2j_pytest_multi.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 20 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; cut -c1-1500 gpurun_out/r2j_bench_n2.json; tail -5 gpurun_out/r2j_bench_n2.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 300 --warmup 20 --exchange nccl --no-extras > gpurun_out/r2j_bench_n2_nccl.json 2> gpurun_out/r2j_bench_n2_nccl.err; cut -c1-400 gpurun_out/r2j_bench_n2_nccl.json
