#!/bin/bash
# 2 GPUs: sharded tests (warp-collective exchange) + bench at N=2 + drop-in tests on the CUDA library
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_ba.py -m gpu -q -k "sharded or two_devices" > gpurun_out/r2m_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_multi.log
tail -12 gpurun_out/r2m_pytest_multi.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 20 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; cut -c1-1800 gpurun_out/r2m_bench_n2.json; tail -5 gpurun_out/r2m_bench_n2.err | cut -c1-300
timeout 900 python -m pytest tests/test_dropin.py -m gpu -q > gpurun_out/r2m_pytest_dropin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_dropin.log
tail -30 gpurun_out/r2m_pytest_dropin.log | cut -c1-400
