#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_coarse.py tests/test_gpu_host.py tests/test_golden.py tests/test_gpu_marg.py -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -8 gpurun_out/r2h_pytest.log | cut -c1-300
timeout 200 python tools/bench_coarse.py --frames 300 > gpurun_out/r2h_bench_coarse_cluster_L4.json 2> gpurun_out/r2h_bench_coarse_cluster_L4.err; cut -c1-900 gpurun_out/r2h_bench_coarse_cluster_L4.json; tail -2 gpurun_out/r2h_bench_coarse_cluster_L4.err
timeout 200 python tools/bench_coarse.py --frames 300 --levels 5 > gpurun_out/r2h_bench_coarse_cluster_L5.json 2> gpurun_out/r2h_bench_coarse_cluster_L5.err; cut -c1-400 gpurun_out/r2h_bench_coarse_cluster_L5.json
