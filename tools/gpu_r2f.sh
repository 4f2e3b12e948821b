#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_coarse.py tests/test_gpu_host.py tests/test_golden.py -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -12 gpurun_out/r2f_pytest.log | cut -c1-300
timeout 200 python tools/bench_coarse.py --frames 200 > gpurun_out/r2f_bench_coarse_cluster_L4.json 2> gpurun_out/r2f_bench_coarse_cluster_L4.err; cut -c1-700 gpurun_out/r2f_bench_coarse_cluster_L4.json; tail -2 gpurun_out/r2f_bench_coarse_cluster_L4.err
DMV_CT_GRID=1 timeout 200 python tools/bench_coarse.py --frames 200 > gpurun_out/r2f_bench_coarse_grid_L4.json 2> gpurun_out/r2f_bench_coarse_grid_L4.err; cut -c1-700 gpurun_out/r2f_bench_coarse_grid_L4.json
timeout 200 python tools/bench_coarse.py --frames 200 --levels 5 > gpurun_out/r2f_bench_coarse_cluster_L5.json 2> gpurun_out/r2f_bench_coarse_cluster_L5.err; cut -c1-700 gpurun_out/r2f_bench_coarse_cluster_L5.json
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/r2f_bench_p16.json 2> gpurun_out/r2f_bench_p16.err; cut -c1-250 gpurun_out/r2f_bench_p16.json; tail -2 gpurun_out/r2f_bench_p16.err
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_coarse.py -m gpu -q -x -k "track or parity" > gpurun_out/r2f_memcheck_ct.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2f_memcheck_ct.log; tail -4 gpurun_out/r2f_memcheck_ct.log | cut -c1-200
