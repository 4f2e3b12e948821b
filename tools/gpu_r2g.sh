#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -12 gpurun_out/r2g_pytest.log | cut -c1-300
timeout 200 python tools/bench_coarse.py --frames 200 > gpurun_out/r2g_bench_coarse_cluster_L4.json 2> gpurun_out/r2g_bench_coarse_cluster_L4.err; cut -c1-700 gpurun_out/r2g_bench_coarse_cluster_L4.json; tail -2 gpurun_out/r2g_bench_coarse_cluster_L4.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ct_track_cluster -s 3 -c 1 -o gpurun_out/r2g_ct_cluster python tools/bench_coarse.py --frames 3 --cpu-frames 1 > gpurun_out/r2g_ncu_ct.log 2>&1; tail -2 gpurun_out/r2g_ncu_ct.log | cut -c1-200
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/r2g_bench_p16.json 2> gpurun_out/r2g_bench_p16.err; cut -c1-250 gpurun_out/r2g_bench_p16.json; tail -2 gpurun_out/r2g_bench_p16.err
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --chunk 32 --batch 16 > gpurun_out/r2g_bench_p32_b16.json 2> gpurun_out/r2g_bench_p32_b16.err; cut -c1-250 gpurun_out/r2g_bench_p32_b16.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ba_fused_kernel -s 40 -c 1 -o gpurun_out/r2g_ba_fused python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2g_ncu.log 2>&1; tail -2 gpurun_out/r2g_ncu.log | cut -c1-200
