#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ba.py tests/test_gpu_marg.py -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log | cut -c1-300
for P in 16 32; do
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --chunk $P > gpurun_out/r2d_bench_p$P.json 2> gpurun_out/r2d_bench_p$P.err; cut -c1-250 gpurun_out/r2d_bench_p$P.json; tail -2 gpurun_out/r2d_bench_p$P.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ba_fused_kernel -s 40 -c 1 -o gpurun_out/r2d_ba_fused python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2d_ncu.log 2>&1; tail -2 gpurun_out/r2d_ncu.log | cut -c1-200
