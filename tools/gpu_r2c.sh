#!/bin/bash
# round 2, visit c: ncu full capture of ba_fused_kernel (where do the 44 us go?), batched mode, new tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ba.py -m gpu -q -k "oob_and_prior or batched or reproducible" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -25 gpurun_out/r2c_pytest.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ba_fused_kernel -s 40 -c 1 -o gpurun_out/r2c_ba_fused python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2c_ncu.log 2>&1; tail -3 gpurun_out/r2c_ncu.log | cut -c1-200
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; cut -c1-300 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --chunk 32 --batch 16 > gpurun_out/r2c_bench_p32_b16.json 2> gpurun_out/r2c_bench_p32_b16.err; cut -c1-200 gpurun_out/r2c_bench_p32_b16.json
