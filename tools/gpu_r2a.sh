#!/bin/bash
# round 2, visit a: GPU tests with the unconditional parity assertions on the round-1 kernels + L2 fetch-granularity experiment
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_gpu.log
tail -15 gpurun_out/r2a_pytest_gpu.log | cut -c1-400
for g in 64 32; do
  DMV_L2_FETCH=$g timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:ba_point -s 30 -c 3 --csv --log-file gpurun_out/r2a_l2fetch_$g.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_l2fetch_$g.log 2>&1
  tail -4 gpurun_out/r2a_l2fetch_$g.csv | cut -c1-300
done
for g in 64 32; do
  DMV_L2_FETCH=$g timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/r2a_bench_l2fetch_$g.json 2> gpurun_out/r2a_bench_l2fetch_$g.err
  cut -c1-400 gpurun_out/r2a_bench_l2fetch_$g.json
done
