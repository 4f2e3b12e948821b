#!/bin/bash
# round 2, visit b: first run of ba_fused_kernel (thread per residual, cooperative in-kernel reduction)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2b_gpu.txt 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2b_smoke.log; tail -5 gpurun_out/r2b_smoke.log | cut -c1-400
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2b_memcheck.log; tail -6 gpurun_out/r2b_memcheck.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_ba.py tests/test_golden.py tests/test_gpu_marg.py tests/test_gpu_host.py -m gpu -q > gpurun_out/r2b_pytest_ba.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_ba.log
tail -40 gpurun_out/r2b_pytest_ba.log | cut -c1-300
for P in 16 32; do
  timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --chunk $P > gpurun_out/r2b_bench_p$P.json 2> gpurun_out/r2b_bench_p$P.err
  cut -c1-600 gpurun_out/r2b_bench_p$P.json; tail -3 gpurun_out/r2b_bench_p$P.err
done
for g in 64 32; do
  DMV_L2_FETCH=$g timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:ba_fused -s 30 -c 3 --csv --log-file gpurun_out/r2b_l2fetch_$g.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_l2fetch_$g.log 2>&1
  tail -12 gpurun_out/r2b_l2fetch_$g.csv | cut -c1-300
done
DMV_L2_FETCH=32 timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/r2b_bench_l2fetch32.json 2> gpurun_out/r2b_bench_l2fetch32.err; cut -c1-300 gpurun_out/r2b_bench_l2fetch32.json
timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --points 8000 > gpurun_out/r2b_bench_c4.json 2> gpurun_out/r2b_bench_c4.err; cut -c1-300 gpurun_out/r2b_bench_c4.json
