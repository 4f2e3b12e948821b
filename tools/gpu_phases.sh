#!/bin/bash
# in-kernel phase timeline of ba_point_kernel + a fresh 1-GPU bench line (clocks sampled by NVML)
mkdir -p gpurun_out
timeout 120 python tools/exp_phases.py 16 > gpurun_out/phases_p16.txt 2>&1
timeout 120 python tools/exp_phases.py 8 > gpurun_out/phases_p8.txt 2>&1
timeout 200 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err
cat gpurun_out/phases_p16.txt; tail -15 gpurun_out/phases_p8.txt; cat gpurun_out/bench1.json; tail -3 gpurun_out/bench1.err
