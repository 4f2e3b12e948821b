#!/bin/bash
# quick GPU visit: parity tests, phase timeline, bench line
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 120 python tools/exp_phases.py 16 > gpurun_out/phases_p16.txt 2>&1
timeout 200 python bench.py --cpu-seconds 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/phases_p16.txt; cat gpurun_out/bench1.json; tail -3 gpurun_out/bench1.err
