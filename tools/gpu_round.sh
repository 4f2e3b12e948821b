#!/bin/bash
# One GPU-box visit: parity tests, bench (ours + reference arm), ncu launch list, one full ncu capture of the dominant kernel.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt; lscpu | head -20 >> gpurun_out/gpu.txt
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 200 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ba_point -s 6 -c 2 -o gpurun_out/prof_ba_point python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench1.json; cat gpurun_out/bench_ref.json
