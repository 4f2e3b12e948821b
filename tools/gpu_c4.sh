#!/bin/bash
mkdir -p gpurun_out
timeout 100 python bench.py --points 8000 --steps 300 --warmup 10 --cpu-seconds 4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; cut -c1-1600 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
