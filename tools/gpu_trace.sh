#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_trace.py -m gpu -x -q > gpurun_out/pytest_trace.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_trace.log
tail -25 gpurun_out/pytest_trace.log | cut -c1-300
timeout 200 python tools/bench_trace.py > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err; cat gpurun_out/bench_trace.json; tail -3 gpurun_out/bench_trace.err
