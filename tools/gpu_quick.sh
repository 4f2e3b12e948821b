#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_host.py tests/test_gpu_marg.py -m gpu -q > gpurun_out/pytest_host.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_host.log
tail -30 gpurun_out/pytest_host.log | cut -c1-300
timeout 60 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_host.py -m gpu -x -q -k "turnover" > gpurun_out/sanitizer_host.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer_host.log; tail -4 gpurun_out/sanitizer_host.log | cut -c1-300
