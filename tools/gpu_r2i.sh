#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trace.py tests/test_gpu_host.py -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -25 gpurun_out/r2i_pytest.log | cut -c1-300
timeout 300 python tools/bench_trace.py > gpurun_out/r2i_bench_trace.json 2> gpurun_out/r2i_bench_trace.err; cut -c1-900 gpurun_out/r2i_bench_trace.json; tail -2 gpurun_out/r2i_bench_trace.err
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_trace.py -m gpu -q -x > gpurun_out/r2i_memcheck_trace.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2i_memcheck_trace.log; tail -4 gpurun_out/r2i_memcheck_trace.log | cut -c1-200
