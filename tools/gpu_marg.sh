#!/bin/bash
# marginalisation launch: parity tests, memcheck of one case, and a quick bench line to confirm the production kernel is unchanged
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_marg.py tests/test_gpu_host.py -m gpu -x -q > gpurun_out/pytest_marg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_marg.log
tail -30 gpurun_out/pytest_marg.log | cut -c1-400
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_marg.py -m gpu -x -q -k "nf4 or empty" > gpurun_out/sanitizer_marg.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer_marg.log
tail -6 gpurun_out/sanitizer_marg.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/bench_marg_check.json 2> gpurun_out/bench_marg_check.err; cat gpurun_out/bench_marg_check.json | cut -c1-1500
