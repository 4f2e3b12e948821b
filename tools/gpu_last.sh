#!/bin/bash
# last visit of the round: re-run the whole GPU suite on the final code, memcheck the keyframe-bookkeeping paths, streaming bench,
# launch list of the marginalisation launch
mkdir -p gpurun_out
T0=$(date +%s); DEADLINE=${DEADLINE:-150}
left() { [ $(( $(date +%s) - T0 )) -lt $DEADLINE ]; }
timeout 120 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log | cut -c1-300
left && { timeout 70 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_host.py -m gpu -x -q -k "finish or flag or turnover" > gpurun_out/sanitizer_host.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer_host.log; tail -5 gpurun_out/sanitizer_host.log | cut -c1-300; }
left && { timeout 60 python tools/bench_stream.py > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err; cat gpurun_out/bench_stream.json | cut -c1-500; }
left && { timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_marg.csv python tools/bench_marg.py --reps 3 > gpurun_out/b_ncu_marg.log 2>&1; grep -c ba_point gpurun_out/launches_marg.csv; }
echo "elapsed $(( $(date +%s) - T0 )) s"
