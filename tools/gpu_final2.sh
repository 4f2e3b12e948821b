#!/bin/bash
# single evidence visit (tight GPU budget): full GPU suite, smoke, bench lines, ncu launch list + full capture, secondary benches.
# Every step has its own timeout; later steps are skipped once the deadline has passed.
mkdir -p gpurun_out
T0=$(date +%s); DEADLINE=${DEADLINE:-330}
left() { [ $(( $(date +%s) - T0 )) -lt $DEADLINE ]; }
timeout 170 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log | cut -c1-300
left && { timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log; }
left && { timeout 120 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; cut -c1-1200 gpurun_out/bench1.json; }
left && { timeout 100 python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/bench_ref.json 2>&1; tail -1 gpurun_out/bench_ref.json | cut -c1-500; }
left && timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
left && timeout 120 ncu --set full --clock-control none --import-source on -k regex:ba_point -s 6 -c 2 -o gpurun_out/prof_ba_point -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
left && { timeout 60 python tools/bench_marg.py > gpurun_out/bench_marg.json 2> gpurun_out/bench_marg.err; cat gpurun_out/bench_marg.json; }
left && { timeout 80 python tools/bench_coarse.py > gpurun_out/bench_coarse.json 2> gpurun_out/bench_coarse.err; cat gpurun_out/bench_coarse.json | cut -c1-600; }
left && { timeout 60 python tools/bench_trace.py > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err; cat gpurun_out/bench_trace.json | cut -c1-400; }
left && { timeout 80 python tools/bench_stream.py > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err; cat gpurun_out/bench_stream.json | cut -c1-400; }
echo "elapsed $(( $(date +%s) - T0 )) s"
