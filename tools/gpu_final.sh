#!/bin/bash
# evidence visit: sanitizer on the small parity cases, ncu launch list + full captures of the three hot kernels, bench lines
mkdir -p gpurun_out
SMALL="tests/test_golden.py tests/test_gpu_coarse.py::test_edge_cases tests/test_gpu_ba.py::test_oob_and_prior_states tests/test_gpu_ba.py::test_resubstitute_and_step"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $SMALL -m gpu -x -q > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest $SMALL -m gpu -x -q > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -4 gpurun_out/sanitizer_memcheck.log; tail -4 gpurun_out/sanitizer_racecheck.log
timeout 200 python bench.py --cpu-seconds 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
timeout 200 python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/bench_ref.json 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ba_point -s 6 -c 2 -o gpurun_out/prof_ba_point python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ba_stitch -s 6 -c 2 -o gpurun_out/prof_ba_stitch python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ct_res_gs -s 10 -c 2 -o gpurun_out/prof_ct_res_gs python tools/bench_coarse.py --frames 3 --cpu-frames 1 > gpurun_out/b_ncu4.log 2>&1
timeout 200 python tools/bench_coarse.py > gpurun_out/bench_coarse.json 2> gpurun_out/bench_coarse.err
cat gpurun_out/bench1.json | cut -c1-900; cat gpurun_out/bench_ref.json | cut -c1-400; cat gpurun_out/bench_coarse.json
