#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_host.py tests/test_gpu_coarse.py -m gpu -x -q > gpurun_out/pytest_ct.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ct.log
tail -15 gpurun_out/pytest_ct.log
timeout 200 python tools/bench_coarse.py > gpurun_out/bench_coarse.json 2> gpurun_out/bench_coarse.err; cat gpurun_out/bench_coarse.json; tail -3 gpurun_out/bench_coarse.err
timeout 200 python tools/bench_coarse.py --levels 5 > gpurun_out/bench_coarse5.json 2>> gpurun_out/bench_coarse.err; cat gpurun_out/bench_coarse5.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
