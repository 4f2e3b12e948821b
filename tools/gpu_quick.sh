#!/bin/bash
# one short 1-GPU visit: the stream record (config 5) with its per-iteration split + the host tests that touch the solve
mkdir -p gpurun_out
timeout 300 python tools/bench_stream.py --keyframes 20 --its 10 --cpu-keyframes 0 > gpurun_out/q_stream.json 2> gpurun_out/q_stream.err; cut -c1-1500 gpurun_out/q_stream.json; tail -2 gpurun_out/q_stream.err
timeout 600 python -m pytest tests/test_gpu_host.py -m gpu -q 2>&1 | tail -3
