#!/bin/bash
# 2 GPUs: sharded adapter tests + initialiser kernel tests + drop-in (initialiser)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_init.py tests/test_dropin.py -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
grep -n "passed\|failed\|^FAILED\|^ERROR\|pytest rc\|^E  " gpurun_out/r2o_pytest.log | cut -c1-300 | tail -60
