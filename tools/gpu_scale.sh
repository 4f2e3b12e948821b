#!/bin/bash
# Multi-GPU visit (gpurun --gpus 8): the sharded tests at 2/4/8 ranks, then the driver's scaling sequence N = 1, 2, 4, 8 of bench.py
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/scale_gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/scale_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/scale_pytest_multi.log
grep -n "passed\|failed\|^FAILED\|^ERROR\|pytest rc" $O/scale_pytest_multi.log | cut -c1-300 | tail -12
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ $N -gt $NG ] && continue
  if [ $N = 1 ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu-baseline > $O/scale_n1.json 2> $O/scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N > $O/scale_n$N.json 2> $O/scale_n$N.err
  fi
  cut -c1-330 $O/scale_n$N.json; tail -2 $O/scale_n$N.err | cut -c1-200
done
