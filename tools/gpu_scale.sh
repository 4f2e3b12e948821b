#!/bin/bash
# Multi-GPU visit (gpurun --gpus 8): the sharded tests at 4 / 8 ranks, then bench.py at N = 2, 4, 8 (the driver's scaling sequence; N = 1 comes from the 1-GPU visit)
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/scale_gpus.txt 2>&1
nvidia-smi topo -m > $O/scale_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "4- or 8-" > $O/scale_pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/scale_pytest_multi.log
grep -n "passed\|failed\|^FAILED\|^ERROR\|pytest rc" $O/scale_pytest_multi.log | cut -c1-300 | tail -12
NG=$(nvidia-smi -L | wc -l)
for N in 2 4 8; do
  [ $N -gt $NG ] && continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N --steps 500 --warmup 20 > $O/scale_n$N.json 2> $O/scale_n$N.err
  cut -c1-330 $O/scale_n$N.json; tail -2 $O/scale_n$N.err | cut -c1-200
done
