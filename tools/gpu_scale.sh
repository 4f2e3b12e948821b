#!/bin/bash
# scaling run like the driver's: bench.py at N = 1, 2, 4, 8 ranks (whatever fits the box), both arms at N=1
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  if [ $N -gt $NG ]; then continue; fi
  if [ $N -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 1000 --warmup 20 --cpu-seconds 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 1000 --warmup 20 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 1000 --warmup 20 --exchange nccl > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err
  fi
  python - <<PY
import json
for f in ("gpurun_out/bench_n$N.json", "gpurun_out/bench_n${N}_nccl.json"):
    try:
        d = json.load(open(f)); print(f, "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"]*1e3,2), "us  value", round(d["value"]/1e6,1), "M/s  e2e", round(d["e2e"]["ms_per_step"]*1e3,2), "us", round(d["e2e"]["value"]/1e6,1), "M/s")
    except Exception as e:
        print(f, "FAILED", e)
PY
done
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n*.err | tail -10
