#!/bin/bash
# same-box A/B: cooperative launch vs plain launch of ba_fused_kernel
mkdir -p gpurun_out
for rep in 1 2 3; do
for v in coop noncoop; do
  case $v in coop) E=0;; noncoop) E=1;; esac
  DMV_BA_NONCOOP=$E timeout 300 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-extras > gpurun_out/ab_${v}_$rep.json 2> gpurun_out/ab_${v}_$rep.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${v}_$rep.json").read().strip().splitlines()[-1])
print("$v $rep", round(d["ms_per_step"]*1e3,2), "e2e", round(d["e2e"]["ms_per_step"]*1e3,2), "parity", d["parity"]["ok"])
PY
done
done
