#!/bin/bash
# same-box A/B: HEAD library vs working tree (+ the BA / marginalisation / batched parity tests on the working tree)
mkdir -p gpurun_out
for rep in 1 2 3; do
for v in HEAD NEW; do
  case $v in HEAD) L="$PWD/gpurun_variants/libHEAD.so";; NEW) L="";; esac
  DMVIO_B200_LIB=$L timeout 300 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-extras > gpurun_out/ab_${v}_$rep.json 2> gpurun_out/ab_${v}_$rep.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${v}_$rep.json").read().strip().splitlines()[-1])
print("$v $rep", round(d["ms_per_step"]*1e3,2), "e2e", round(d["e2e"]["ms_per_step"]*1e3,2), "parity", d["parity"]["ok"], d["parity"]["Hsc"])
PY
done
done
timeout 600 python -m pytest tests/test_gpu_ba.py tests/test_gpu_marg.py tests/test_golden.py -m gpu -q 2>&1 | tail -2
