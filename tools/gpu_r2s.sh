#!/bin/bash
# L2 fetch granularity A/B on ONE box (A = HEAD, B = per-chunk Gram partials)
mkdir -p gpurun_out
for rep in 1 2; do
for v in A32 A64 A128 B32 B64; do
  case $v in A*) L="";; B*) L="$PWD/gpurun_variants/libB_gram.so";; esac
  DMV_L2_FETCH=${v:1} DMVIO_B200_LIB=$L timeout 300 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-extras > gpurun_out/r2s_${v}_$rep.json 2> gpurun_out/r2s_${v}_$rep.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2s_${v}_$rep.json").read().strip().splitlines()[-1])
print("$v $rep", round(d["ms_per_step"]*1e3,2), "e2e", round(d["e2e"]["ms_per_step"]*1e3,2), "parity", d["parity"]["ok"])
PY
done
done
