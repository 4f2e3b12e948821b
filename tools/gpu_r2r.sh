#!/bin/bash
# A/B/C on ONE box: A = HEAD (two-pass Schur tiles, warp-collective exchange), B = per-chunk Gram partials (float4 items), C = HEAD with the fused kernel of 79f278e
mkdir -p gpurun_out
for rep in 1 2; do
for v in A B C; do
  case $v in A) L="";; B) L="gpurun_variants/libB_gram.so";; C) L="gpurun_variants/libC_oldxchg.so";; esac
  DMVIO_B200_LIB=${L:+$PWD/$L} timeout 300 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-extras > gpurun_out/r2r_${v}_$rep.json 2> gpurun_out/r2r_${v}_$rep.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2r_${v}_$rep.json").read().strip().splitlines()[-1])
print("$v $rep", round(d["ms_per_step"]*1e3,2), "e2e", round(d["e2e"]["ms_per_step"]*1e3,2), "parity", d["parity"]["ok"])
PY
done
done
