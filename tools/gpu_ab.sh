#!/bin/bash
# same-box A/B of the batched record: A = committed library, V3 = batch kernel (P = 32) compiled for 3 CTAs per SM
mkdir -p gpurun_out
for rep in 1 2; do
for v in A V3; do
  case $v in A) L="";; V3) L="$PWD/gpurun_variants/libV3.so";; esac
  DMVIO_B200_LIB=$L timeout 400 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --chunk 32 --batch 16 > gpurun_out/ab_${v}_$rep.json 2> gpurun_out/ab_${v}_$rep.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${v}_$rep.json").read().strip().splitlines()[-1])
rb=d.get("roofline_batched") or {}
print("$v $rep single", round(d["ms_per_step"]*1e3,2), "batched kernel_ms", rb.get("kernel_ms"), "frac", rb.get("frac"), "e2e_ms", rb.get("e2e_ms"))
PY
done
done
# auto chunk shape: BA parity tests + the default bench (config4_one_gpu should now run with 32-point chunks)
timeout 600 python -m pytest tests/test_gpu_ba.py tests/test_gpu_host.py -m gpu -q 2>&1 | tail -2
timeout 400 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/ab_autoP.json 2> gpurun_out/ab_autoP.err
python - <<PY
import json
d=json.loads(open("gpurun_out/ab_autoP.json").read().strip().splitlines()[-1])
c=d.get("config4_one_gpu",{})
print("autoP single", round(d["ms_per_step"]*1e3,2), d["config"].get("chunk_points"), "c4", round(c.get("ms_per_step",0)*1e3,2), round(c.get("value",0)/1e6), c.get("parity",{}).get("ok"))
PY
