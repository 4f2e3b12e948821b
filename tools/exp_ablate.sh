#!/bin/bash
# kernel ablation: DMV_DBG bit0 = skip global REDs, bit1 = skip phase C (Gram), bit2 = skip image gathers, bit3 = empty kernels
for P in ${PS:-16}; do for D in ${DS:-0 1 2 4 7 8}; do
  echo -n "P=$P dbg=$D: "
  DMV_DBG=$D python bench.py --steps 50 --warmup 3 --no-cpu-baseline --chunk $P 2>/tmp/err.txt > /tmp/out.json
  python - <<'PY'
import json
try:
    d=json.load(open('/tmp/out.json'))
    print('step_us',round(d['ms_per_step']*1e3,1),'point_us',round(d['roofline']['kernel_ms']*1e3,1),'e2e_us',round(d['e2e']['ms_per_step']*1e3,1))
except Exception as e:
    print('FAILED', open('/tmp/err.txt').read()[-300:])
PY
done; done
