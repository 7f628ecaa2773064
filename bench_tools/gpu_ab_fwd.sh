#!/bin/bash
# forward A/B: C3D_DEBUG values ($VARIANTS) x configs ($CFGS), bench.py device time + in-run parity against the C oracle
mkdir -p gpurun_out
rm -f gpurun_out/ab_*.log
for v in ${VARIANTS:-0}; do for cfg in ${CFGS:-c2 c2d2}; do
  C3D_DEBUG=$v timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-extras > gpurun_out/ab_${v}_$cfg.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${v}_$cfg.log").read().strip().splitlines()[-1])
    print("v=$v $cfg ms %.3f  parity %s" % (d["ms_per_step"], json.dumps(d.get("parity"))[:260]))
except Exception as e:
    print("v=$v $cfg FAILED", e)
PY
done; done
