#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -x -q 2>&1 | tail -3
for cfg in c2 c2d2 c4 c1 c3; do
echo "== $cfg"; timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ms', round(d['ms_per_step'], 3), 'min', round(d['ms_per_step_min'], 3), 'TF', round(d['roofline']['achieved'], 1), 'clk', d['clocks'])
    elif 'rror' in l or 'c3d' in l: print(l.strip())
"; done
