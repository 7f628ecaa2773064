#!/bin/bash
run() { echo "== debug=$1 cfg=$2"; C3D_LIB=$PWD/bench_tools/_variants/libc3dpp_0x00.so C3D_STAGGER=0 C3D_DEBUG=$1 timeout 300 python bench.py --config $2 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ms', round(d['ms_per_step'], 3), 'min', round(d['ms_per_step_min'], 3), 'TF', round(d['roofline']['achieved'], 1), 'clk', d['clocks'])
    elif 'rror' in l: print(l.strip())
"; }
run 0 c2
run 1 c2
run 0 c2d2
run 1 c2d2
