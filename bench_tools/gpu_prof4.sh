#!/bin/bash
export C3D_FWD=pair
C3D_LIB=$PWD/bench_tools/_variants/libc3dpp_0x00.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "bf16 or render" 2>&1 | tail -3
for egw in 8 4; do for m in 0x00 0x92; do
echo "== egw $egw mask $m"
for cfg in c2 c2d2; do
C3D_EGW=$egw C3D_LIB=$PWD/bench_tools/_variants/libc3dpp_$m.so timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ms', round(d['ms_per_step'], 3), 'min', round(d['ms_per_step_min'], 3), 'TF', round(d['roofline']['achieved'], 1), 'clk', d['clocks'])
    elif 'rror' in l or 'c3d' in l: print(l.strip())
"; done; done; done
