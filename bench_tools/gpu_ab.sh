#!/bin/bash
# A/B: slot stagger and FMA-pipe sine fraction on the forward kernel (c2 = D8, c2d2 = D2)
mkdir -p gpurun_out
run() { echo "== lib=$1 stagger=$2 cfg=$3"; C3D_LIB=$PWD/bench_tools/_variants/libc3dpp_$1.so C3D_STAGGER=$2 timeout 300 python bench.py --config $3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ms', round(d['ms_per_step'], 3), 'min', round(d['ms_per_step_min'], 3), 'TF', round(d['roofline']['achieved'], 1), 'clk', d['clocks'])
    elif 'rror' in l: print(l.strip())
"; }
for cfg in c2 c2d2; do
  run 0x00 0 $cfg
  run 0x00 5 $cfg
  run 0x00 2 $cfg
  run 0x88 0 $cfg
  run 0x88 5 $cfg
  run 0xAA 5 $cfg
  run 0xAA 2 $cfg
done
C3D_LIB=$PWD/bench_tools/_variants/libc3dpp_0x88.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
