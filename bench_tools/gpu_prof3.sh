#!/bin/bash
for cfg in c2 c2d2; do
echo "== $cfg"
C3D_LIB=$PWD/bench_tools/_variants/libc3dpp_prof.so C3D_DEBUG=2 timeout 300 python bench.py --config $cfg --steps 1 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | grep "c3d prof" | tail -3
done
