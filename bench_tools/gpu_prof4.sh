#!/bin/bash
for dbg in 2 14; do
echo "== debug $dbg"
C3D_LIB=$PWD/bench_tools/_variants/libc3dpp_prof.so C3D_DEBUG=$dbg timeout 300 python bench.py --config c2 --steps 1 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | grep "c3d timeline" | tail -3
done
