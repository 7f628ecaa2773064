#!/bin/bash
mkdir -p gpurun_out
export C3D_FWD=pair
C3D_LIB=$PWD/bench_tools/_variants/libc3dpp_0x00.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_forward_pair -s 2 -c 1 -f -o gpurun_out/prof_pair python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_pair.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_pair.log
