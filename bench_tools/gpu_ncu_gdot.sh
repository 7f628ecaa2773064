#!/bin/bash
# ncu --set full of the two small backward kernels of one flip-inversion step (32 images, D=2).
mkdir -p gpurun_out
D=2 TARGETS=16 STEPS=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gdot_kernel|composite_bwd_kernel' -s 6 -c 2 -f -o gpurun_out/prof_gdot python bench_tools/bench_inversion.py > gpurun_out/ncu_gdot.log 2>&1
tail -n 4 gpurun_out/ncu_gdot.log
