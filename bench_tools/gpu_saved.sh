#!/bin/bash
# Flip-inversion step: saved-forward path (default) against the recompute path (C3D_SAVE_FWD_GB=0).
mkdir -p gpurun_out; rm -f gpurun_out/inv.log
for d in 2 8; do
  for gb in 48 0; do
    D=$d TARGETS=16 C3D_SAVE_FWD_GB=$gb timeout 300 python bench_tools/bench_inversion.py >> gpurun_out/inv.log 2>&1
  done
done
cat gpurun_out/inv.log
