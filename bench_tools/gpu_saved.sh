#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -q -m gpu -x 2>&1 | tail -30 > gpurun_out/bwd_tests.log
for d in 2 8; do
  for gb in 48 0; do
    echo "D=$d C3D_SAVE_FWD_GB=$gb" >> gpurun_out/inv.log
    D=$d TARGETS=16 C3D_SAVE_FWD_GB=$gb timeout 300 python bench_tools/bench_inversion.py >> gpurun_out/inv.log 2>&1
  done
done
tail -5 gpurun_out/bwd_tests.log; cat gpurun_out/inv.log
