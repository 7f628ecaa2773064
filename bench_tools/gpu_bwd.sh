#!/bin/bash
# backward-path session: gradient tests, then A/B of a tuning option on the 32-image inversion step (D = 2, 8)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/inv_*.log
run() { name=$1; shift; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_bwd python -m pytest tests/test_gpu_backward.py -q -m gpu -x
for v in ${VARIANTS:-0}; do for d in 2 8; do
  C3D_DEBUG=$v D=$d TARGETS=16 STEPS=50 run inv_v${v}_d$d python bench_tools/bench_inversion.py
done; done
cat gpurun_out/summary.txt
tail -n 4 gpurun_out/pytest_bwd.log
for f in gpurun_out/inv_v*.log; do echo $f; grep -h "ms_fwd_bwd\|full_step" $f | cut -c1-200; done
