#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run bwd_tests python -m pytest tests/test_gpu_backward.py -q --timeout 300 -s
run fwd_tests python -m pytest tests/test_gpu_parity.py -q --timeout 300
cat gpurun_out/summary.txt; tail -n 30 gpurun_out/bwd_tests.log; tail -n 3 gpurun_out/fwd_tests.log
