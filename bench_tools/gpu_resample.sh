#!/bin/bash
# Importance-resampling extension on the GPU box: parity tests, device timing against the HBM roofline, ncu capture.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_resample.py -q -m gpu -x 2>&1 | tail -30 > gpurun_out/resample_tests.log
timeout 200 python bench_tools/resample_probe.py ${PROBE_ARGS} > gpurun_out/resample_probe.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sample_pdf -s 6 -c 1 -f -o gpurun_out/prof_resample \
  python bench_tools/resample_probe.py > gpurun_out/ncu_resample.log 2>&1
tail -3 gpurun_out/resample_tests.log
grep -v Traceback gpurun_out/resample_probe.log | cut -c1-250
tail -2 gpurun_out/ncu_resample.log
