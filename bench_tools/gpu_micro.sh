#!/bin/bash
# per-SM micro-benchmarks (bench_tools/micro/*.cu); the binaries are built on the CPU box and travel with the snapshot
mkdir -p gpurun_out
for b in ${PROBES:-sm_probe tma_probe}; do
  timeout 300 bench_tools/micro/$b > gpurun_out/micro_$b.log 2>&1
  echo "rc=$?" >> gpurun_out/micro_$b.log
done
tail -n 3 gpurun_out/micro_*.log
