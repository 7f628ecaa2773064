#!/bin/bash
# A/B of a forward-kernel change on one box: parity prints of the forward tests + bench lines of the three forward configs
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; }
run parity_prints python -m pytest tests/test_gpu_parity.py -q -s -k "edge_shapes or bf16_matches_reference_golden or full_size"
for cfg in ${CFGS:-c2 c2d2 c3}; do
  run ab_$cfg python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-extras
  tail -n 1 gpurun_out/ab_$cfg.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], d['ms_per_step'], d['ms_per_step_min'], d['clocks'])"
done
tail -n 3 gpurun_out/parity_prints.log
