#!/bin/bash
# GPU session after the operand-format change: full GPU suite, smoke, gradient errors printed, bench lines with in-run parity
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gpu python -m pytest tests -q -m gpu
run grads python -m pytest tests/test_gpu_backward.py -q -s -k "gradients_match_reference_autograd or eikonal"
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench_c2 python bench.py --config c2 --steps 10 --warmup 3
for cfg in ${CFGS:-c2d2 c5}; do
  run bench_$cfg python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-extras
done
cat gpurun_out/summary.txt
tail -n 15 gpurun_out/pytest_gpu.log
grep -E "^(ffhq|cars)|errs|passed|failed" gpurun_out/grads.log | cut -c1-400
tail -n 2 gpurun_out/smoke.log
for cfg in c2 ${CFGS:-c2d2 c5}; do tail -n 1 gpurun_out/bench_$cfg.log | cut -c1-2500; done
