#!/bin/bash
# round 2 GPU session: full GPU suite, smoke, bench lines of the main configs
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gpu python -m pytest tests -q -m gpu -x
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
for cfg in ${CFGS:-c2 c2d2}; do
  run bench_$cfg python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-extras
done
cat gpurun_out/summary.txt
tail -n 5 gpurun_out/pytest_gpu.log
tail -n 2 gpurun_out/smoke.log
for cfg in ${CFGS:-c2 c2d2}; do tail -n 1 gpurun_out/bench_$cfg.log | cut -c1-330; done
