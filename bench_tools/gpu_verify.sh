#!/bin/bash
# final verification of a build: full GPU suite, smoke, the default bench line, the fp32-mode line
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gpu python -m pytest tests -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench_default python bench.py
run bench_ref python bench.py --impl reference --steps 3 --warmup 1
run bench_c2_fp32 python bench.py --config c2 --precision fp32 --steps 3 --warmup 3 --no-extras
run bench_c5 python bench.py --config c5 --steps 200 --warmup 3
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/pytest_gpu.log
tail -n 2 gpurun_out/smoke.log
for f in bench_default bench_ref bench_c2_fp32 bench_c5; do tail -n 1 gpurun_out/$f.log | cut -c1-260; done
