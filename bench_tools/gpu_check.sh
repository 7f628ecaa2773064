#!/bin/bash
# Mid-round health check on the GPU box: full GPU suite, smoke, headline bench (with side lines), resample ncu capture.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gpu python -m pytest tests -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench_c2 python bench.py --steps 10 --warmup 3
run ncu_resample ncu --set full --clock-control none --import-source on -k regex:sample_pdf -s 30 -c 1 -f -o gpurun_out/prof_resample python bench_tools/resample_probe.py
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/pytest_gpu.log
tail -n 1 gpurun_out/bench_c2.log | cut -c1-3000
