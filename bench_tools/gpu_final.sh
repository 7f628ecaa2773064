#!/bin/bash
# End-of-round GPU session: full GPU suite, smoke, headline bench, launch list + one full ncu capture of the dominant kernel.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gpu python -m pytest tests -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench_c2 python bench.py --steps 10 --warmup 3
run bench_c2d2 python bench.py --config c2d2 --steps 10 --warmup 3 --no-cpu-baseline --no-extras
run bench_c4 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras
run bench_c3 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras
run bench_c1 python bench.py --config c1 --steps 20 --warmup 3 --no-cpu-baseline --no-extras
run bench_c5 python bench.py --config c5 --steps 200 --warmup 3
run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras
run ncu_full ncu --set full --clock-control none --import-source on -k regex:fused_forward_kernel -s 2 -c 1 -f -o gpurun_out/prof_fused python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/pytest_gpu.log
for f in bench_c2 bench_c2d2 bench_c4 bench_c3 bench_c1 bench_c5; do tail -n 1 gpurun_out/$f.log | cut -c1-260; done
