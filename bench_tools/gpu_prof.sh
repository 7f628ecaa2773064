#!/bin/bash
# One gpurun call: full GPU test suite, bench in both cluster modes, ncu launch list + one full capture.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gpu python -m pytest tests -q -m gpu --timeout 300
run bench_cl2 python bench.py --steps 10 --warmup 3
C3D_CLUSTER=1 run bench_cl1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run bench_c4 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline
run bench_c1 python bench.py --config c1 --steps 20 --warmup 3 --no-cpu-baseline
run bench_fp32 python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline
run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline
run ncu_full ncu --set full --clock-control none --import-source on -k regex:fused_forward -s 2 -c 1 -f -o gpurun_out/prof_fused python bench.py --steps 1 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt
for f in bench_cl2 bench_cl1 bench_c4 bench_c1 bench_fp32; do tail -n 1 gpurun_out/$f.log; done
