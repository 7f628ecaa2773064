#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gpu python -m pytest tests -q -m gpu --timeout 300
run bench_c2 python bench.py --steps 10 --warmup 3
run bench_c4 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline
run bench_c1 python bench.py --config c1 --steps 20 --warmup 3 --no-cpu-baseline
D=2 TARGETS=16 run inv_d2 python bench_tools/bench_inversion.py
D=8 TARGETS=16 run inv_d8 python bench_tools/bench_inversion.py
D=2 TARGETS=2 run inv_d2_t2 python bench_tools/bench_inversion.py
D=2 TARGETS=16 C3D_BWD=simt run inv_d2_simt python bench_tools/bench_inversion.py
D=2 TARGETS=16 run ncu_bwd_list ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/launches_inv.csv python bench_tools/bench_inversion.py
D=8 TARGETS=16 run ncu_bwd ncu --set full --clock-control none --import-source on -k regex:fused_backward -s 2 -c 1 -f -o gpurun_out/prof_bwd python bench_tools/bench_inversion.py
cat gpurun_out/summary.txt
for f in bench_c2 bench_c4 bench_c1 inv_d2 inv_d8 inv_d2_t2 inv_d2_simt; do tail -n 1 gpurun_out/$f.log | cut -c 1-260; done
