#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/dbg_*.log
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
for egw in 4 8; do
for cs in ffhq_d2_n24 ffhq_d8_n24 ffhq_d2_n128_static cars_d6_n36_b2_beta; do
  C3D_EGW=$egw run dbg_${cs}_e$egw python bench_tools/debug_fused.py $cs bf16 points
done
C3D_EGW=$egw run bench_e$egw python bench.py --steps 10 --warmup 3 --no-cpu-baseline
C3D_EGW=$egw C3D_DEBUG=2 run prof_e$egw python bench.py --steps 1 --warmup 3 --no-cpu-baseline
done
C3D_CLUSTER=1 run bench_cl1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run pytest_gpu python -m pytest tests -q -m gpu --timeout 300
run ncu_full ncu --set full --clock-control none --import-source on -k regex:fused_forward -s 2 -c 1 -f -o gpurun_out/prof_fused_v3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt
tail -q -n 1 gpurun_out/dbg_*.log | cut -c 1-200
for f in bench_e4 bench_e8 bench_cl1; do tail -n 1 gpurun_out/$f.log | cut -c 1-330; done
grep -h "c3d prof" gpurun_out/prof_e4.log | tail -n 3; grep -h "c3d prof" gpurun_out/prof_e8.log | tail -n 3
tail -n 3 gpurun_out/pytest_gpu.log
