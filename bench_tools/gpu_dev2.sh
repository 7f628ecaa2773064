#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/dbg_*.log
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run k16 python bench_tools/k16_probe.py
for cl in 1 2; do
  for cs in ffhq_d2_n24 ffhq_d8_n24 ffhq_d2_n128_static cars_d6_n36_b2_beta ffhq_d8_n24_b2_wplus_perturb; do
    C3D_CLUSTER=$cl run dbg_${cs}_c$cl python bench_tools/debug_fused.py $cs bf16 points
  done
  C3D_CLUSTER=$cl run dbg_poses_c$cl python bench_tools/debug_fused.py ffhq_d8_n24 bf16 poses
done
run bench_v2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
C3D_CLUSTER=1 run bench_v2_cl1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
C3D_FUSED=1 run bench_v1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run pytest_gpu python -m pytest tests -q -m gpu --timeout 300
run ncu_full ncu --set full --clock-control none --import-source on -k regex:fused_forward -s 2 -c 1 -f -o gpurun_out/prof_fused_v2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt; cat gpurun_out/k16.log | tail -n 3
tail -q -n 1 gpurun_out/dbg_*.log
for f in bench_v2 bench_v2_cl1 bench_v1; do tail -n 1 gpurun_out/$f.log | cut -c 1-330; done
tail -n 3 gpurun_out/pytest_gpu.log
