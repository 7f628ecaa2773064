#!/bin/bash
# Dev battery for one gpurun call: each group in its own process so a trap in one kernel cannot hide the rest.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run t_aux python -m pytest tests/test_gpu_parity.py -q --timeout 120 -k "library or style_prep or raygen or camera or composite"
run t_umma python -m pytest tests/test_gpu_parity.py -q --timeout 120 -k umma
run t_fp32 python -m pytest tests/test_gpu_parity.py -q --timeout 200 -k "forward_fp32"
for cl in 1 2; do
  for cs in ffhq_d2_n24 ffhq_d8_n24 ffhq_d2_n128_static cars_d6_n36_b2_beta ffhq_d8_n24_b2_wplus_perturb; do
    C3D_CLUSTER=$cl run dbg_${cs}_c$cl python bench_tools/debug_fused.py $cs bf16 points
  done
  C3D_CLUSTER=$cl run dbg_poses_c$cl python bench_tools/debug_fused.py ffhq_d8_n24 bf16 poses
done
run t_rest python -m pytest tests/test_gpu_parity.py -q --timeout 200 -k "render or argument"
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/dbg_*.log
