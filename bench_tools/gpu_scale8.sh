#!/bin/bash
# 8-GPU session (trimmed): weak-scaling headline bench at N=8 and the sharded inversion (strong scaling of 16 targets).
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/scale_8.json 2> gpurun_out/scale_8.err
echo "n=8 rc=$?"; tail -n 1 gpurun_out/scale_8.json | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench_tools/bench_inversion_dist.py > gpurun_out/inv_dist_8.json 2> gpurun_out/inv_dist_8.err
echo "inv n=8 rc=$?"; tail -n 1 gpurun_out/inv_dist_8.json
timeout 300 python bench_tools/bench_inversion_dist.py > gpurun_out/inv_dist_1.json 2> gpurun_out/inv_dist_1.err
echo "inv n=1 rc=$?"; tail -n 1 gpurun_out/inv_dist_1.json
timeout 300 python -m pytest tests/test_gpu_dist.py tests/test_gpu_resample.py -q -m gpu -k "dist or perturbed" 2>&1 | tail -2
