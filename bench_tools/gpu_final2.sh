#!/bin/bash
# Final 2-GPU check of the end-of-round build: NCCL tests, torchrun bench (both arms), sharded inversion.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -n 2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/final_scale_2.json 2> gpurun_out/final_scale_2.err
echo "bench n=2 rc=$?"; tail -n 1 gpurun_out/final_scale_2.json | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/final_ref_2.json 2> gpurun_out/final_ref_2.err
echo "ref n=2 rc=$?"; tail -n 1 gpurun_out/final_ref_2.json | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench_tools/bench_inversion_dist.py > gpurun_out/final_inv_dist_2.json 2> gpurun_out/final_inv_dist_2.err
echo "inv n=2 rc=$?"; tail -n 1 gpurun_out/final_inv_dist_2.json
