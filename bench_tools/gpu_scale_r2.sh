#!/bin/bash
# round-2 multi-GPU session on N GPUs (N = $1): the headline bench with the map gather inside the timed region (p2p, fused, NCCL
# all timed), the shared-latent inversion (gradient all-reduce in the step's CUDA graph), the 2-GPU tests
N=${1:-8}
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err
echo "c2 n=$N rc=$?"; tail -n 1 gpurun_out/r2_scale_$N.json | cut -c1-1500
tail -n 3 gpurun_out/r2_scale_$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config c5 --shared-latent --steps 20 --warmup 3 > gpurun_out/r2_c5_shared_$N.json 2> gpurun_out/r2_c5_shared_$N.err
echo "c5 shared n=$N rc=$?"; tail -n 1 gpurun_out/r2_c5_shared_$N.json | cut -c1-700
timeout 300 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -2
