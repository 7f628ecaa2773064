#!/bin/bash
# multi-GPU session: weak-scaling bench (N = 8, 4, 2, 1) and sharded inversion
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  echo "n=$n rc=$?"; tail -n 1 gpurun_out/scale_$n.json | cut -c1-400
done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err; tail -n 1 gpurun_out/scale_1.json | cut -c1-300
for n in 8 1; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench_tools/bench_inversion_dist.py > gpurun_out/inv_dist_$n.json 2> gpurun_out/inv_dist_$n.err
  echo "inv n=$n rc=$?"; tail -n 1 gpurun_out/inv_dist_$n.json
done
