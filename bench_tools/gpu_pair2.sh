#!/bin/bash
# round 2: CTA-pair forward kernel (half-split jobs): parity on the pair parametrisations, then bench lines pair vs single-CTA
mkdir -p gpurun_out
export C3D_FWD=pair
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "pair or poses" 2>&1 | tail -8 > gpurun_out/pair2_tests.log
cat gpurun_out/pair2_tests.log
for fwd in pair v3; do
  for cfg in c2 c2d2; do
    echo "== $fwd $cfg"
    C3D_FWD=$fwd timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('ms_per_step_min'), d['roofline']['frac'])"
  done
done 2>&1 | tee gpurun_out/pair2_bench.log
