#!/bin/bash
# round 2: CTA-pair forward kernel: parity on the pair parametrisations, then bench lines
mkdir -p gpurun_out
export C3D_FWD=pair
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "pair or poses" 2>&1 | tail -5
for cfg in ${CFGS:-c2 c2d2}; do
  echo "== pair $cfg"
  timeout 120 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('ms_per_step_min'), d['roofline']['frac'])"
done 2>&1 | tee gpurun_out/pair2_bench.log
