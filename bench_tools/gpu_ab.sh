#!/bin/bash
# A/B of forward-kernel variants: parity tests on the default build, then bench lines per variant.
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/ab_tests.log
for v in default $VARIANTS; do
  for cfg in c2 c2d2; do
    if [ $v = default ]; then L=""; else L=$PWD/bench_tools/_variants/libc3dpp_$v.so; fi
    echo "== $v $cfg" >> gpurun_out/ab.log
    C3D_LIB=$L timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['ms_per_step_min'], d['roofline']['frac'])" >> gpurun_out/ab.log
  done
done
cat gpurun_out/ab_tests.log gpurun_out/ab.log
