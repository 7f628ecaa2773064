#!/bin/bash
# forward A/B of two builds on ONE box: $LIBS (paths of libc3dpp variants; "default" = the in-tree build) x $CFGS, $REPS rounds
mkdir -p gpurun_out
for rep in $(seq 1 ${REPS:-2}); do for lib in ${LIBS:-default bench_tools/_variants/libc3dpp_base.so}; do for cfg in ${CFGS:-c2 c2d2}; do
  if [ "$lib" = default ]; then unset C3D_LIB; else export C3D_LIB=$PWD/$lib; fi
  timeout 600 python bench.py --config $cfg --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/abl.log 2>&1
  tail -n 1 gpurun_out/abl.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib $cfg rep $rep: ms %.3f min %.3f clk %s' % (d['ms_per_step'], d.get('ms_per_step_min', 0), d['clocks']['sm_mhz']))"
done; done; done
