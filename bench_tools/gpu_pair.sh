#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "bf16 or render or smoke" 2>&1 | tail -15
for cfg in c2 c2d2 c4 c1 c3; do
echo "== $cfg"; timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ms', round(d['ms_per_step'], 3), 'min', round(d['ms_per_step_min'], 3), 'TF', round(d['roofline']['achieved'], 1), 'clk', d['clocks'])
    elif 'rror' in l or 'c3d' in l: print(l.strip())
"; done
#!/bin/bash
for cfg in c2 c2d2; do
echo "== $cfg"
C3D_LIB=$PWD/bench_tools/_variants/libc3dpp_prof.so C3D_DEBUG=2 timeout 300 python bench.py --config $cfg --steps 1 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | grep "c3d prof" | sort | tail -6
done
