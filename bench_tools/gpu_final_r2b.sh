#!/bin/bash
# Late round-2 GPU session (after the fp16-operand, cta-scope-barrier and fp32-mode changes): full GPU suite, smoke, bench
# lines of every config, launch lists, full-size parity prints.
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" >> gpurun_out/summary.txt; }
run pytest_gpu python -m pytest tests -q -m gpu
run fullsize_prints python -m pytest tests/test_gpu_fullsize.py -q -s -k "full_size or oracle"
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench_c2 python bench.py --steps 10 --warmup 3
run bench_ref python bench.py --impl reference --steps 3 --warmup 1
for cfg in c2d2 c4 c3 c1; do
  run bench_$cfg python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-extras
done
run bench_c5 python bench.py --config c5 --steps 200 --warmup 3
run bench_c3full python bench.py --config c3full --steps 5 --warmup 3
run bench_c2_fp32 python bench.py --config c2 --precision fp32 --steps 3 --warmup 3 --no-extras
D=8 TARGETS=16 STEPS=50 run inv_d8 python bench_tools/bench_inversion.py
run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras
run ncu_list_fp32 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r02b_fp32.csv python bench.py --precision fp32 --steps 1 --warmup 3 --no-cpu-baseline --no-extras
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/pytest_gpu.log
grep -E "^c[0-9]|bf16|fp32" gpurun_out/fullsize_prints.log | cut -c1-250 | head
for f in bench_c2 bench_ref bench_c2d2 bench_c4 bench_c3 bench_c1 bench_c5 bench_c3full bench_c2_fp32; do tail -n 1 gpurun_out/$f.log | cut -c1-200; done
tail -n 4 gpurun_out/inv_d8.log
