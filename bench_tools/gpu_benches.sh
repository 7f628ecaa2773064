#!/bin/bash
# Bench lines of every BASELINE config on the current build (no profiler).
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -n 1 gpurun_out/$name.log | cut -c1-200; }
run fin_c2 python bench.py --steps 10 --warmup 3
run fin_c2d2 python bench.py --config c2d2 --steps 10 --warmup 3 --no-cpu-baseline --no-extras
run fin_c4 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras
run fin_c3 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras
run fin_c1 python bench.py --config c1 --steps 20 --warmup 3 --no-cpu-baseline --no-extras
run fin_c5 python bench.py --config c5 --steps 200 --warmup 3
run fin_ref python bench.py --impl reference --steps 3 --warmup 1
