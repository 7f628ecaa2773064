#!/bin/bash
# round-1 third GPU session: new tests (param grads, graph inversion), bench with side lines, inversion graph timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
D=2 TARGETS=16 timeout 300 python bench_tools/bench_inversion.py > gpurun_out/inv_d2.log 2>&1; echo "inv rc=$?"; cat gpurun_out/inv_d2.log | tail -5
D=2 TARGETS=2 timeout 300 python bench_tools/bench_inversion.py > gpurun_out/inv_d2_t2.log 2>&1; echo "inv rc=$?"; cat gpurun_out/inv_d2_t2.log | tail -5
D=8 TARGETS=16 timeout 300 python bench_tools/bench_inversion.py > gpurun_out/inv_d8.log 2>&1; echo "inv rc=$?"; cat gpurun_out/inv_d8.log | tail -5
