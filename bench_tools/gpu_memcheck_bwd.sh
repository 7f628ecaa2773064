#!/bin/bash
# compute-sanitizer memcheck of the save-mode forward + tensor-core backward (fp16 saved tiles, gdot recompute) on the final build.
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_backward.py -q -m gpu -x -k "saved_forward or ffhq_d2_n24" > gpurun_out/memcheck_bwd.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_bwd.log | tail -n 4
