#!/bin/bash
# compute-sanitizer memcheck of the whole GPU test suite on the final build.
mkdir -p gpurun_out
timeout 240 compute-sanitizer --tool memcheck python -m pytest tests -q -m gpu > gpurun_out/memcheck_all.log 2>&1
echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_all.log | tail -n 4
