#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_parity.py -x -q -k "eikonal or mlp_init or error" 2>&1 | tail -12
