#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_backward.py -x -q -k "eikonal or mlp_init" 2>&1 | tail -25
