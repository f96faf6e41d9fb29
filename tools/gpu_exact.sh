#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exact.py -m gpu -q -x -s -k cfg2 > gpurun_out/r02q_pytest_exact.log 2>&1; grep "cfg2: plain\|still different" gpurun_out/r02q_pytest_exact.log | cut -c1-600
