#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exact.py -m gpu -q -x -s > gpurun_out/r02q_pytest_exact.log 2>&1; tail -15 gpurun_out/r02q_pytest_exact.log | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02q_pytest.log 2>&1; tail -4 gpurun_out/r02q_pytest.log
