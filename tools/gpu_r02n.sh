#!/bin/bash
N=${1:-8}
timeout 300 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
bash tools/gpu_e2e_debug.sh $N
