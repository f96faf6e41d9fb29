#!/bin/bash
# r02d: multi-GPU behind the C-ABI (2 GPUs) — parity suite incl. the group tests, torchrun worker
T=r02d; mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
grep -i "MULTI-RANK\|identical\|DIFFERS\|beauty pass" gpurun_out/${T}_pytest.log | head -20
