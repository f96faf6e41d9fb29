#!/bin/bash
# usage: gpurun -- 'bash tools/gpu_resched_bisect.sh v1 v2 ...' : the k3_fast parity tests against each library variant
for V in "$@"; do
  L=newman_b200/libnewman_b200_$V.so; [ "$V" = "-" ] && L=newman_b200/libnewman_b200.so
  NEWMAN_B200_LIB=$PWD/$L timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants_agree or bit_exact_vs_oraclep" 2>&1 | grep -E "^(FAILED|ERROR)|passed|failed" | cut -c1-200 | sed "s/^/$V: /"
done
