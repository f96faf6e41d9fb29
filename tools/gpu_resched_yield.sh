#!/bin/bash
# usage: gpurun -- 'bash tools/gpu_resched_yield.sh [variant...]' : control-field experiments on the re-registered quiet block
run() { # label, lib, env...
  L=$2; lab=$1; shift; shift
  env "$@" NEWMAN_B200_LIB=$PWD/$L timeout 120 python bench.py --no-cpu-baseline --no-extras --workload ${W:-cfg2} --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lab', round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'Giter/s frac', round(d['roofline']['frac'],4))
"
}
run default newman_b200/libnewman_b200.so X=1
for V in "$@"; do
  run $V newman_b200/libnewman_b200_$V.so X=1
  NEWMAN_B200_LIB=$PWD/newman_b200/libnewman_b200_$V.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants_agree or bit_exact_vs_oraclep" 2>&1 | tail -1
done
run default newman_b200/libnewman_b200.so X=1
