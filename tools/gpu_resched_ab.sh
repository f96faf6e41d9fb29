#!/bin/bash
# usage: gpurun --timeout 900 -- 'bash tools/gpu_resched_ab.sh <tag> [variants...]'
# A/B of the instruction orders tools/sass_resched.py produces for k3_fast's quiet block: the libraries
# newman_b200/libnewman_b200_<variant>.so (built here beforehand: `make -C newman_b200/csrc RESCHED=0 OUT=$PWD/newman_b200/libnewman_b200_plain.so`,
# then `python tools/sass_resched.py ..._plain.so ..._<variant>.so [flags]`; "-" = the in-tree library) are benchmarked back to
# back on the same box, and the parity tests that compare k3_fast with the oracle bit for bit run on each.
T=${1:-ab}; shift; mkdir -p gpurun_out
VARS=${@:-plain - plain -}
for V in $VARS; do
  if [ "$V" = "-" ]; then L=newman_b200/libnewman_b200.so; else L=newman_b200/libnewman_b200_$V.so; fi
  [ -f $L ] || continue
  for W in ${WORKLOADS:-cfg2 cfg3 cfg4}; do
    NEWMAN_B200_LIB=$PWD/$L timeout 160 python bench.py --no-cpu-baseline --no-extras --workload $W --steps 10 --warmup 3 2> gpurun_out/${T}_err.log | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$V', '$W', round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'Giter/s frac', round(d['roofline']['frac'],4), 'clk', d['clocks'].get('sm_mhz'))
"
  done
  NEWMAN_B200_LIB=$PWD/$L timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_finish.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/${T}_pytest_$V.log 2>&1
  grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${T}_pytest_$V.log | cut -c1-300
done
NM_DEBUG_LEVELS=1 timeout 90 python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 3 2>&1 | grep "nm level" | tail -18 | awk '{print $6, $7, $11, $12}' > gpurun_out/${T}_levels.txt
sed -n 3,9p gpurun_out/${T}_levels.txt
