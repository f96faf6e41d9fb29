#!/bin/bash
# usage: gpurun --timeout 900 -- 'bash tools/gpu_resched_ab.sh <tag>'
# A/B of the instruction orders tools/sass_resched.py produces for k3_fast's quiet block: the libraries
# newman_b200/libnewman_b200_<variant>.so (built here beforehand) are benchmarked back to back on the same box.
T=${1:-ab}; mkdir -p gpurun_out
for V in plain "" noyield chain plain ""; do
  L=newman_b200/libnewman_b200${V:+_$V}.so
  [ -f $L ] || continue
  for W in cfg2 cfg3; do
    NEWMAN_B200_LIB=$PWD/$L timeout 120 python bench.py --no-cpu-baseline --no-extras --workload $W --steps 10 --warmup 3 2> gpurun_out/${T}_err.log | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('${V:-resched}', '$W', round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'Giter/s frac', round(d['roofline']['frac'],4), 'clk', d['clocks'].get('sm_mhz'))
"
  done
done
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_finish.py -m gpu -x -q 2>&1 | tail -2
NEWMAN_B200_LIB=$PWD/newman_b200/libnewman_b200_chain.so timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
NM_DEBUG_LEVELS=1 timeout 90 python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 3 2>&1 | grep "nm level" | tail -18 > gpurun_out/${T}_levels_resched.txt
NEWMAN_B200_LIB=$PWD/newman_b200/libnewman_b200_plain.so NM_DEBUG_LEVELS=1 timeout 90 python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 3 2>&1 | grep "nm level" | tail -18 > gpurun_out/${T}_levels_plain.txt
paste gpurun_out/${T}_levels_plain.txt gpurun_out/${T}_levels_resched.txt | cut -c1-200 | awk '{print $6, $7, $11, $12, "|", $21, $22}'
