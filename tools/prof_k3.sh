#!/bin/bash
# ncu --set full captures of the dominant kernel on full 1024-iteration levels of cfg3 at half size
# (33.2 M samples): the plain FP64 kernel and the scaled (floatexp delta) kernel. Run under gpurun;
# summaries: python tools/ncu_summary.py metrics gpurun_out/<name>.ncu-rep > profiles/<name>_metrics.txt
if [ "$1" != "scaled" ]; then
ncu --set full --clock-control none --import-source on -k regex:k3_fast -s 396 -c 2 -o gpurun_out/r01f_k3fast_cfg3 -f python bench.py --workload cfg3 --scale 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b_prof.log 2>&1
fi
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k3_fast.*bool.1' -s 127 -c 2 -o gpurun_out/r01f_k3fast_scaled -f python bench.py --workload cfg3 --scale 2 --steps 1 --warmup 3 --no-cpu-baseline --floatexp 2 > gpurun_out/b_prof2.log 2>&1
tail -n 3 gpurun_out/b_prof2.log
