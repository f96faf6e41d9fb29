#!/bin/bash
# One gpurun call that validates and measures a kernel generation on one B200 (≈ 6 GPU-minutes in all):
#   gpurun --timeout 480 -- 'bash tools/gpu_round.sh r02a'            everything below
#   gpurun --timeout 150 -- 'bash tools/gpu_round.sh r02a quick'      parity tests + cfg2/cfg3 bench lines only (≈ 80 s)
# Outputs land in gpurun_out/<tag>_*; afterwards, here:
#   for f in cfg1 cfg2 cfg3 cfg4 cfg5; do grep '^{' gpurun_out/<tag>_bench_$f.log | tail -1 > profiles/<tag>_bench_${f}_n1.json; done
#   python tools/ncu_summary.py launches gpurun_out/<tag>_bench_cfg2_launches.csv > profiles/<tag>_bench_cfg2_launches.txt
#   python tools/ncu_summary.py metrics gpurun_out/<tag>_k3fast.ncu-rep > profiles/<tag>_k3_fast_cfg3_metrics.txt
#   ncu -i gpurun_out/<tag>_k3fast.ncu-rep --page source --csv --print-source sass   (per-instruction executed counts / stall samples)
# A/B of kernel variants in the same call: build with `make -C newman_b200/csrc BUILD=/tmp/b OUT=$PWD/newman_b200/_variants/x.so
# EXTRA=-DK3F_QUIET=0` and run bench.py with NEWMAN_B200_LIB=.../x.so (the variant must export every symbol _lib.py binds).
T=${1:-r02a}; MODE=${2:-full}
mkdir -p gpurun_out
summary() { python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads([x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1])
        r = d.get("roofline", {})
        print(f, round(d["value"], 3), "Giter/s", round(d["ms_per_step"], 2), "ms  e2e", round(d["e2e"]["value"], 2), " frac", r.get("frac"),
              " k3 ms", r.get("k3_ms_per_step"), " glitched", d.get("glitched_per_step"), " clocks", d.get("clocks"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.log").read()[-1200:])
PY
}
timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
if [ "$MODE" = quick ]; then
  timeout 40 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_cfg2.log 2>&1
  timeout 60 python bench.py --no-cpu-baseline --workload cfg3 --steps 2 --warmup 3 > gpurun_out/${T}_bench_cfg3.log 2>&1
  summary ${T}_bench_cfg2 ${T}_bench_cfg3
  exit 0
fi
timeout 120 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log | cut -c1-160
timeout 120 python bench.py > gpurun_out/${T}_bench_cfg2.log 2>&1
timeout 90 python bench.py --impl reference > gpurun_out/${T}_bench_cfg2_ref.log 2>&1
timeout 60 python bench.py --no-cpu-baseline --workload cfg1 > gpurun_out/${T}_bench_cfg1.log 2>&1
timeout 120 python bench.py --no-cpu-baseline --workload cfg3 --steps 2 --warmup 3 > gpurun_out/${T}_bench_cfg3.log 2>&1
timeout 120 python bench.py --no-cpu-baseline --workload cfg4 --steps 2 --warmup 3 > gpurun_out/${T}_bench_cfg4.log 2>&1
timeout 120 python bench.py --no-cpu-baseline --workload cfg5 --steps 2 --warmup 3 > gpurun_out/${T}_bench_cfg5.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_bench_cfg2_launches.csv \
  python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/${T}_launches_run.log 2>&1
timeout 200 bash tools/prof_k3.sh ${T} --workload cfg3 --scale 2
NM_DEBUG_LEVELS=1 timeout 60 python bench.py --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/${T}_levels_cfg2.log 2>&1
NM_DEBUG_LEVELS=1 timeout 90 python bench.py --no-cpu-baseline --workload cfg3 --steps 1 --warmup 3 > gpurun_out/${T}_levels_cfg3.log 2>&1
summary ${T}_bench_cfg2 ${T}_bench_cfg2_ref ${T}_bench_cfg1 ${T}_bench_cfg3 ${T}_bench_cfg4 ${T}_bench_cfg5
