#!/bin/bash
# r02c: early export + k3_finish on the event queue + rounded orbit table — parity suite + A/B (one B200)
T=r02c; mkdir -p gpurun_out
V=$PWD/newman_b200/_variants
timeout 400 python -m pytest tests -m gpu -x -q -s > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
grep "k3 ms with\|checked steps" gpurun_out/${T}_pytest.log
b() { tag=$1; shift
  env "$@" timeout 150 python bench.py --no-cpu-baseline ${BARGS} > gpurun_out/${T}_${tag}.log 2>&1
  python - ${T}_$tag <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1])
    r = d["roofline"]
    print(f, round(d["value"], 1), "Giter/s", round(d["ms_per_step"], 3), "ms  frac", round(r["frac"], 4), " k3 ms", round(r["k3_ms_per_step"], 3),
          "launches", d["gpu_launches"], "sm", d["clocks"]["sm_mhz"], "e2e", round(d["e2e"]["ms_per_step"], 2))
except Exception as e:
    print(f, "FAILED", e); print(open(f"gpurun_out/{f}.log").read()[-800:])
PY
}
BARGS="--steps 10 --warmup 3"
b cfg2_loudq1 NM_K3_LOUDQ=1
b cfg2_loudq0 NM_K3_LOUDQ=0
b cfg2_lq16 NEWMAN_B200_LIB=$V/lq16.so
b cfg2_lq4 NEWMAN_B200_LIB=$V/lq4.so
b cfg2_lq32 NEWMAN_B200_LIB=$V/lq32.so
BARGS="--workload cfg3 --steps 2 --warmup 3"
b cfg3_loudq1 NM_K3_LOUDQ=1
b cfg3_loudq0 NM_K3_LOUDQ=0
BARGS="--workload cfg4 --steps 2 --warmup 3"
b cfg4_loudq1 NM_K3_LOUDQ=1
b cfg4_loudq0 NM_K3_LOUDQ=0
BARGS="--workload cfg5 --steps 2 --warmup 3"
b cfg5_loudq1 NM_K3_LOUDQ=1
NM_DEBUG_LEVELS=1 timeout 60 python bench.py --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/${T}_levels_cfg2.log 2>&1
grep "nm level" gpurun_out/${T}_levels_cfg2.log | tail -18
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_bench_cfg2_launches.csv \
  python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/${T}_launches_run.log 2>&1
