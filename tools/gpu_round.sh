#!/bin/bash
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_round.sh <tag>'
# The round's evidence on one B200 — parity suite, smoke, the default bench line (+ configs), reference arm,
# ncu launch list of a bench run, ncu --set full of a full and of an escape level of k3_fast, per-level timings
T=${1:-r02z}; MODE=${2:-full}; mkdir -p gpurun_out   # MODE=noprof skips the ncu captures and the level tables
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
grep "cfg2 vs converged\|(sample id" gpurun_out/${T}_pytest.log | cut -c1-400
timeout 200 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log | cut -c1-200
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python - $T <<'PY'
import json, sys
try:
    d = json.loads([x for x in open(f"gpurun_out/{sys.argv[1]}_bench.log") if x.startswith("{")][-1])
    r = d["roofline"]
    print("cfg2", round(d["value"], 1), "Giter/s", round(d["ms_per_step"], 3), "ms frac", round(r["frac"], 4), "e2e", round(d["e2e"]["ms_per_step"], 2),
          "ms  e2e_view ms", d["e2e_view"]["ms_per_step"], "host", d["e2e_view"]["host_precompute_s_per_step"])
    p = d["cpu_baseline"]["parity_on_sample"]; print("parity", {k: p[k] for k in p if k not in ("truth", "explanation")})
    for k, c in d.get("configs", {}).items():
        print(k, round(c["value"], 1), "Giter/s", round(c["ms_per_step"], 3), "ms frac", c["frac"], "e2e ms", c["e2e"]["ms_per_step"], "host", c["host_precompute_s"],
              "view", (c.get("e2e_view") or {}).get("ms_per_step"), "refs", c.get("secondary_references"), "glitched", c.get("glitched_per_step"),
              c.get("frames_per_s_device"), c.get("tween_frames_per_step"))
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/{sys.argv[1]}_bench.err").read()[-3000:])
PY
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/${T}_ref.log 2> gpurun_out/${T}_ref.err; tail -3 gpurun_out/${T}_ref.err; cut -c1-300 gpurun_out/${T}_ref.log
if [ "$MODE" = full ]; then
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_bench_cfg2_launches.csv \
  python bench.py --no-cpu-baseline --no-extras --steps 2 --warmup 1 > gpurun_out/${T}_launches_run.log 2>&1
timeout 300 bash tools/prof_k3.sh ${T} --workload cfg3 --scale 2 --no-extras
timeout 300 bash tools/prof_k3_full.sh ${T}f --workload cfg3 --scale 2
NM_DEBUG_LEVELS=1 timeout 90 python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 3 > gpurun_out/${T}_levels_cfg2.log 2>&1
NM_DEBUG_LEVELS=1 timeout 120 python bench.py --no-cpu-baseline --no-extras --workload cfg3 --steps 1 --warmup 3 > gpurun_out/${T}_levels_cfg3.log 2>&1
grep "nm level" gpurun_out/${T}_levels_cfg2.log | tail -18
fi
# the zoom video with all of its 600 key frames on this one GPU
( time timeout 1500 python bench.py --workload cfg5 --frames 600 --steps 1 --warmup 3 --no-extras ) > gpurun_out/${T}_cfg5_600.log 2> gpurun_out/${T}_cfg5_600.err; tail -3 gpurun_out/${T}_cfg5_600.err
python - $T <<'PY'
import json, sys
try:
    d = json.loads([x for x in open(f"gpurun_out/{sys.argv[1]}_cfg5_600.log") if x.startswith("{")][-1])
    print("cfg5 x600", round(d["value"], 1), "Giter/s", round(d["ms_per_step"], 1), "ms per pass; key frames/s", d["frames_per_s_device"], "e2e", d["frames_per_s_e2e"], "host", d["host_precompute_s"])
except Exception as e:
    print("cfg5 FAILED", e)
PY
