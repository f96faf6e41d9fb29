#!/bin/bash
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_final.sh <tag>'
# Round evidence on one B200 for the library as built: the whole GPU suite, smoke, the default bench line, the reference
# arm, and an ncu --set full capture of one full (quiet) level of k3_fast.
T=${1:-r02z}; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log | cut -c1-120
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python - $T <<'PY'
import json, sys
try:
    d = json.loads([x for x in open(f"gpurun_out/{sys.argv[1]}_bench.log") if x.startswith("{")][-1])
    r = d["roofline"]
    print("cfg2", round(d["value"], 1), "Giter/s", round(d["ms_per_step"], 3), "ms frac", round(r["frac"], 4), "mix", round(r.get("frac_of_mix_ceiling") or 0, 4), "e2e", round(d["e2e"]["ms_per_step"], 2),
          "ms  e2e_view ms", d["e2e_view"]["ms_per_step"], "host", d["e2e_view"]["host_precompute_s_per_step"])
    p = d["cpu_baseline"]["parity_on_sample"]; print("parity", {k: p[k] for k in p if k not in ("truth", "explanation")})
    for k, c in d.get("configs", {}).items():
        print(k, round(c["value"], 1), "Giter/s", round(c["ms_per_step"], 3), "ms frac", c["frac"], "e2e ms", c["e2e"]["ms_per_step"], "host", c["host_precompute_s"])
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/{sys.argv[1]}_bench.err").read()[-3000:])
PY
( time timeout 200 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/${T}_ref.log 2> gpurun_out/${T}_ref.err; cut -c1-200 gpurun_out/${T}_ref.log
timeout 200 bash tools/prof_k3_full.sh ${T}f --workload cfg3 --scale 2
