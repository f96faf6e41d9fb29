#!/bin/bash
# r02e: restructured bench.py on one B200: default run (main line + configs + e2e_view), reference arm
T=r02e; mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python - <<'PY'
import json
try:
    d = json.loads([x for x in open("gpurun_out/r02e_bench.log") if x.startswith("{")][-1])
    r = d["roofline"]
    print("cfg2", round(d["value"], 1), "Giter/s", round(d["ms_per_step"], 3), "ms frac", round(r["frac"], 4), "e2e", round(d["e2e"]["ms_per_step"], 2),
          "ms  e2e_view", d.get("e2e_view"))
    print("parity", d["cpu_baseline"]["parity_on_sample"])
    for k, c in d.get("configs", {}).items():
        print(k, round(c["value"], 1), "Giter/s", round(c["ms_per_step"], 3), "ms frac", c["frac"], "e2e ms", c["e2e"]["ms_per_step"], "host", c["host_precompute_s"],
              "view", (c.get("e2e_view") or {}).get("ms_per_step"), c.get("frames_per_s_device"), c.get("tween_frames_per_step"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r02e_bench.err").read()[-3000:])
PY
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${T}_ref.log 2> gpurun_out/${T}_ref.err; tail -3 gpurun_out/${T}_ref.err; cut -c1-400 gpurun_out/${T}_ref.log
