#!/bin/bash
# r02k: kernel-written read-backs — parity suite + cfg2 line
T=r02k; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
NM_BENCH_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err
grep "e2e debug" gpurun_out/${T}_bench.err
python - <<'PY'
import json
d = json.loads([x for x in open("gpurun_out/r02k_bench.log") if x.startswith("{")][-1])
print("device", round(d["ms_per_step"], 3), "ms", round(d["value"], 1), " e2e", round(d["e2e"]["ms_per_step"], 3), "ms", round(d["e2e"]["value"], 1))
PY
