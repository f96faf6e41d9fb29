#!/bin/bash
# r02h: full GPU suite after the guard fix, host-time trace, full-level ncu capture
T=r02h; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
NM_DEBUG_HOST=1 timeout 300 python - > gpurun_out/${T}_host_trace.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import newman_b200
from newman_b200 import workloads
for name in ("cfg2", "cfg3", "cfg4"):
    c = workloads.config(name)
    v = newman_b200.Mandelbrot(c["nr"], c["nc"], N=c["N"], sz=c["sz"], center=c["center"], tol=c["tol"])
    for rep in range(2):
        print("====", name, "call", rep, file=sys.stderr, flush=True)
        t0 = time.perf_counter(); v.render(); dt = time.perf_counter() - t0
        i = v.frame_info()
        print(f"==== {name} call {rep}: {dt:.3f} s  host {i['host_precompute_s']:.3f}  device {i['device_ms']:.1f} ms  probe_exact {i['probe_exact']} consistent {i['probe_consistent']}", file=sys.stderr, flush=True)
PY
grep "====\|nm host" gpurun_out/${T}_host_trace.log | tail -60
timeout 400 bash tools/prof_k3_full.sh ${T} --workload cfg3 --scale 2
