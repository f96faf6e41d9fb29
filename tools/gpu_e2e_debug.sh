#!/bin/bash
# r02j: where the e2e step of the 8-GPU weak line spends its time (NM_BENCH_DEBUG), shm return vs NCCL gather
T=r02n; N=${1:-8}; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741"
for g in shm; do
  if [ $g = resident ]; then export NM_BENCH_E2E_RESIDENT=1; G=shm; else G=$g; fi
  NM_BENCH_DEBUG=1 timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline --gather ${G:-$g} > gpurun_out/${T}_n${N}_$g.log 2> gpurun_out/${T}_n${N}_$g.err
  grep "e2e debug" gpurun_out/${T}_n${N}_$g.err | sort | head -8
  python - $N $g <<'PY'
import json, sys
n, g = sys.argv[1], sys.argv[2]
try:
    d = json.loads([x for x in open(f"gpurun_out/r02n_n{n}_{g}.log") if x.startswith("{")][-1])
    print(g, "device", round(d["ms_per_step"], 2), "ms  e2e", round(d["e2e"]["ms_per_step"], 2), "ms", round(d["e2e"]["value"], 1), "Giter/s", d["e2e"]["gather"])
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/r02n_n{n}_{g}.err").read()[-2000:])
PY
done
nproc; free -g | head -2; lscpu | grep -i "numa\|socket\|model name" | head -8
