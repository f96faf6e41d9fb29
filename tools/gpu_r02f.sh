#!/bin/bash
# r02f: bench.py at N = 2 (weak main line + strong cfg3 block + e2e_view through nmm_render)
T=r02f; mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/${T}_bench_n2.log 2> gpurun_out/${T}_bench_n2.err; tail -4 gpurun_out/${T}_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads([x for x in open("gpurun_out/r02f_bench_n2.log") if x.startswith("{")][-1])
    r = d["roofline"]
    print("cfg2 x2", round(d["value"], 1), "Giter/s", round(d["ms_per_step"], 3), "ms frac", round(r["frac"], 4), "e2e", round(d["e2e"]["ms_per_step"], 2), d["e2e"]["value"])
    print("e2e_view", d.get("e2e_view"))
    print("strong", json.dumps(d.get("strong"), indent=1))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r02f_bench_n2.err").read()[-3000:])
PY
