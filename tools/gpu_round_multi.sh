#!/bin/bash
# usage: gpurun --gpus 8 --timeout 2400 -- 'bash tools/gpu_round_multi.sh 8 <tag>'
# The multi-GPU evidence: group tests, bench.py --gpus 8 (weak cfg2 + strong cfg3 + e2e_view), cfg5 with all 600 key
# frames, and the C++ caller's beauty render of the cfg3 view on all GPUs
N=${1:-8}; T=${2:-r02z}; mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_dropin_cpp.py tests/test_render_cli.py -m gpu -q -s > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
grep -i "MULTI-RANK\|DIFFERS\|beauty pass\|beauty identical" gpurun_out/${T}_pytest.log | head
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733"
( time timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/${T}_bench_n${N}.log 2> gpurun_out/${T}_bench_n${N}.err; tail -4 gpurun_out/${T}_bench_n${N}.err
( time timeout 1500 $TR bench.py --gpus $N --workload cfg5 --frames 600 --steps 1 --warmup 3 ) > gpurun_out/${T}_cfg5_n${N}.log 2> gpurun_out/${T}_cfg5_n${N}.err; tail -4 gpurun_out/${T}_cfg5_n${N}.err
python - $N $T <<'PY'
import json, sys
n, t = sys.argv[1], sys.argv[2]
for f in (f"{t}_bench_n{n}", f"{t}_cfg5_n{n}"):
    try:
        d = json.loads([x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1])
        r = d["roofline"]
        print(f, round(d["value"], 1), "Giter/s", round(d["ms_per_step"], 3), "ms frac", round(r["frac"], 4), "e2e", round(d["e2e"]["ms_per_step"], 2), round(d["e2e"]["value"], 1),
              "fps", d.get("frames_per_s_device"), d.get("frames_per_s_e2e"), "host", d.get("host_precompute_s"))
        print("  e2e_view", d.get("e2e_view"))
        print("  strong", json.dumps(d.get("strong")))
    except Exception as e:
        print(f, "FAILED", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
# the C++ caller: FractalViewer::beautyRender of the cfg3 view (3840x2160, 4x) on one GPU and on all of them
python - <<'PY'
import sys
sys.path.insert(0, ".")
import newman_b200
from newman_b200 import workloads
c = workloads.config("cfg3")
from fractions import Fraction
from newman_b200.workloads import _dec
d = Fraction(1, 10 ** 100)
src = newman_b200.Mandelbrot(600, 800, N=c["N"], sz=(_dec(4 * d / 800, 60), _dec(3 * d / 600, 60)), center=c["center"], tol=c["tol"])
src.save("gpurun_out/cfg3_view.txt")
PY
g++ -std=c++11 -O2 -Iinclude/newman_b200 -Inewman_b200/csrc/compat tests/dropin/headless_viewer.cpp -o /tmp/headless_viewer -Lnewman_b200 -l:libnewman_b200.so -Wl,-rpath,$PWD/newman_b200 -l:libgmp.so.10
( cd newman_b200 && timeout 600 /tmp/headless_viewer beauty 2160 3840 4 1048576 $N ../gpurun_out/cfg3_view.txt 120 ) > gpurun_out/${T}_beauty_cpp.log 2>&1; cat gpurun_out/${T}_beauty_cpp.log
