"""Host time of Mandelbrot::renderFrame per stage (NM_DEBUG_HOST=1 prints the laps) for the bench views, two calls each,
with and without the speculative primary reference build.  usage (GPU box): NM_DEBUG_HOST=1 python tools/host_trace.py cfg2 cfg3"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import newman_b200  # noqa: E402
from newman_b200 import workloads  # noqa: E402

for w in sys.argv[1:] or ["cfg2"]:
    cfg = workloads.config(w)
    for spec in (1, 0):
        if spec:
            os.environ.pop("NM_NO_SPECULATION", None)
        else:
            os.environ["NM_NO_SPECULATION"] = "1"
        view = newman_b200.Mandelbrot(cfg["nr"], cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
        out = torch.empty((cfg["nr"], cfg["nc"], 2), dtype=torch.int32).pin_memory()
        for call in range(2):
            print("==== %s speculation=%d call %d" % (w, spec, call), flush=True)
            t0 = time.perf_counter()
            view.render(out)
            i = view.frame_info()
            print("==== %s speculation=%d call %d: %.3f s  host %.3f  device %.1f ms" % (w, spec, call, time.perf_counter() - t0, i["host_precompute_s"], i["device_ms"]), flush=True)
