"""Device-resident timing of the cfg2 primary frame under kernel options (tuning aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import newman_b200
from newman_b200 import workloads, pipeline
from newman_b200 import _lib as L

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = workloads.config("cfg2", scale=scale)
m = newman_b200.Mandelbrot(cfg["nr"], cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
t0 = time.time(); h = m.host_tables(); print("host tables %.2fs M=%d" % (time.time() - t0, h["M"]))
ts = pipeline.TableSet(h, cfg["N"], cfg["tol"], 1e-6)
dev = newman_b200.Device(0)
for group in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "2", "4"])]:
    dev.set_option(L.OPT_K3_GROUP, group)
    best = None
    for rep in range(4):
        dev.frame_deep(ts.tables(), ts.arr["eps_re"], ts.arr["eps_im"])
        dev.launch()
        st = dev.stats()
        if best is None or st["ms_k3"] < best["ms_k3"]:
            best = st
    it = best["executed_iters"]
    print("group %d: k2 %.2f ms  k3 %.2f ms  executed %.3e  -> %.1f Giter/s, %.1f%% of 1.861e13 inst/s ; glitched %d launches %d"
          % (group, best["ms_k2"], best["ms_k3"], it, it / best["ms_k3"] / 1e6, it * 10 / (best["ms_k3"] * 1e-3) / 1.861e13 * 100,
             best["glitched"], best["kernel_launches"]), "checked lane-steps %.3e" % best["checked_steps"])
