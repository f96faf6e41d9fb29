"""Host time of one reference build (orbit + A/B/C series, hp_host.cpp: build_tables) per bench view: serial form,
4-stage and 6-stage pipeline. No GPU involved.   python tools/host_tables_bench.py > profiles/<tag>_host_tables.txt"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import newman_b200
from newman_b200 import workloads


def main():
    print("# one reference build per view (the frame's centre sample as reference point), best of 5, seconds; %d host cores" % os.cpu_count())
    print("# view   M        bits   serial   4 stages   6 stages   identical")
    for name, scale in (("cfg2", 40), ("cfg3", 80), ("cfg4", 80)):
        cfg = workloads.config(name, scale=scale)
        best, tabs = {}, {}
        for threads in (1, 4, 6):
            v = newman_b200.Mandelbrot(cfg["nr"], cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"],
                                       host_threads=threads)
            ts = []
            for _ in range(5):
                t0 = time.perf_counter()
                h = v.host_tables(cfg["nr"] // 2, cfg["nc"] // 2)
                ts.append(time.perf_counter() - t0)
            best[threads], tabs[threads] = min(ts), h
            bits = v.precision_bits()
        same = all(np.ascontiguousarray(tabs[1][k]).tobytes() == np.ascontiguousarray(tabs[t][k]).tobytes()
                   for t in (4, 6) for k in ("x_hi", "x_lo", "a", "b", "c", "a_m", "a_e", "b_m", "b_e", "c_m", "c_e"))
        print("%-6s %-8d %-6d %-8.3f %-10.3f %-10.3f %s" % (name, tabs[1]["M"], bits, best[1], best[4], best[6], same))


if __name__ == "__main__":
    main()
