"""Greedy auto-zoom that produced the deep bench/test views in newman_b200/workloads.py.

At every depth D = 10^-k a small raster (24 x 32 samples spanning the 4D x 3D extent of the view,
SURVEY.md 8d) is rendered on the CPU — host tables from the product's own C++/GMP code, pixels by
Oracle-P (test infrastructure; this is a fixture generator, not a product path) — and the view is
re-centred on a high-count, non-interior sample before zooming in by `--step`. Prints one line per
level and the final centre with enough digits for the target depth.

  python tools/zoom_view.py --start-re -0.75 --start-im 1e-5 --depth 100 --N 1048576
"""
import argparse
import os
import sys
import time
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import newman_b200  # noqa: E402
import oracles  # noqa: E402
from newman_b200 import pipeline  # noqa: E402


def dec(fr, digits):
    fr = Fraction(fr)
    neg = fr < 0
    fr = abs(fr)
    ip = fr.numerator // fr.denominator
    frac = fr - ip
    s = str((frac.numerator * 10 ** digits) // frac.denominator).rjust(digits, "0")
    return ("-" if neg else "") + f"{ip}.{s}"


def sci(fr, digits=30):
    fr = Fraction(fr)
    e = 0
    while fr >= 10:
        fr /= 10; e += 1
    while fr < 1:
        fr *= 10; e -= 1
    m = (fr.numerator * 10 ** digits) // fr.denominator
    s = str(m)
    return f"{s[0]}.{s[1:]}e{e}"


def render(nr, nc, N, pitch_re, pitch_im, cre, cim, digits):
    m = newman_b200.Mandelbrot(nr, nc, N=N, sz=(sci(pitch_re), sci(pitch_im)), center=(dec(cre, digits), dec(cim, digits)))
    if m.useHardware():
        c_re, c_im = m.host_coords()
        out, _ = oracles.p_render_hw(c_re, c_im, N)
        return out["iterations"], 0, 0
    mk = lambda d: pipeline.TableSet(d, N, 1e-10, 1e-6, pipeline.floatexp_level(d))
    h = m.host_tables(nr // 2, nc // 2)
    od = oracles.OracleDevice()
    res = pipeline.render_rounds(od, mk(h), lambda gp: mk(m.host_tables(gp // nc, gp % nc)), nc, np.arange(nr), max_secondary=2)
    ex = sum(s["executed_iters"] for s in res["stats"])
    return od.out["iterations"], h["M"], ex


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--start-re", default="-0.75")
    ap.add_argument("--start-im", default="1e-5")
    ap.add_argument("--start-depth", type=int, default=3, help="first level: D = 10^-k")
    ap.add_argument("--depth", type=int, default=100)
    ap.add_argument("--N", type=int, default=1 << 20)
    ap.add_argument("--nr", type=int, default=24)
    ap.add_argument("--nc", type=int, default=32)
    ap.add_argument("--step-decades", type=int, default=1, help="zoom factor per level = 10^this")
    ap.add_argument("--pick", type=float, default=1.0, help="quantile of the finite counts to re-centre on (1 = max)")
    a = ap.parse_args()
    cre, cim = Fraction(a.start_re), Fraction(a.start_im)
    nr, nc, N = a.nr, a.nc, a.N
    digits = a.depth + 25
    levels = list(range(a.start_depth, a.depth, a.step_decades)) + [a.depth]
    for k in levels:
        D = Fraction(1, 10 ** k)
        pr, pi = 4 * D / nc, 3 * D / nr
        t0 = time.time()
        it, M, ex = render(nr, nc, N, pr, pi, cre, cim, digits)
        fin = it[(it < N) & (it >= 0)]
        interior = float((it >= N).mean())
        line = f"k={k:3d} M={M:7d} counts {fin.min() if len(fin) else -1}..{fin.max() if len(fin) else -1} " \
               f"mean {fin.mean() if len(fin) else 0:.0f} interior {interior:.2f} executed {ex} {time.time() - t0:.1f}s"
        print(line, flush=True)
        if k == a.depth:
            break
        # re-centre: a high finite count, as far from interior samples as ties allow
        cand = np.argwhere((it < N) & (it >= 0))
        if len(cand) == 0:
            print("all interior; stop")
            break
        vals = it[cand[:, 0], cand[:, 1]]
        target = np.quantile(vals, a.pick)
        order = np.argsort(np.abs(vals - target), kind="stable")
        r, c = cand[order[0]]
        cre = cre + (int(c) - nc // 2) * pr        # pixel map of mandelbrot.cpp:271, 275
        cim = cim + (nr // 2 - int(r) - 1) * pi
    print("center_re", dec(cre, digits))
    print("center_im", dec(cim, digits))


if __name__ == "__main__":
    main()
