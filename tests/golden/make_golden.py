"""Generate the committed golden fixtures from Oracle-R (the reference's own mandelbrot.cpp compiled
unmodified, oracle/_ref). Run HERE (needs /root/reference once):  python tests/golden/make_golden.py

  kats.json       per known-answer view: inputs, orbit length, precision, sums and sha256 digests of the
                  reference raster (iterations / smoothing planes) and of the Oracle-P raster
  kat_1c.npz      48x64 plain-double view: coordinates + reference raster
  kat_d30.npz     96x128 view at 1e-30: descended tables, eps arrays, reference raster, Oracle-P raster
  kat_s.npz       30x40 seahorse view (M=9082): same, exercises long continuation / rebasing / glitches
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracles import KATS, RefView, p_render_deep, p_render_hw  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    meta = {}
    for name, k in KATS.items():
        v = RefView(**k)
        hw = v.use_hardware()
        if not hw:
            v.precompute()
        ref, _ = v.render_all()
        m = dict(inputs={kk: vv for kk, vv in k.items()}, hardware=hw, precision_bits=v.precision_bits(),
                 sum_iterations=int(ref["iterations"].sum(dtype=np.int64)), min_it=int(ref["iterations"].min()),
                 max_it=int(ref["iterations"].max()), ref_iterations_sha256=digest(ref["iterations"]),
                 ref_smoothing_sha256=digest(ref["smoothing"]))
        if hw:
            cre, cim = v.coords()
            N = k["N"]
            m["executed"] = int(((ref["iterations"] + 1) * (ref["iterations"] < N)).sum())
            if name == "KAT-1c":
                np.savez_compressed(os.path.join(HERE, "kat_1c.npz"), c_re=cre, c_im=cim, ref=ref,
                                    cardioid=v.cardioid_mask())
        else:
            t = v.tables()
            er, ei = v.eps()
            out, rq_pix, rq_it, st = p_render_deep(t, er, ei)
            m.update(M=t.M, has_escape=t.has_escape, oraclep_iterations_sha256=digest(out["iterations"]),
                     oraclep_smoothing_sha256=digest(out["smoothing"]), oraclep_glitched=int(len(rq_pix)),
                     oraclep_stats={kk: int(vv) for kk, vv in st.items()},
                     count_mismatch_vs_ref=int(((out["iterations"] != ref["iterations"]) & (out["iterations"] >= 0)).sum()),
                     tables_sha256={n: digest(getattr(t, n)) for n in ("x_hi", "x_lo", "a", "b", "c")},
                     eps_sha256=[digest(er), digest(ei)])
            if name in ("KAT-D30", "KAT-S"):
                fn = "kat_d30.npz" if name == "KAT-D30" else "kat_s.npz"
                np.savez_compressed(os.path.join(HERE, fn), x_hi=t.x_hi, x_lo=t.x_lo, a=t.a, b=t.b, c=t.c,
                                    eps_re=er, eps_im=ei, N=t.N, tol=t.tol, glitch_tol=t.glitch_tol, ref=ref,
                                    oraclep=out, rq_pix=rq_pix, rq_iter=rq_it)
        meta[name] = m
        print(name, {kk: vv for kk, vv in m.items() if kk not in ("inputs", "tables_sha256")})
    json.dump(meta, open(os.path.join(HERE, "kats.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
