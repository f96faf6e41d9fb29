"""Adjudication fixtures for K3 (the FP64 perturbation continuation) against the reference.

K3 replaces phase 3 of Mandelbrot::getIterations (reference mandelbrot.cpp:209-224: `Yn = Y^2 + Y0` in mpf from
`Y = X[L-1] + d[L-1]`). Where its escape counts differ from the compiled reference's, who is right? For every sample
of (a) the strided 6 144-sample set bench.py compares on cfg2 and (b) all 1 200 samples of KAT-S this script records

  ref      the compiled reference (Oracle-R) at the view's own precision
  t1       the reference's OWN algorithm with phase 3 run at 2x the view's bits (oracle/ref_driver.cpp:
           ref_compute_pixels_wide; phases 1-2 are double arithmetic and do not change) — and again at 4x (t1b);
           t1 == t1b everywhere means the continuation is converged: that is the truth for phase 3
  direct   brute force from the pixel itself, no series skip, at 2x bits (ref_truth_pixels): how far the series
           truncation (error_tolerance) itself moves counts — context, not a parity target

Run HERE (needs /root/reference, ~6 min on 8 cores):  python tests/golden/make_k3_truth.py
Writes tests/golden/k3_truth_cfg2.npz, k3_truth_kat_s.npz and k3_truth.json (summary)."""
import ctypes as C
import json
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracles  # noqa: E402
from newman_b200 import workloads  # noqa: E402


def _lib():
    L = oracles.RefView.lib()
    L.ref_truth_pixels.restype = C.c_double
    L.ref_truth_pixels.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_compute_pixels_wide.restype = C.c_double
    L.ref_compute_pixels_wide.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return L


def bench_sample(nr, nc, n):
    """the strided sample of bench.py (cpu_reference_sample)"""
    total = nr * nc
    return (np.arange(n, dtype=np.int64) * (total // n) + (total // n) // 3).astype(np.int32)


def _work(a):
    kw, probe, pix, mult = a
    L = _lib()
    v = oracles.RefView(**kw)
    if probe is None:
        v.precompute()
    else:
        v.precompute_at(*probe)
    bits = v.precision_bits()
    ref, _ = v.compute_pixels(pix)
    out = {"ref": ref, "bits": bits, "M": v.orbit_len()}
    for key, m in (("t1", mult[0]), ("t1b", mult[1])):
        o = np.zeros(len(pix), dtype=oracles.ESC)
        L.ref_compute_pixels_wide(v.h, oracles.vp(pix), len(pix), m * bits, oracles.vp(o))
        out[key] = o
    ref2, _ = v.compute_pixels(pix)
    assert np.array_equal(ref, ref2), "the wide run disturbed the reference"
    it = np.zeros(len(pix), dtype=np.int32)
    r2 = np.zeros(len(pix))
    L.ref_truth_pixels(v.h, oracles.vp(pix), len(pix), mult[0] * bits, oracles.vp(it), oracles.vp(r2))
    out["direct_it"], out["direct_r2"] = it, r2
    return out


def run(kw, probe, pix, mult=(2, 4), procs=None):
    procs = procs or os.cpu_count() or 1
    chunks = [np.ascontiguousarray(pix[i::procs]) for i in range(procs)]
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(_work, [(kw, probe, c, mult) for c in chunks])
    out = {}
    for key in ("ref", "t1", "t1b", "direct_it", "direct_r2"):
        a = np.zeros(len(pix), dtype=res[0][key].dtype)
        for i, r in enumerate(res):
            a[i::procs] = r[key]
        out[key] = a
    out["bits"], out["M"] = res[0]["bits"], res[0]["M"]
    return out


def summary(name, pix, r):
    ref, t1, t1b = r["ref"]["iterations"], r["t1"]["iterations"], r["t1b"]["iterations"]
    d = {"n": int(len(pix)), "view_bits": int(r["bits"]), "orbit_len": int(r["M"]),
         "t1_converged_frac": float((t1 == t1b).mean()), "ref_equals_t1_frac": float((ref == t1b).mean()),
         "ref_max_abs_diff_vs_t1": int(np.abs(ref - t1b).max()),
         "direct_equals_t1_frac": float((r["direct_it"] == t1b).mean()),
         "direct_max_abs_diff_vs_t1": int(np.abs(r["direct_it"] - t1b).max())}
    print(name, d)
    return d


def main():
    meta = {}
    cfg = workloads.config("cfg2")
    kw = dict(nr=cfg["nr"], nc=cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
    pix = bench_sample(cfg["nr"], cfg["nc"], 6144)
    r = run(kw, tuple(cfg["probe"]), pix)
    np.savez_compressed(os.path.join(HERE, "k3_truth_cfg2.npz"), pix=pix, probe=np.array(cfg["probe"]), ref=r["ref"], t1=r["t1"],
                        t1b=r["t1b"], direct_it=r["direct_it"], direct_r2=r["direct_r2"])
    meta["cfg2"] = summary("cfg2", pix, r)
    k = oracles.KATS["KAT-S"]
    pix = np.arange(k["nr"] * k["nc"], dtype=np.int32)
    r = run(dict(k), None, pix, mult=(2, 8))
    np.savez_compressed(os.path.join(HERE, "k3_truth_kat_s.npz"), pix=pix, ref=r["ref"], t1=r["t1"], t1b=r["t1b"],
                        direct_it=r["direct_it"], direct_r2=r["direct_r2"])
    meta["KAT-S"] = summary("KAT-S", pix, r)
    json.dump(meta, open(os.path.join(HERE, "k3_truth.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
