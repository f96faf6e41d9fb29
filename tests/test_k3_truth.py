"""K3 (FP64 perturbation continuation) adjudicated against the reference and against the converged continuation.

tests/golden/k3_truth_*.npz (generator: tests/golden/make_k3_truth.py) hold, for the 6 144-sample set bench.py compares
on cfg2 and for all of KAT-S: the compiled reference's records (`ref`), the reference's own algorithm with its phase-3
continuation at 2x and at 4x/8x the view's precision (`t1`, `t1b` — equal everywhere, i.e. converged: the truth for
phase 3, reference mandelbrot.cpp:209-224) and brute force without the series skip (`direct`).

What they establish:
  * cfg2 (192-bit view): the reference equals the converged continuation on ALL 6 144 samples. Where K3 differs
    (~0.8 % of the samples) K3 is wrong: FP64 perturbation carries a relative error of ~5e-15 in delta after 7 000
    iterations, and samples whose last few hundred iterations are chaotic amplify that by > 1e10.
  * KAT-S (64-bit floor of setPrecision): the reference itself is not converged (2.2 % of its counts are wrong); K3 is
    wrong on about as many, mostly the same samples.
The CPU tests below pin those levels for Oracle-P (== the CUDA path bit for bit); the GPU test pins the CUDA path."""
import json
import os

import numpy as np
import pytest

from oracles import KATS, RefView, Tables, have_ref, p_render_deep

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")
META = json.load(open(os.path.join(G, "k3_truth.json")))
needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref (compiled reference) not built on this box")


def load(name):
    return np.load(os.path.join(G, name))


def test_truth_fixture_is_converged():
    for fn, key in (("k3_truth_cfg2.npz", "cfg2"), ("k3_truth_kat_s.npz", "KAT-S")):
        z = load(fn)
        assert np.array_equal(z["t1"]["iterations"], z["t1b"]["iterations"]), "2x and 4x/8x precision disagree"
        assert np.array_equal(z["t1"]["smoothing"].view(np.uint32), z["t1b"]["smoothing"].view(np.uint32))
        assert META[key]["t1_converged_frac"] == 1.0
    z = load("k3_truth_cfg2.npz")
    # on cfg2 the compiled reference IS the converged continuation, sample for sample
    assert np.array_equal(z["ref"]["iterations"], z["t1b"]["iterations"])
    # ... while the series skip itself (error_tolerance 1e-10) moves half of the counts away from brute force: the parity
    # target is the reference's algorithm, not the Mandelbrot set
    assert 0.3 < (z["direct_it"] == z["t1b"]["iterations"]).mean() < 0.7


@needs_ref
def test_fixture_matches_the_compiled_reference_live():
    from newman_b200 import workloads
    cfg = workloads.config("cfg2")
    z = load("k3_truth_cfg2.npz")
    v = RefView(cfg["nr"], cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
    v.precompute_at(*[int(x) for x in z["probe"]])
    sel = np.arange(0, len(z["pix"]), 96)
    out, _ = v.compute_pixels(z["pix"][sel])
    assert np.array_equal(out["iterations"], z["ref"]["iterations"][sel])
    assert np.array_equal(out["smoothing"].view(np.uint32), z["ref"]["smoothing"].view(np.uint32)[sel])


def adjudicate(port, z):
    t = z["t1b"]["iterations"]
    r = z["ref"]["iterations"]
    ok = port >= 0
    return dict(n=int(ok.sum()), port_agree=float((port[ok] == t[ok]).mean()), ref_agree=float((r[ok] == t[ok]).mean()),
                port_max_abs_diff=int(np.abs(port[ok] - t[ok]).max()), ref_max_abs_diff=int(np.abs(r[ok] - t[ok]).max()),
                both_wrong=int(((port != t) & (r != t) & ok).sum()), port_wrong=int(((port != t) & ok).sum()),
                ref_wrong=int(((r != t) & ok).sum()))


@needs_ref
def test_kat_s_port_vs_truth():
    """64-bit view: neither the reference nor FP64 perturbation is converged; they are wrong about equally often."""
    k = KATS["KAT-S"]
    v = RefView(**k)
    v.precompute()
    t = v.tables()
    er, ei = v.eps()
    z = load("k3_truth_kat_s.npz")
    out, rq, _, _ = p_render_deep(t, er, ei, mode=1)      # single rebasing pass: every sample resolved
    a = adjudicate(out["iterations"].reshape(-1), z)
    print("KAT-S", a)
    assert a["ref_wrong"] == 26 and a["ref_max_abs_diff"] == 442
    assert a["port_wrong"] <= a["ref_wrong"] + 4
    assert a["both_wrong"] >= 20    # the same chaotic samples


@needs_ref
def test_cfg2_port_vs_truth_subsample():
    """192-bit view: the reference is converged; FP64 perturbation misses ~0.8 % of the samples (768 of the 6 144 here;
    the GPU test checks all of them on the CUDA path)."""
    from newman_b200 import workloads
    cfg = workloads.config("cfg2")
    z = load("k3_truth_cfg2.npz")
    v = RefView(cfg["nr"], cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
    v.precompute_at(*[int(x) for x in z["probe"]])
    t = v.tables()
    er, ei = v.eps()
    sel = np.arange(0, len(z["pix"]), 8)
    pix = np.ascontiguousarray(z["pix"][sel])
    out, rq, _, _ = p_render_deep(t, er, ei, pix_list=pix, mode=1)
    sub = {k: z[k][sel] for k in ("t1b", "ref")}
    a = adjudicate(out.reshape(-1)[pix]["iterations"], sub)
    print("cfg2 (768 samples)", a)
    assert a["ref_wrong"] == 0
    assert a["port_agree"] >= 0.985 and a["port_max_abs_diff"] <= 400


@needs_ref
def test_exact_mode_on_the_cpu_oracle():
    """The exact mode restated on the CPU (Oracle-P == the CUDA path bit for bit): the probe — the sample set rendered against
    the rounded and against the truncated orbit — flags every sample the FP64 pass gets wrong, and the double-double pass
    (oraclep_refine_dd) gives the flagged samples the count of the converged continuation, i.e. the compiled reference's."""
    from newman_b200 import workloads
    from oracles import p_refine_dd
    cfg = workloads.config("cfg2")
    z = load("k3_truth_cfg2.npz")
    v = RefView(cfg["nr"], cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
    v.precompute_at(*[int(x) for x in z["probe"]])
    t = v.tables()
    er, ei = v.eps()
    sel = np.arange(0, len(z["pix"]), 3)
    pix = np.ascontiguousarray(z["pix"][sel])
    truth = z["t1b"]["iterations"][sel]
    a, _, _, _ = p_render_deep(t, er, ei, pix_list=pix, mode=1)
    t.orbit_truncated = True
    b, _, _, _ = p_render_deep(t, er, ei, pix_list=pix, mode=1)
    t.orbit_truncated = False
    a, b = a.reshape(-1)[pix]["iterations"], b.reshape(-1)[pix]["iterations"]
    wrong, flagged = a != truth, a != b
    print("cfg2 (2 048 samples): FP64 wrong", int(wrong.sum()), " flagged by the probe", int(flagged.sum()))
    assert wrong.sum() > 0 and not (wrong & ~flagged).any()
    out, st = p_refine_dd(t, er, ei, pix[flagged])
    fixed = a.copy()
    fixed[flagged] = out.reshape(-1)[pix[flagged]]["iterations"]
    assert np.array_equal(fixed, truth)
