"""-m gpu: the exact mode (Mandelbrot::exact / NM_MODE_DD): samples whose escape count FP64 perturbation cannot resolve are found
by rendering the frame a second time against the truncated orbit and repeated in double-double arithmetic (csrc/k3_dd.cuh).
Parity of the double-double pass: bit for bit against its CPU restatement (oracle/oracle_p.c: oraclep_refine_dd); of the mode
as a whole: against the converged continuation of the reference's own algorithm (tests/golden/k3_truth_cfg2.npz), which the
compiled reference equals on every sample."""
import ctypes as C
import os

import numpy as np
import pytest

import newman_b200
import oracles
from newman_b200 import _lib as L
from newman_b200 import workloads
from oracles import KATS, RefView

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
needs_ref = pytest.mark.skipif(not oracles.have_ref(), reason="oracle/_ref not built")


def refine_dd_oracle(t, er, ei, er_lo, ei_lo, pix, nr, nc):
    out, st = oracles.p_refine_dd(t, er, ei, pix, eps_lo=(er_lo, ei_lo))
    return out, st


@needs_ref
@pytest.mark.parametrize("kat", ["KAT-D30", "KAT-D90", "KAT-S"])
def test_dd_pass_bit_exact_vs_oracle(dev, kat):
    """NM_MODE_DD on every second sample of a fixture == oraclep_refine_dd, records and executed steps; the other samples
    keep the values of the frame before."""
    k = KATS[kat]
    v = RefView(**k)
    v.precompute()
    t = v.tables()
    er, ei = v.eps()
    nr, nc = k["nr"], k["nc"]
    rng = np.random.default_rng(3)
    er_lo = er * rng.uniform(-1, 1, nc) * 2.0 ** -54        # any low parts will do: both sides must use them alike
    ei_lo = ei * rng.uniform(-1, 1, nr) * 2.0 ** -54
    base = dev.render_deep(dev.make_tables(t.x_hi, t.x_lo, t.a, t.b, t.c, t.N, t.tol, t.glitch_tol), er, ei, mode=L.MODE_REBASE).copy()
    pix = np.arange(0, nr * nc, 2, dtype=np.int32)
    want, st = refine_dd_oracle(t, er, ei, er_lo, ei_lo, pix, nr, nc)
    tabs = dev.make_tables(t.x_hi, t.x_lo, t.a, t.b, t.c, t.N, t.tol, t.glitch_tol, eps_lo=(er_lo, ei_lo))
    dev.frame_deep(tabs, er, ei, pix_list=pix, mode=L.MODE_DD)
    dev.launch()
    got = dev.read_rows()
    gs = dev.stats()
    refined = want.reshape(-1)[pix]["iterations"] != 0          # (phase-2 samples are left to K2 by both)
    g, w, b = got.reshape(-1), want.reshape(-1), base.reshape(-1)
    sel = pix[refined]
    assert np.array_equal(g[sel]["iterations"], w[sel]["iterations"])
    assert np.array_equal(g[sel]["smoothing"].view(np.uint32), w[sel]["smoothing"].view(np.uint32))
    rest = np.setdiff1d(np.arange(nr * nc), pix)
    assert np.array_equal(g[rest].view(np.uint8), b[rest].view(np.uint8))
    assert gs["executed_iters"] == st["executed_iters"] and gs["rebased"] == st["rebased"]


def test_exact_mode_resolves_every_adjudicated_sample_of_cfg2():
    """The full-size cfg2 frame in exact mode: every one of the 6 144 adjudicated samples carries the count of the
    reference's converged continuation (== the compiled reference); the plain frame misses 29 of them."""
    cfg = workloads.config("cfg2")
    z = np.load(os.path.join(HERE, "golden", "k3_truth_cfg2.npz"))
    truth = z["t1b"]["iterations"]
    m = newman_b200.Mandelbrot(cfg["nr"], cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
    plain = m.render().copy()
    i0 = m.frame_info()
    assert (i0["probe_row"], i0["probe_col"]) == tuple(int(x) for x in z["probe"])
    m.set_exact(True)
    exact = m.render()
    i1 = m.frame_info()
    p6, e6 = plain.reshape(-1)[z["pix"]]["iterations"], exact.reshape(-1)[z["pix"]]["iterations"]
    print("cfg2: plain frame %d of %d sampled counts differ from the truth; exact mode %d. refined %d samples (%.2f %% of the frame), "
          "device %.1f + %.1f ms" % ((p6 != truth).sum(), len(truth), (e6 != truth).sum(), i1["refined"],
                                     100.0 * i1["refined"] / plain.size, i1["device_ms"], i1["refine_ms"]))
    bad = np.nonzero(e6 != truth)[0]
    print("   still different (sample id, plain, exact, truth):", [(int(z["pix"][i]), int(p6[i]), int(e6[i]), int(truth[i])) for i in bad])
    assert (p6 != truth).sum() > 0
    assert np.array_equal(e6, truth)
    changed = int((plain["iterations"] != exact["iterations"]).sum())
    assert 0 < changed <= i1["refined"] < 0.03 * plain.size
    m.set_exact(False)
    again = m.render()
    assert np.array_equal(again.view(np.uint8), plain.view(np.uint8))
