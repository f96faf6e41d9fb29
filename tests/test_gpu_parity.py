"""-m gpu parity tests proper: the CUDA path (through the C-ABI) against Oracle-P bit for bit, and
against Oracle-R (the compiled reference, oracle/_ref — travels to the GPU box prebuilt) wherever
the reference is defined."""
import numpy as np
import pytest

import oracles
from oracles import KATS, RefView, p_render_deep, p_render_hw

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not oracles.have_ref(), reason="oracle/_ref not built (needs /root/reference once)")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@needs_ref
@pytest.mark.parametrize("kat", ["KAT-1c", "KAT-1b", "KAT-1"])
def test_k1_bit_exact_vs_reference(dev, kat):
    """Plain-double path: iterations AND float32 smoothing bit-identical to the compiled reference."""
    k = KATS[kat]
    v = RefView(**k)
    assert v.use_hardware()
    ref, _ = v.render_all()
    cre, cim = v.coords()
    dev.frame_hw(cre, cim, k["N"])
    dev.launch()
    amb = dev.ambiguous()
    for pix in amb:  # host verdict with the reference's own mpf test
        r, c = divmod(int(pix), k["nc"])
        if v.in_cardioid(r, c):
            dev.poke(pix, k["N"], 0.0)
    out = dev.read_rows()
    assert np.array_equal(out["iterations"], ref["iterations"])
    assert np.array_equal(bits(out["smoothing"]), bits(ref["smoothing"]))
    st = dev.stats()
    exp_exec = int(((ref["iterations"] + 1) * (ref["iterations"] < k["N"])).sum())
    assert st["executed_iters"] >= exp_exec  # + N for every non-cardioid interior sample


@needs_ref
@pytest.mark.parametrize("kat", ["KAT-D30", "KAT-D60", "KAT-D90", "KAT-B", "KAT-T3", "KAT-S"])
def test_deep_bit_exact_vs_oraclep(dev, kat):
    """Series + perturbation: GPU == Oracle-P bit for bit, including which pixels are flagged as
    glitched (order-independent), executed-iteration count and series evaluations."""
    k = KATS[kat]
    v = RefView(**k)
    v.precompute()
    t = v.tables()
    er, ei = v.eps()
    exp, rq_pix, rq_it, st = p_render_deep(t, er, ei)
    tabs = dev.make_tables(t.x_hi, t.x_lo, t.a, t.b, t.c, t.N, t.tol, t.glitch_tol)
    out = dev.render_deep(tabs, er, ei)
    gpix, git = dev.requeue()
    assert np.array_equal(out["iterations"], exp["iterations"])
    assert np.array_equal(bits(out["smoothing"]), bits(exp["smoothing"]))
    o1 = np.argsort(gpix); o2 = np.argsort(rq_pix)
    assert np.array_equal(gpix[o1], rq_pix[o2]) and np.array_equal(git[o1], rq_it[o2])
    gs = dev.stats()
    assert gs["executed_iters"] == st["executed_iters"]
    assert gs["series_evals"] == st["series_evals"]
    assert gs["rebased"] == st["rebased"]


@needs_ref
@pytest.mark.parametrize("kat", ["KAT-D30", "KAT-S"])
def test_deep_rebase_mode_vs_oraclep(dev, kat):
    k = KATS[kat]
    v = RefView(**k)
    v.precompute()
    t = v.tables()
    er, ei = v.eps()
    exp, rq_pix, _, st = p_render_deep(t, er, ei, mode=1)
    assert len(rq_pix) == 0
    tabs = dev.make_tables(t.x_hi, t.x_lo, t.a, t.b, t.c, t.N, t.tol, t.glitch_tol)
    out = dev.render_deep(tabs, er, ei, mode=1)
    assert np.array_equal(out["iterations"], exp["iterations"])
    assert np.array_equal(bits(out["smoothing"]), bits(exp["smoothing"]))
    assert dev.stats()["rebased"] == st["rebased"]
    assert (out["iterations"] >= 0).all()


@needs_ref
def test_deep_vs_reference_counts(dev):
    """Where the reference's arbitrary-precision continuation is short, escape counts of every
    non-glitched pixel equal the compiled reference's exactly."""
    for kat in ["KAT-D30", "KAT-D60", "KAT-D90", "KAT-B"]:
        k = KATS[kat]
        v = RefView(**k)
        v.precompute()
        ref, _ = v.render_all()
        t = v.tables()
        er, ei = v.eps()
        tabs = dev.make_tables(t.x_hi, t.x_lo, t.a, t.b, t.c, t.N, t.tol, t.glitch_tol)
        out = dev.render_deep(tabs, er, ei)
        ok = out["iterations"] >= 0
        assert ok.mean() > 0.999
        assert np.array_equal(out["iterations"][ok], ref["iterations"][ok])


def test_fp64_peak_probe(dev):
    info = dev.info()
    assert info["sm_count"] >= 100
    ips, ms = dev.fp64_peak(0, 1 << 14)
    per_sm_clk = ips / (info["sm_count"] * info["sm_clock_khz"] * 1e3)
    print("DFMA inst/s %.3e  => %.1f lanes/clk/SM at max clock" % (ips, per_sm_clk))
    assert 8 < per_sm_clk < 140
