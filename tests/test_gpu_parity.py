"""-m gpu parity tests proper: the CUDA path (through the C-ABI) against Oracle-P bit for bit, and
against Oracle-R (the compiled reference, oracle/_ref — travels to the GPU box prebuilt) wherever
the reference is defined."""
import numpy as np
import pytest

import newman_b200
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
@pytest.mark.parametrize("literal", [1, 0])
@pytest.mark.parametrize("kat", ["KAT-D30", "KAT-D60", "KAT-D90", "KAT-B", "KAT-T3", "KAT-S"])
def test_deep_bit_exact_vs_oraclep(dev, kat, literal):
    """Series + perturbation: GPU == Oracle-P bit for bit, including which pixels are flagged as
    glitched (order-independent) and the executed-iteration count. K2 runs both as the reference's
    literal index scan and with the per-index filter; both must give the oracle's L for every sample."""
    k = KATS[kat]
    v = RefView(**k)
    v.precompute()
    t = v.tables()
    er, ei = v.eps()
    exp, rq_pix, rq_it, st = p_render_deep(t, er, ei)
    tabs = dev.make_tables(t.x_hi, t.x_lo, t.a, t.b, t.c, t.N, t.tol, t.glitch_tol)
    dev.set_option(newman_b200._lib.OPT_K2_LITERAL, literal)
    try:
        out = dev.render_deep(tabs, er, ei)
    finally:
        dev.set_option(newman_b200._lib.OPT_K2_LITERAL, 0)
    gpix, git = dev.requeue()
    assert np.array_equal(out["iterations"], exp["iterations"])
    assert np.array_equal(bits(out["smoothing"]), bits(exp["smoothing"]))
    o1 = np.argsort(gpix); o2 = np.argsort(rq_pix)
    assert np.array_equal(gpix[o1], rq_pix[o2]) and np.array_equal(git[o1], rq_it[o2])
    gs = dev.stats()
    assert gs["executed_iters"] == st["executed_iters"]
    assert gs["rebased"] == st["rebased"]
    if literal:
        assert gs["series_evals"] == st["series_evals"]
    else:  # the filter leaves ~1 exact test per sample (KAT-B: tol = 1e9 is far outside its sweet spot)
        assert gs["series_evals"] <= max(4 * out.size, st["series_evals"] // 8)
        print(kat, "exact tests per sample: %.3f (literal %.1f)" % (gs["series_evals"] / out.size, st["series_evals"] / out.size))


@needs_ref
@pytest.mark.parametrize("group", [0, 2, 4])
@pytest.mark.parametrize("kat", ["KAT-D30", "KAT-D90", "KAT-S"])
def test_k3_kernel_variants_agree(dev, kat, group):
    """The simple (1 pixel/lane) and fast (2 or 4 same-index pixels/lane, branch-free blocks) kernels
    take identical decisions: rasters, glitch lists, executed-iteration and rebase counts."""
    k = KATS[kat]
    v = RefView(**k)
    v.precompute()
    t = v.tables()
    er, ei = v.eps()
    exp, rq_pix, rq_it, st = p_render_deep(t, er, ei)
    tabs = dev.make_tables(t.x_hi, t.x_lo, t.a, t.b, t.c, t.N, t.tol, t.glitch_tol)
    dev.set_option(newman_b200._lib.OPT_K3_GROUP, group)
    try:
        out = dev.render_deep(tabs, er, ei)
        gpix, git = dev.requeue()
        gs = dev.stats()
    finally:
        dev.set_option(newman_b200._lib.OPT_K3_GROUP, 4)
    assert np.array_equal(out["iterations"], exp["iterations"])
    assert np.array_equal(bits(out["smoothing"]), bits(exp["smoothing"]))
    o1 = np.argsort(gpix); o2 = np.argsort(rq_pix)
    assert np.array_equal(gpix[o1], rq_pix[o2]) and np.array_equal(git[o1], rq_it[o2])
    assert gs["executed_iters"] == st["executed_iters"] and gs["rebased"] == st["rebased"]


def test_series_filter_irregular_ranges(dev):
    """Synthetic coefficient tables that push the filter into its unsafe / literal branches (tiny and
    huge |B|,|C|, zeros, denormal products): filtered K2 must still reproduce the literal scan."""
    rng = np.random.default_rng(11)
    M, N = 400, 1000
    nr, nc = 24, 32
    for trial, (eps_mag, growth, tol) in enumerate([(1e-40, 0.9, 1e-10), (1e-95, 2.0, 1e-10), (1e-3, 0.2, 1e-10),
                                                    (1e-60, 1.4, 1e5), (1e-102, 2.2, 1e-10), (1e-20, 0.0, 1e-300)]):
        i = np.arange(M)
        mag = 10.0 ** np.clip(growth * i, None, 300)
        ph = rng.random((3, M)) * 2 * np.pi
        a = np.stack([mag * np.cos(ph[0]), mag * np.sin(ph[0])], 1).reshape(-1)
        b = np.stack([np.minimum(mag ** 2, 1e300) * np.cos(ph[1]), np.minimum(mag ** 2, 1e300) * np.sin(ph[1])], 1).reshape(-1)
        c = np.stack([np.minimum(mag ** 3, 1e305) * np.cos(ph[2]), np.minimum(mag ** 3, 1e305) * np.sin(ph[2])], 1).reshape(-1)
        for arr in (b, c):  # sprinkle exact zeros and tiny entries
            arr[rng.integers(2, 2 * M, 12)] = 0.0
            arr[rng.integers(2, 2 * M, 12)] *= 1e-200
        b[:2] = 0; c[:2] = 0; a[:2] = (1.0, 0.0)
        x_hi = np.zeros(2 * (M + 1)); x_hi[0::2] = 0.3 * np.cos(i_ := np.arange(M + 1) * 0.7); x_hi[1::2] = 0.3 * np.sin(i_)
        x_hi[-2:] = (2000.0, 0.0)
        x_lo = np.zeros(2 * M)
        er = (rng.random(nc) - 0.5) * eps_mag * 10.0 ** (-8 * rng.random(nc))
        ei = (rng.random(nr) - 0.5) * eps_mag * 10.0 ** (-8 * rng.random(nr))
        er[3] = 0.0; ei[5] = 0.0
        t = oracles.Tables(x_hi, x_lo, a, b, c, N, tol)
        exp, rq_pix, rq_it, st = p_render_deep(t, er, ei)
        tabs = dev.make_tables(t.x_hi, t.x_lo, t.a, t.b, t.c, t.N, t.tol, t.glitch_tol)
        for literal in (1, 0):
            dev.set_option(newman_b200._lib.OPT_K2_LITERAL, literal)
            try:
                out = dev.render_deep(tabs, er, ei)
            finally:
                dev.set_option(newman_b200._lib.OPT_K2_LITERAL, 0)
            assert np.array_equal(out["iterations"], exp["iterations"]), (trial, literal)
            assert np.array_equal(bits(out["smoothing"]), bits(exp["smoothing"])), (trial, literal)
            assert dev.stats()["executed_iters"] == st["executed_iters"], (trial, literal)  # sum of L over samples


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


@needs_ref
@pytest.mark.parametrize("group", [4, 2])
@pytest.mark.parametrize("kat", ["KAT-D30", "KAT-D90", "KAT-S"])
def test_post_link_pass_changes_no_result(dev, kat, group):
    """tools/sass_resched.py re-orders and re-registers k3_fast's FP64 blocks inside the shipped library; the same
    objects linked without it (libnewman_b200_ptxas.so, ptxas's own instruction order) must give the same raster, the
    same glitch list and the same counters on the same tables — every rounding is the same instruction."""
    import os
    from newman_b200 import _lib as L
    from newman_b200.device import Device
    other = os.path.join(os.path.dirname(L.LIB_PATH), "libnewman_b200_ptxas.so")
    if not os.path.exists(other):
        pytest.skip("libnewman_b200_ptxas.so not built")
    k = KATS[kat]
    v = RefView(**k)
    v.precompute()
    t = v.tables()
    er, ei = v.eps()
    got = []
    for d in (dev, Device(0, lib=L.load_from(other))):
        tabs = d.make_tables(t.x_hi, t.x_lo, t.a, t.b, t.c, t.N, t.tol, t.glitch_tol)
        d.set_option(L.OPT_K3_GROUP, group)
        try:
            out = d.render_deep(tabs, er, ei)
            gpix, git = d.requeue()
            gs = d.stats()
        finally:
            d.set_option(L.OPT_K3_GROUP, 4)
        o = np.argsort(gpix)
        got.append((out["iterations"].copy(), bits(out["smoothing"]).copy(), gpix[o].copy(), git[o].copy(),
                    gs["executed_iters"], gs["rebased"], gs["glitched"]))
        if d is not dev:
            d.close()
    a, b = got
    for x, y in zip(a[:4], b[:4]):
        assert np.array_equal(x, y)
    assert a[4:] == b[4:]
    assert a[4] > 0
