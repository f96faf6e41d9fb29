"""-m gpu: the run-to-completion kernel (newman_b200/csrc/k3_finish.cuh) that takes frames — and what is left
of a frame after a sweep — with few states instead of the level launches. Same decisions as the level kernels
and the oracle, bit for bit: rasters, glitch lists, executed-iteration and rebase counts; plain, floatexp-series
and scaled forms; both modes; as the whole frame, as the remainder of a k3_fast sweep, and through the drop-in
class (where it also serves the secondary-reference rounds). The other GPU modules pin NM_OPT_K3_FINISH_MAX to 0
at the device level so that their small fixtures keep exercising the level kernels."""
import numpy as np
import pytest

import newman_b200
import oracles
from newman_b200 import _lib as L
from oracles import KATS, RefView, p_render_deep

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not oracles.have_ref(), reason="oracle/_ref not built (needs /root/reference once)")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def kat_inputs(kat):
    k = KATS[kat]
    v = RefView(**k)
    v.precompute()
    er, ei = v.eps()
    return v.tables(), er, ei


def check(dev, tt, a, b, mode, base):
    exp, rq_pix, rq_it, st = base
    tabs = dev.make_tables(tt.x_hi, tt.x_lo, tt.a, tt.b, tt.c, tt.N, tt.tol, tt.glitch_tol, exps=tt.exps, eps_exps=tt.eps_exps)
    out = dev.render_deep(tabs, a, b, mode=mode)
    gpix, git = dev.requeue()
    gs = dev.stats()
    assert np.array_equal(out["iterations"], exp["iterations"])
    assert np.array_equal(bits(out["smoothing"]), bits(exp["smoothing"]))
    o1 = np.argsort(gpix); o2 = np.argsort(rq_pix)
    assert np.array_equal(gpix[o1], rq_pix[o2]) and np.array_equal(git[o1], rq_it[o2])
    assert gs["executed_iters"] == st["executed_iters"] and gs["rebased"] == st["rebased"]
    return gs


@needs_ref
@pytest.mark.parametrize("finish_max", [65536, 300])
@pytest.mark.parametrize("kat", ["KAT-D30", "KAT-D90", "KAT-S", "KAT-B"])
def test_finish_kernel_vs_oraclep(kat, finish_max):
    """finish_max = 65536: the whole frame in one launch. 300: the frame (1 200 .. 12 288 states) goes through
    the level kernels and only the event queue of the sweep is run to completion."""
    t, er, ei = kat_inputs(kat)
    dev = newman_b200.Device(0)
    try:
        dev.set_option(L.OPT_K3_FINISH_MAX, finish_max)
        for mode in (0, 1):
            base = p_render_deep(t, er, ei, mode=mode)
            gs = check(dev, t, er, ei, mode, base)
            if finish_max == 65536:
                assert gs["kernel_launches"] <= 12, gs   # K2 (+ its filter preparation) and ONE K3 launch
            tf = t.floatexp()
            ts, mr, mi = t.floatexp(er, ei)
            check(dev, tf, er, ei, mode, base)     # floatexp series
            check(dev, ts, mr, mi, mode, base)     # + floatexp eps, scaled delta states
    finally:
        dev.close()


def mk(k):
    return newman_b200.Mandelbrot(k["nr"], k["nc"], N=k["N"], sz=k["sz"], center=k["center"], tol=k.get("tol", 1e-10))


@needs_ref
@pytest.mark.parametrize("kat", ["KAT-D30", "KAT-S", "KAT-T3"])
def test_class_frames_agree_with_and_without_finish(kat, monkeypatch):
    """The drop-in class (which owns its context: the environment sets its default) renders the same frame
    through the level kernels alone as with the production default, where these small frames and their
    secondary-reference rounds run to completion in single launches."""
    k = KATS[kat]
    monkeypatch.setenv("NM_K3_FINISH_MAX", "0")
    m0 = mk(k)
    a = m0.render()
    ia = m0.frame_info()
    monkeypatch.delenv("NM_K3_FINISH_MAX")
    m1 = mk(k)
    b = m1.render()
    ib = m1.frame_info()
    assert np.array_equal(a["iterations"], b["iterations"]) and np.array_equal(bits(a["smoothing"]), bits(b["smoothing"]))
    assert (b["iterations"] >= 0).all()
    for key in ("references", "executed_iters", "glitched", "rebased", "orbit_len", "probe_row", "probe_col"):
        assert ia[key] == ib[key], key
    assert ib["kernel_launches"] < ia["kernel_launches"]


def test_finish_option_validation():
    dev = newman_b200.Device(0)
    try:
        with pytest.raises(newman_b200.NmError):
            dev.set_option(L.OPT_K3_FINISH_MAX, -1)
        dev.set_option(L.OPT_K3_FINISH_MAX, 0)
        dev.set_option(L.OPT_K3_FINISH_MAX, 1 << 20)
    finally:
        dev.close()


@needs_ref
@pytest.mark.parametrize("group", [4, 2])
@pytest.mark.parametrize("kat", ["KAT-D30", "KAT-D90", "KAT-S"])
def test_split_levels_vs_oraclep(kat, group):
    """NM_OPT_K3_SPLIT: chunks that follow an escape-heavy chunk run as four quarter-chunk launches with a global
    compaction after each (k3_fast.cuh: K3Work). With the state-count minimum lowered to 2 the small fixtures
    take that path wherever samples escape; rasters, glitch lists and counts must not change — plain, floatexp
    series and scaled forms."""
    t, er, ei = kat_inputs(kat)
    dev = newman_b200.Device(0)
    try:
        dev.set_option(L.OPT_K3_FINISH_MAX, 0)
        dev.set_option(L.OPT_K3_GROUP, group)
        base = p_render_deep(t, er, ei, mode=0)
        launches = []
        for split in (0, 2):
            dev.set_option(L.OPT_K3_SPLIT, split)
            gs = check(dev, t, er, ei, 0, base)
            launches.append(gs["kernel_launches"])
            ts, mr, mi = t.floatexp(er, ei)
            check(dev, t.floatexp(), er, ei, 0, base)
            check(dev, ts, mr, mi, 0, base)
        assert launches[1] > launches[0]
    finally:
        dev.close()

