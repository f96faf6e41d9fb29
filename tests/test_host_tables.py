"""Host side of precompute() (probe search, orbit, series, eps, coordinates, cardioid, view
transforms) against the compiled reference, bit for bit, and against the committed digests."""
import hashlib
import json
import os

import numpy as np
import pytest

import newman_b200
import oracles
from oracles import KATS, RefView

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "kats.json")))
needs_ref = pytest.mark.skipif(not oracles.have_ref(), reason="oracle/_ref not built")
DEEP = ["KAT-D30", "KAT-D60", "KAT-D90", "KAT-T3", "KAT-S"]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def mk(k, **kw):
    return newman_b200.Mandelbrot(k["nr"], k["nc"], N=k["N"], sz=k.get("sz"), center=k.get("center"),
                                  tol=k.get("tol", 1e-10), **kw)


@pytest.mark.parametrize("kat", DEEP)
def test_tables_match_golden_digests(kat):
    """findProbe + computeOrbit + computeSeries + eps, descended: equal to the reference's (digests
    generated from Oracle-R by tests/golden/make_golden.py)."""
    g = GOLD[kat]
    h = mk(KATS[kat]).host_tables()
    assert h["M"] == g["M"] and h["has_escape"] == g["has_escape"]
    for n in ("x_hi", "x_lo", "a", "b", "c"):
        assert digest(h[n]) == g["tables_sha256"][n], n
    assert [digest(h["eps_re"]), digest(h["eps_im"])] == g["eps_sha256"]
    assert mk(KATS[kat]).precision_bits() == g["precision_bits"]


@needs_ref
@pytest.mark.parametrize("kat", DEEP)
@pytest.mark.parametrize("threads", [1, 5])
def test_tables_match_reference_live(kat, threads):
    k = KATS[kat]
    v = RefView(**k)
    v.precompute()
    t = v.tables()
    er, ei = v.eps()
    h = mk(k, host_threads=threads).host_tables()
    assert h["M"] == t.M and h["has_escape"] == t.has_escape
    for n in ("x_hi", "x_lo", "a", "b", "c"):
        assert np.array_equal(h[n].view(np.uint64), getattr(t, n).view(np.uint64)), n
    assert np.array_equal(h["eps_re"].view(np.uint64), er.view(np.uint64))
    assert np.array_equal(h["eps_im"].view(np.uint64), ei.view(np.uint64))


@needs_ref
def test_forced_reference_point_matches():
    k = KATS["KAT-D60"]
    v = RefView(**k)
    v.precompute_at(10, 17)
    t = v.tables()
    h = mk(k).host_tables(10, 17)
    assert h["M"] == t.M
    for n in ("x_hi", "x_lo", "a", "b", "c"):
        assert np.array_equal(h[n].view(np.uint64), getattr(t, n).view(np.uint64)), n


@needs_ref
@pytest.mark.parametrize("kat", ["KAT-1c", "KAT-1b", "KAT-1"])
def test_hw_coords_and_cardioid(kat):
    k = KATS[kat]
    v = RefView(**k)
    m = mk(k)
    assert m.useHardware() and v.use_hardware()
    a, b = v.coords(), m.host_coords()
    assert np.array_equal(a[0].view(np.uint64), b[0].view(np.uint64))
    assert np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64))
    rng = np.random.default_rng(1)
    for r, c in zip(rng.integers(0, k["nr"], 200), rng.integers(0, k["nc"], 200)):
        assert m.host_in_cardioid(int(r), int(c)) == v.in_cardioid(int(r), int(c))


def test_cardioid_classification_deep():
    # far from the cardioid: decided once for the whole view
    mode, _ = mk(KATS["KAT-D30"]).host_cardioid()
    assert mode == newman_b200.CARDIOID_NONE
    # a deep view well inside the main cardioid: every sample is interior
    m = newman_b200.Mandelbrot(16, 16, N=100, sz=("1e-30", "1e-30"), center=("-0.1", "0.1"))
    mode, _ = m.host_cardioid()
    assert mode == newman_b200.CARDIOID_ALL
    # a deep view straddling the cardioid boundary on the real axis (c = 0.25 is the cusp)
    m = newman_b200.Mandelbrot(8, 8, N=100, sz=("1e-30", "1e-30"), center=("0.25", "0"))
    mode, mask = m.host_cardioid()
    assert mode == newman_b200.CARDIOID_MASK
    assert 0 < mask.sum() < mask.size


@needs_ref
def test_cardioid_mask_matches_reference():
    kw = dict(nr=8, nc=8, N=100, sz=("1e-30", "1e-30"), center=("0.25", "0"))
    v = RefView(**kw)
    m = newman_b200.Mandelbrot(8, 8, N=100, sz=kw["sz"], center=kw["center"])
    mode, mask = m.host_cardioid()
    assert np.array_equal(mask, v.cardioid_mask())


@needs_ref
def test_view_transforms_match_reference(tmp_path):
    """zoom / translate / zoomAt / scaleUp / scaleDown / loadLegacy leave bit-identical view state."""
    v = RefView(60, 80, N=300)
    m = newman_b200.Mandelbrot(60, 80, N=300)
    L = RefView.lib()
    steps = [("zoom", (3.0,)), ("translate", (5, -7, 1)), ("zoom_at", (32.0, 10, 20, 1)), ("zoom", (1e6,)),
             ("zoom_at", (1e9, 40, 3, 1)), ("translate", (-2, 9, 2)), ("zoom", (2.5e7,)), ("zoom_at", (4e8, 1, 1, 2))]
    for name, args in steps:
        if name == "zoom":
            L.ref_zoom(v.h, args[0]); m.zoom(args[0])
        elif name == "translate":
            L.ref_translate(v.h, *args); m.translate(*args)
        else:
            L.ref_zoom_at(v.h, *args); m.zoomAt(*args)
        assert m.view_strings() == v.view_strings(), (name, args)
        assert m.precision_bits() == v.precision_bits()
    assert not m.useHardware()
    # multisample rescale of the raster (mandelbrot.cpp:320-360), incl. the float32 accumulator
    rng = np.random.default_rng(7)
    g = np.zeros((60, 80), dtype=newman_b200.ESCAPE_DTYPE)
    g["iterations"] = rng.integers(0, 300, g.shape)
    g["smoothing"] = rng.random(g.shape, dtype=np.float32)
    v2 = RefView(60, 80, N=300)
    m2 = newman_b200.Mandelbrot(60, 80, N=300)
    m2.set_grid(g)
    # push the same raster into the reference through its own grid memory
    import ctypes as C
    L.ref_write_grid.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_write_grid(v2.h, oracles.vp(g))
    for sc, up in ((2, 0), (3, 1), (3, 0), (4, 1), (4, 0), (2, 1)):
        L.ref_scale(v2.h, sc, up)
        (m2.scaleUp if up else m2.scaleDown)(sc)
        assert (m2.rows(), m2.cols()) == (L.ref_rows(v2.h), L.ref_cols(v2.h))
        rg = np.zeros((m2.rows(), m2.cols()), dtype=newman_b200.ESCAPE_DTYPE)
        L.ref_read_grid(v2.h, oracles.vp(rg))
        assert np.array_equal(m2.grid().view(np.uint64), rg.view(np.uint64)), (sc, up)
        assert m2.view_strings() == v2.view_strings()
    # legacy 5-line view file
    fn = str(tmp_path / "view.txt")
    open(fn, "w").write("5000\n-0.743643887037158704752191506114774\n0.131825904205311970493132056385139\n"
                        "1.25e-20\n1.25e-20\n")
    v3 = RefView(60, 80)
    m3 = newman_b200.Mandelbrot(60, 80)
    L.ref_load_legacy(v3.h, fn.encode())
    m3.loadLegacy(fn)
    assert m3.view_strings() == v3.view_strings()
    # save() writes the viewer's format (viewer.cpp:12-23); it round-trips at the viewer's 800x600
    m5 = newman_b200.Mandelbrot(600, 800, N=777, sz=("3e-40", "3e-40"), center=("-1.25", "0.0625"))
    fn2 = str(tmp_path / "view2.txt")
    m5.save(fn2)
    m6 = newman_b200.Mandelbrot(600, 800)
    m6.zoom(1e30)  # enough precision to parse the file (the reference's loader has the same caveat)
    m6.loadLegacy(fn2)
    assert m6.view_strings() == m5.view_strings()


ALL_TABLES = ("x_hi", "x_lo", "a", "b", "c", "a_m", "a_e", "b_m", "b_e", "c_m", "c_e", "eps_re", "eps_im", "eps_re_m",
              "eps_re_e", "eps_im_m", "eps_im_e")


def same_tables(h1, h4):
    assert (h1["M"], h1["has_escape"], h1["finite"]) == (h4["M"], h4["has_escape"], h4["finite"])
    for n in ALL_TABLES:
        a, b = np.ascontiguousarray(h1[n]), np.ascontiguousarray(h4[n])
        assert a.shape == b.shape and a.tobytes() == b.tobytes(), n


@pytest.mark.parametrize("name,scale", [("cfg2", 40), ("cfg3", 80)])
def test_pipelined_tables_equal_serial_tables(name, scale):
    """build_tables with 4+ host threads runs orbit, A, B, C as a pipeline (hp_host.cpp: tables_pipelined; 6+ threads: two
    more stages for the products); it must
    deliver what the one-thread form does, bit for bit — the long orbits of the bench views, double and
    mantissa/exponent forms."""
    from newman_b200 import workloads
    cfg = workloads.config(name, scale=scale)
    hs = []
    for threads in (1, 4, 6):   # serial, 4 stages, 6 stages (A^2 and A B products on their own threads)
        v = newman_b200.Mandelbrot(cfg["nr"], cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"],
                                   host_threads=threads)
        hs.append(v.host_tables(cfg["nr"] // 2, cfg["nc"] // 2))
    assert hs[0]["M"] > 50000
    same_tables(hs[0], hs[1])
    same_tables(hs[0], hs[2])


@pytest.mark.parametrize("N", [1, 2, 3, 33, 4097, 9000])
def test_pipelined_tables_edge_lengths(N):
    """Orbits that never escape (M == N: a reference inside the main cardioid, lengths around the pipeline's block
    and batch sizes) and ones that escape at once."""
    for center in (("-0.1", "0.05"), ("0.4", "0.3"), ("2.5", "0.0")):
        hs = []
        for threads in (1, 4, 6):
            v = newman_b200.Mandelbrot(12, 16, N=N, sz=("1e-30", "1e-30"), center=center, host_threads=threads)
            hs.append(v.host_tables(6, 8))
        same_tables(hs[0], hs[1])
        same_tables(hs[0], hs[2])
        if center[0] == "-0.1":
            assert hs[0]["M"] == N and not hs[0]["has_escape"]


def test_pooled_mpf_layout_selfcheck():
    """hp_host.cpp lays arbitrary-precision values out by hand in limb pools (the pipelined table build); the library
    checks once per process that such values behave exactly like mpf_init2 values under the running libgmp."""
    from newman_b200 import view
    assert view._lib().nmv_host_selfcheck() == 1
