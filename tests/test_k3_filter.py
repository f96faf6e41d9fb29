"""The K3 candidate filter (newman_b200/csrc/k3_filter.cuh) must have NO FALSE NEGATIVES: whenever the exact
comparisons of the perturbation step (k3_checked.cuh == oracle/oracle_p.c:369-383) would report a glitch or
an escape, the integer test on the high words of delta that k3_fast evaluates instead has to fire. The host
build of the same functions is attacked here with states placed on and around the decision boundaries; the
exact side is evaluated with rational arithmetic rounded once (== IEEE fma / add / mul)."""
import ctypes as C
import math
from fractions import Fraction as Fr

import numpy as np
import pytest

import newman_b200

U5 = C.c_uint32 * 5


def entry(lib, zr, zi, gb):
    e = U5()
    lib.nm_k3_filter_entry(zr, zi, gb, e)
    return e


def fires(lib, e, dr, di, scaled=0):
    g, x = C.c_int(), C.c_int()
    lib.nm_k3_filter_fires(e, dr, di, scaled, C.byref(g), C.byref(x))
    return bool(g.value), bool(x.value)


def fl(x):
    return float(x)  # Fraction -> nearest double (ties to even)


def exact_tests(zr, zi, dr, di, gb, S=1.0):
    """(glitch, escape) exactly as checked_step decides them for the new delta paired with Z = (zr, zi)."""
    if S == 1.0:
        r, i = zr + dr, zi + di                        # DADD
    else:
        r, i = fl(Fr(S) * Fr(dr) + Fr(zr)), fl(Fr(S) * Fr(di) + Fr(zi))   # fma(S, d, Z)
    if not (math.isfinite(r) and math.isfinite(i)):
        return False, True
    p = r * r                                           # DMUL
    zmag = fl(Fr(i) * Fr(i) + Fr(p)) if math.isfinite(p) else math.inf   # fma(zi, zi, zr*zr)
    esc = zmag > 1048576.0
    return (not esc) and zmag < gb, esc


def glitch_bound(zr, zi, tol):
    return (zr * zr + zi * zi) * tol                    # k_glitch_bounds / oracle_p.c:314


def test_glitch_filter_has_no_false_negatives():
    lib = newman_b200.load()
    rng = np.random.default_rng(20261017)
    hits = 0
    for k in range(60000):
        mag = 10.0 ** rng.uniform(-120, 0.3)
        th = rng.uniform(0, 2 * math.pi)
        mode = k % 6
        if mode == 0:      # nearly axis-aligned: the minor component straddles zero
            th = rng.choice([0, 0.5, 1, 1.5]) * math.pi + rng.normal() * 10.0 ** rng.uniform(-12, -2)
        zr, zi = mag * math.cos(th), mag * math.sin(th)
        if mode == 1:
            zi = 0.0
        tol = float(rng.choice([1e-6, 1e-6, 1e-6, 1e-3, 1e-9, 1e-12, 0.3, 0.9]))
        gb = glitch_bound(zr, zi, tol)
        g = math.sqrt(gb)
        # delta = -Z + u with |u| around the glitch radius (inside, on, outside), any direction
        rad = g * float(rng.choice([0.0, 1e-3, 0.5, 0.999999, 1 - 1e-12, 1 - 1e-15, 1.0, 1 + 1e-15, 1 + 1e-9, 1.2]))
        ph = rng.uniform(0, 2 * math.pi) if k % 3 else rng.choice([0, 0.25, 0.5, 0.75, 1, 1.25, 1.5, 1.75]) * math.pi
        dr, di = -zr + rad * math.cos(ph), -zi + rad * math.sin(ph)
        e = entry(lib, zr, zi, gb)
        gl, _ = exact_tests(zr, zi, dr, di, gb)
        fg, fx = fires(lib, e, dr, di)
        if gl:
            hits += 1
            assert fg, ("glitch missed", zr, zi, dr, di, gb, list(e))
    assert hits > 20000   # the attack really produces glitches


def test_glitch_filter_interval_edges():
    """States whose high words sit exactly on the ends of the table's intervals."""
    lib = newman_b200.load()
    rng = np.random.default_rng(7)
    n = 0
    for k in range(4000):
        zr, zi = rng.normal() * 1.3, rng.normal() * 1.3
        if k % 5 == 0:
            zr *= 1e-4
        gb = glitch_bound(zr, zi, 1e-6)
        e = entry(lib, zr, zi, gb)
        # walk delta_c over neighbouring doubles around -Z_c +- sqrt(gb)
        g = math.sqrt(gb)
        for sr in (-1, 0, 1):
            for si in (-1, 0, 1):
                dr, di = -zr + sr * g, -zi + si * g
                for _ in range(3):
                    for cand in ((dr, di), (np.nextafter(dr, 0), di), (dr, np.nextafter(di, 0)),
                                 (np.nextafter(dr, math.copysign(math.inf, dr)), di)):
                        gl, _ = exact_tests(zr, zi, float(cand[0]), float(cand[1]), gb)
                        if gl:
                            n += 1
                            assert fires(lib, e, float(cand[0]), float(cand[1]))[0], (zr, zi, cand, gb, list(e))
                    dr, di = dr * (1 - 2e-16), di * (1 - 2e-16)
    assert n > 1000


def test_glitch_filter_degenerate_entries():
    lib = newman_b200.load()
    # gb == 0 (index 0, the escaped iterate, padding): the exact test can never fire and neither does the filter
    e = entry(lib, 0.0, 0.0, 0.0)
    assert list(e)[:4] == [0xffffffff, 0, 0xffffffff, 0]
    assert not fires(lib, e, 0.0, 0.0)[0] and not fires(lib, e, -1e-300, 1e-300)[0]
    # tiny reference iterate / denormal bound: every state is a candidate, scaled ones included
    for z, tol in ((1e-120, 1e-6), (3e-31, 1e-6), (1e-140, 1e-6)):
        e = entry(lib, z, -z, glitch_bound(z, -z, tol))
        assert list(e)[:4] == [0, 0xffffffff, 0, 0xffffffff]
        assert fires(lib, e, 1.0, 1.0)[0] and fires(lib, e, 1.5, -1.25, scaled=1)[0]
    # a regular entry never fires for a scaled state (|delta| < 2^-127 cannot reach -Z)
    e = entry(lib, 0.3, -0.2, glitch_bound(0.3, -0.2, 1e-6))
    assert fires(lib, e, -0.3, 0.2)[0]
    assert not fires(lib, e, -0.3, 0.2, scaled=1)[0]
    # glitch_tol so large that no component can be bounded away from zero: always a candidate (slow, still exact)
    e = entry(lib, 0.3, -0.2, glitch_bound(0.3, -0.2, 4.0))
    assert fires(lib, e, 5.0, 5.0)[0]


def test_glitch_filter_is_selective():
    """False alarms cost a replay, not correctness — but the filter must not fire on ordinary states."""
    lib = newman_b200.load()
    rng = np.random.default_rng(3)
    alarms = 0
    for _ in range(20000):
        zr, zi = rng.normal(), rng.normal()
        gb = glitch_bound(zr, zi, 1e-6)
        e = entry(lib, zr, zi, gb)
        dr, di = rng.normal(), rng.normal()   # |delta| ~ |Z|: the worst phase of a pixel's life
        gl, _ = exact_tests(zr, zi, dr, di, gb)
        alarms += fires(lib, e, dr, di)[0] and not gl
    assert alarms <= 2


def test_escape_filter_has_no_false_negatives():
    lib = newman_b200.load()
    rng = np.random.default_rng(11)
    hits = 0
    for k in range(40000):
        zmag = rng.uniform(0, 2.0) if k % 4 else 10.0 ** rng.uniform(0, 4)     # orbit entries, and the escaped tail
        th = rng.uniform(0, 2 * math.pi)
        zr, zi = zmag * math.cos(th), zmag * math.sin(th)
        e = entry(lib, zr, zi, 0.0 if zmag > 1024 else glitch_bound(zr, zi, 1e-6))
        r = 1024.0 * float(rng.choice([1 - 1e-3, 1 - 1e-9, 1 - 1e-15, 1.0, 1 + 1e-15, 1 + 1e-12, 1.001, 2.0, 1e3, 1e40]))
        ph = rng.uniform(0, 2 * math.pi) if k % 3 else rng.choice([0, 0.25, 0.5, 0.75, 1, 1.25, 1.5, 1.75]) * math.pi
        dr, di = r * math.cos(ph) - zr, r * math.sin(ph) - zi
        _, esc = exact_tests(zr, zi, dr, di, 0.0)
        if esc:
            hits += 1
            assert fires(lib, e, dr, di)[1], ("escape missed", zr, zi, dr, di, list(e))
    assert hits > 15000
    # scaled states escape only together with the reference (entry = the escaped iterate)
    e = entry(lib, 900.0, 800.0, 0.0)
    assert list(e)[4] == 0 and fires(lib, e, 1.0, 1.0, scaled=1)[1]
    e = entry(lib, 0.5, 0.5, glitch_bound(0.5, 0.5, 1e-6))
    assert not fires(lib, e, 1.9, 1.9, scaled=1)[1] and not fires(lib, e, 1.9, 1.9)[1]
    assert fires(lib, e, 800.0, 0.0)[1]


def test_scaled_step_uses_the_same_bounds():
    """z = fma(S, d, Z) with S = 2^e: the filter looks at d of PLAIN states only; for e == 0 the fma form equals
    the add the bound was derived for."""
    zr, zi, dr, di = 0.3, -0.2, -0.3 + 1e-5, 0.2 - 3e-5
    gb = glitch_bound(zr, zi, 1e-6)
    assert exact_tests(zr, zi, dr, di, gb) == exact_tests(zr, zi, dr, di, gb, S=1.0)
    a = exact_tests(zr, zi, dr, di, gb)
    r, i = fl(Fr(1) * Fr(dr) + Fr(zr)), fl(Fr(1) * Fr(di) + Fr(zi))
    assert (r, i) == (zr + dr, zi + di) and a[0]
