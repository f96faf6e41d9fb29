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


# ---- quiet segments (k3_filter.cuh: k3_seg_bound) -------------------------------------------------------------------
def hi_word(x):
    return int(np.float64(abs(x)).view(np.uint64) >> np.uint64(32))


def seg_bound(lib, z, gb, e_max, j0=0, steps=16):
    """the quiet bound of k3_fast's 16-iteration segments"""
    z = np.ascontiguousarray(z, np.float64)
    gb = np.ascontiguousarray(gb, np.float64)
    assert steps == 16
    fn = lib.nm_k3_seg_bound
    return fn(z.ctypes.data_as(C.POINTER(C.c_double)), gb.ctypes.data_as(C.POINTER(C.c_double)), j0, len(gb) - 1, e_max)


def admissible(T, dr, di):
    """The kernel's test at the start of a segment (plain state)."""
    return hi_word(dr) < T and hi_word(di) < T


def fast_step(zr, zi, dr, di, er, ei):
    """One iteration exactly as k3_fast's quiet block computes it (k3_fast.cuh: k3_block_quiet, plain state)."""
    wr, wi = 2.0 * zr + dr, 2.0 * zi + di               # DADD against the exact 2Z table
    t = fl(Fr(dr) * Fr(wr) + Fr(er))
    ndr = fl(Fr(-di) * Fr(wi) + Fr(t))
    t = fl(Fr(dr) * Fr(wi) + Fr(ei))
    ndi = fl(Fr(di) * Fr(wr) + Fr(t))
    return ndr, ndi


def random_orbit_piece(rng, n=18):
    """n table entries of arbitrary size (the bound may not rely on Z being a Mandelbrot orbit)."""
    style = rng.integers(0, 4)
    if style == 0:
        mag = 10.0 ** rng.uniform(-0.6, 0.3, n)
    elif style == 1:
        mag = 10.0 ** rng.uniform(-6, 0.3, n)
    else:
        mag = 10.0 ** rng.uniform(-0.5, 0.3, n)
        mag[rng.integers(1, n)] = 10.0 ** rng.uniform(-14, -1)   # an approach to zero
    th = rng.uniform(0, 2 * math.pi, n)
    return np.stack([mag * np.cos(th), mag * np.sin(th)], 1)


@pytest.mark.parametrize("steps", [16])
def test_quiet_bound_never_admits_a_glitching_state(steps):
    """Build a trajectory that DOES glitch at index k of the segment (Z_k is placed on -delta_k afterwards): the
    bound computed from that table must reject the trajectory's start state."""
    lib = newman_b200.load()
    rng = np.random.default_rng(20261018 + steps)
    hits = 0
    for trial in range(6000 if steps == 16 else 4000):
        z = random_orbit_piece(rng, steps + 2)
        tol = float(rng.choice([1e-6, 1e-6, 1e-6, 1e-3, 1e-9, 0.01]))
        k = int(rng.integers(1, steps + 1))
        e_max = 10.0 ** rng.uniform(-60, -3)
        ea = rng.uniform(0, 2 * math.pi)
        er, ei = e_max * math.cos(ea) * 0.999, e_max * math.sin(ea) * 0.999
        # the start state: anything from far below eps to the size of Z
        d0 = 10.0 ** rng.uniform(math.log10(e_max) - 3, 0.0)
        da = rng.uniform(0, 2 * math.pi) if trial % 3 else rng.choice([0, 0.25, 0.5, 0.75, 1.0, 1.5]) * math.pi
        dr0, di0 = d0 * math.cos(da), d0 * math.sin(da)
        dr, di = dr0, di0
        ok = True
        for i in range(k):
            dr, di = fast_step(z[i, 0], z[i, 1], dr, di, er, ei)
            if not (math.isfinite(dr) and math.isfinite(di)) or max(abs(dr), abs(di)) > 1e3:
                ok = False
                break
        if not ok or max(abs(dr), abs(di)) < 1e-90:
            continue
        # put Z_k inside the glitch disc around -delta_k: |Z_k + delta_k| = rad * sqrt(tol) * |Z_k|
        rad = float(rng.choice([0.0, 0.3, 0.9, 0.999, 0.999999]))
        ph = rng.uniform(0, 2 * math.pi)
        m = math.hypot(dr, di)
        z[k, 0] = -dr + rad * math.sqrt(tol) * m * math.cos(ph)
        z[k, 1] = -di + rad * math.sqrt(tol) * m * math.sin(ph)
        gb = np.array([glitch_bound(a, b, tol) for a, b in z])
        gb[0] = 0.0
        gl, _ = exact_tests(z[k, 0], z[k, 1], dr, di, gb[k])
        if not gl:
            continue
        hits += 1
        T = seg_bound(lib, z, gb, e_max, steps=steps)
        assert not admissible(T, dr0, di0), ("glitch inside a quiet segment", trial, k, T, dr0, di0, z.tolist(), tol, e_max)
    assert hits > (3000 if steps == 16 else 1500)


@pytest.mark.parametrize("steps", [16])
def test_quiet_bound_largest_admissible_states_do_not_glitch(steps):
    """The other direction: the largest states the bound admits (high words T - 1, all-ones low words, every sign
    pattern, eps of full size in the worst directions) are iterated through the segment with the kernel's arithmetic;
    the exact glitch test may fire nowhere."""
    lib = newman_b200.load()
    rng = np.random.default_rng(5 + steps)
    checked = 0
    for trial in range(1500):
        z = random_orbit_piece(rng, steps + 2)
        tol = float(rng.choice([1e-6, 1e-6, 1e-3, 1e-9]))
        gb = np.array([glitch_bound(a, b, tol) for a, b in z])
        gb[0] = 0.0
        e_max = 10.0 ** rng.uniform(-60, -2)
        T = seg_bound(lib, z, gb, e_max, steps=steps)
        if T <= 0:
            continue
        top = float(np.uint64((T - 1) << 32 | 0xffffffff).view(np.float64))
        for sr, si in ((1, 1), (1, -1), (-1, 1), (-1, -1), (1, 0), (0, -1)):
            dr, di = sr * top, si * top
            ea = rng.uniform(0, 2 * math.pi)
            er, ei = e_max * math.cos(ea) * 0.9999, e_max * math.sin(ea) * 0.9999
            for i in range(steps):
                dr, di = fast_step(z[i, 0], z[i, 1], dr, di, er, ei)
                if not (math.isfinite(dr) and math.isfinite(di)):
                    break
                gl, _ = exact_tests(z[i + 1, 0], z[i + 1, 1], dr, di, gb[i + 1])
                assert not gl, ("admitted state glitches", trial, i, T, z.tolist(), tol, e_max)
            checked += 1
    assert checked > (3000 if steps == 16 else 1000)


def test_quiet_bound_is_useful_and_degenerates_safely():
    lib = newman_b200.load()
    # an ordinary stretch of orbit (|Z| around 1, growth ~2 per step) admits deltas up to ~1e-6
    n = 40
    z = np.stack([np.full(n, 0.7), np.full(n, -0.6)], 1)
    gb = np.array([glitch_bound(a, b, 1e-6) for a, b in z]); gb[0] = 0.0
    T = seg_bound(lib, z, gb, 1e-52, j0=16)
    lim = float(np.uint64(T << 32).view(np.float64))
    assert 1e-7 < lim < 1e-3, lim
    # pixel offsets of the size of Z: nothing is quiet
    assert seg_bound(lib, z, gb, 1.0, j0=16) == 0
    # a segment that reaches beyond the table
    assert seg_bound(lib, z, gb, 1e-52, j0=32) == 0 and seg_bound(lib, z[:33], gb[:33], 1e-52, j0=16) > 0
    # an "always a candidate" entry (tiny reference iterate) inside the segment
    z2 = z.copy(); z2[20] = (1e-120, 1e-121)
    gb2 = np.array([glitch_bound(a, b, 1e-6) for a, b in z2]); gb2[0] = 0.0
    assert seg_bound(lib, z2, gb2, 1e-52, j0=16) == 0 and seg_bound(lib, z2, gb2, 1e-52, j0=0) > 0
    # gb == 0 entries (index 0, the escaped iterate) constrain nothing
    z3 = z.copy(); z3[32] = (1500.0, 200.0)
    gb3 = gb.copy(); gb3[32] = 0.0
    assert seg_bound(lib, z3, gb3, 1e-52, j0=16) >= T   # (the last index was the binding one: the bound relaxes)
    # NaN / inf offsets
    assert seg_bound(lib, z, gb, math.nan, j0=16) == 0 and seg_bound(lib, z, gb, math.inf, j0=16) == 0


@pytest.mark.parametrize("steps", [16])
def test_quiet_bound_on_a_real_frame(steps):
    """Every sample of a small cfg2 frame (1e-50, M = 54 512) is iterated through the whole reference orbit with the
    perturbation recurrence (complex128: the bound's 2^-40 slack dwarfs the difference to the kernel's fma order); a
    sample that the segment bound admits at a segment start may not satisfy the glitch test anywhere in that segment.
    This pins the table indexing (Z[0] = 0, Z[j] = X[j-1], the new delta pairs with Z[j+1]) and the use of the
    frame's largest pixel offset on real data; the frame does contain glitches, all of them in loud segments."""
    from newman_b200 import workloads
    lib = newman_b200.load()
    cfg = workloads.config("cfg2", scale=40)
    v = newman_b200.Mandelbrot(cfg["nr"], cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
    h = v.host_tables(cfg["nr"] // 2, cfg["nc"] // 2)
    xh = np.asarray(h["x_hi"]).reshape(-1, 2)
    J = len(xh)                                   # last valid table index (the escaped iterate included)
    z = np.zeros((J + 1, 2))
    z[1:] = xh
    gtol = 1e-6
    gb = (z[:, 0] ** 2 + z[:, 1] ** 2) * gtol
    gb[0] = 0.0
    if h["has_escape"]:
        gb[J] = 0.0                               # k_glitch_bounds: the escaped iterate never flags
    er, ei = np.asarray(h["eps_re"], float), np.asarray(h["eps_im"], float)
    e_max = math.hypot(np.abs(er).max(), np.abs(ei).max()) * (1 + 2.0 ** -40)
    nseg = J // steps
    T = np.array([seg_bound(lib, z, gb, e_max, j0=steps * s, steps=steps) for s in range(nseg)], dtype=np.int64)
    assert (T > 0).mean() > 0.99

    Z = z[:, 0] + 1j * z[:, 1]
    eps = (er[None, :] + 1j * ei[:, None]).reshape(-1)
    d = np.zeros_like(eps)
    alive = np.ones(eps.shape, bool)
    admitted = np.zeros(eps.shape, bool)
    glitches = loud_glitches = checked = 0
    with np.errstate(over="ignore", invalid="ignore"):
        for j in range(J):
            if j % steps == 0:
                s = j // steps
                if s < nseg and j + steps <= J:
                    hi = np.maximum(np.abs(d.real), np.abs(d.imag)).view(np.uint64) >> np.uint64(32)
                    admitted = alive & (hi.astype(np.int64) < T[s])
                    checked += int(admitted.sum())
                else:
                    admitted[:] = False
            d = d * (2.0 * Z[j] + d) + eps
            zz = Z[j + 1] + d
            m2 = zz.real ** 2 + zz.imag ** 2
            g = alive & (m2 < gb[j + 1])
            if g.any():
                glitches += int(g.sum())
                loud_glitches += int((g & ~admitted).sum())
                assert not (g & admitted).any(), ("glitch inside an admitted segment", j, np.nonzero(g & admitted)[0][:5])
            alive &= ~g & ~(m2 > 1048576.0) & np.isfinite(m2)
            if not alive.any():
                break
    assert glitches > 0 and glitches == loud_glitches
    assert checked > 1000000 * 16 // steps
