"""ctypes access to the two CPU oracles (TEST INFRASTRUCTURE):
  Oracle-R  oracle/_ref/libnewman_ref.so   — the reference's mandelbrot.cpp compiled unmodified
  Oracle-P  oracle/_build/liboracle_p.so    — the plain-C restatement (oracle/oracle_p.c)
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libnewman_ref.so")
P_SO = os.path.join(ROOT, "oracle", "_build", "liboracle_p.so")
ESC = np.dtype([("iterations", "<i4"), ("smoothing", "<f4")])


def build_oracles():
    """(Re)build what can be built here: Oracle-P always; Oracle-R only where /root/reference exists."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oraclep"], check=True)
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)


def have_ref():
    return os.path.exists(REF_SO)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


class OpTables(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("has_escape", C.c_int32), ("flags", C.c_int32),
                ("tol", C.c_double), ("glitch_tol", C.c_double),
                ("x_hi", C.c_void_p), ("x_lo", C.c_void_p), ("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p),
                ("a_exp", C.c_void_p), ("b_exp", C.c_void_p), ("c_exp", C.c_void_p),
                ("eps_re_exp", C.c_void_p), ("eps_im_exp", C.c_void_p)]


class OpStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("executed_iters", "series_evals", "skipped_pixels", "glitched", "rebased")]


_p = None


def oraclep():
    global _p
    if _p is None:
        if not os.path.exists(P_SO):
            build_oracles()
        _p = C.CDLL(P_SO)
        _p.oraclep_render_deep.restype = C.c_int64
        _p.oraclep_render_deep.argtypes = [C.POINTER(OpTables), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.POINTER(OpStats)]
        _p.oraclep_render_hw.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.POINTER(OpStats)]
        _p.oraclep_series_L.argtypes = [C.POINTER(OpTables), C.c_double, C.c_double, C.POINTER(C.c_double),
                                        C.POINTER(C.c_double)]
        _p.oraclep_pick_reference.restype = C.c_int64
        _p.oraclep_pick_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        _p.oraclep_resolve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p]
        _p.oraclep_refine_dd.restype = C.c_int64
        _p.oraclep_refine_dd.argtypes = [C.POINTER(OpTables), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(OpStats)]
        _p.oraclep_smoothing.restype = C.c_float
        _p.oraclep_smoothing.argtypes = [C.c_double]
        _p.oraclep_trunc_add3.restype = C.c_double
        _p.oraclep_trunc_add3.argtypes = [C.c_double, C.c_double, C.c_double]
    return _p


def vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Tables:
    """Host-side deep tables (numpy, float64, interleaved re/im) as both oracles and the device want them."""

    def __init__(self, x_hi, x_lo, a, b, c, N, tol, glitch_tol=1e-6, exps=None, eps_exps=None):
        """exps = (a_exp, b_exp, c_exp): floatexp series, a/b/c are then mantissas (0.5 <= |m| < 1).
        eps_exps = (eps_re_exp, eps_im_exp): floatexp eps (the eps arrays handed to the render call
        are then mantissas) => scaled delta states."""
        self.x_hi, self.x_lo, self.a, self.b, self.c = [np.ascontiguousarray(v, dtype=np.float64) for v in
                                                        (x_hi, x_lo, a, b, c)]
        self.M = len(self.a) // 2
        self.has_escape = 1 if len(self.x_hi) // 2 == self.M + 1 else 0
        self.N, self.tol, self.glitch_tol = N, tol, glitch_tol
        self.exps = None if exps is None else [np.ascontiguousarray(v, dtype=np.int32) for v in exps]
        self.eps_exps = None if eps_exps is None else [np.ascontiguousarray(v, dtype=np.int32) for v in eps_exps]
        self.orbit_truncated = False   # True: K3 against the truncated orbit (the exact mode's probe rendering)

    def op(self):
        ex = [None] * 3 if self.exps is None else [v.ctypes.data for v in self.exps]
        ee = [None] * 2 if self.eps_exps is None else [v.ctypes.data for v in self.eps_exps]
        return OpTables(M=self.M, N=self.N, has_escape=self.has_escape, flags=1 if self.orbit_truncated else 0, tol=self.tol,
                        glitch_tol=self.glitch_tol, x_hi=self.x_hi.ctypes.data, x_lo=self.x_lo.ctypes.data,
                        a=self.a.ctypes.data, b=self.b.ctypes.data, c=self.c.ctypes.data,
                        a_exp=ex[0], b_exp=ex[1], c_exp=ex[2], eps_re_exp=ee[0], eps_im_exp=ee[1])

    def floatexp(self, eps_re=None, eps_im=None):
        """The same tables in floatexp form (numpy frexp == mpf_get_d_2exp for doubles): for tests that
        run both modes on one view. With eps arrays: also returns their (mantissa, exponent) split."""
        ms, es = zip(*[np.frexp(v) for v in (self.a, self.b, self.c)])
        t = Tables(self.x_hi, self.x_lo, ms[0], ms[1], ms[2], self.N, self.tol, self.glitch_tol, exps=es)
        if eps_re is None:
            return t
        (mr, er), (mi, ei) = np.frexp(eps_re), np.frexp(eps_im)
        t.eps_exps = [np.ascontiguousarray(er, dtype=np.int32), np.ascontiguousarray(ei, dtype=np.int32)]
        return t, np.ascontiguousarray(mr), np.ascontiguousarray(mi)


def p_render_deep(t, eps_re, eps_im, cardioid_mode=0, mask=None, pix_list=None, mode=0, out=None):
    nr, nc = len(eps_im), len(eps_re)
    if out is None:
        out = np.zeros((nr, nc), dtype=ESC)
    W = nr * nc if pix_list is None else len(pix_list)
    rq_pix = np.zeros(max(W, 1), dtype=np.int32)
    rq_it = np.zeros(max(W, 1), dtype=np.int32)
    st = OpStats()
    ot = t.op()
    n = oraclep().oraclep_render_deep(C.byref(ot), vp(eps_re), nc, vp(eps_im), nr, cardioid_mode, vp(mask),
                                      vp(pix_list), 0 if pix_list is None else len(pix_list), mode, vp(out),
                                      vp(rq_pix), vp(rq_it), C.byref(st))
    return out, rq_pix[:n].copy(), rq_it[:n].copy(), {k: getattr(st, k) for k, _ in st._fields_}


def p_refine_dd(t, eps_re, eps_im, pix_list, eps_lo=None, out=None):
    """oraclep_refine_dd: phase 3 of the listed samples in double-double arithmetic (the exact mode's second pass)."""
    nr, nc = len(eps_im), len(eps_re)
    if out is None:
        out = np.zeros((nr, nc), dtype=ESC)
    pix = np.ascontiguousarray(pix_list, dtype=np.int32)
    lo = [None, None] if eps_lo is None else [np.ascontiguousarray(x, dtype=np.float64) for x in eps_lo]
    st = OpStats()
    ot = t.op()
    rc = oraclep().oraclep_refine_dd(C.byref(ot), vp(eps_re), vp(lo[0]), nc, vp(eps_im), vp(lo[1]), nr, vp(pix), len(pix), vp(out),
                                     C.byref(st))
    assert rc == 0, "scaled frames are not refined"
    return out, {k: getattr(st, k) for k, _ in st._fields_}


def p_render_hw(c_re, c_im, N, mask=None):
    out = np.zeros((len(c_im), len(c_re)), dtype=ESC)
    st = OpStats()
    oraclep().oraclep_render_hw(vp(c_re), len(c_re), vp(c_im), len(c_im), N, vp(mask), vp(out), C.byref(st))
    return out, {k: getattr(st, k) for k, _ in st._fields_}


def p_resolve(grid, pal, N, sc, smooth):
    nr, nc = grid.shape
    out = np.zeros((nr // sc, nc // sc, 3), dtype=np.uint8)
    oraclep().oraclep_resolve(vp(grid), nr, nc, vp(pal), len(pal.reshape(-1)) // 3, N, sc, int(smooth), vp(out))
    return out


class RefView:
    """Oracle-R view (SURVEY.md §8d construction order)."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not have_ref():
                build_oracles()
            L = C.CDLL(REF_SO)
            L.ref_create.restype = C.c_void_p
            L.ref_create.argtypes = [C.c_int, C.c_int]
            for f in ("ref_precompute", "ref_render_all", "ref_compute_rows", "ref_compute_pixels"):
                getattr(L, f).restype = C.c_double
            L.ref_destroy.argtypes = [C.c_void_p]
            L.ref_set_view.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_double]
            L.ref_precompute.argtypes = [C.c_void_p]
            L.ref_precompute_at.argtypes = [C.c_void_p, C.c_int, C.c_int]
            L.ref_render_all.argtypes = [C.c_void_p, C.c_void_p]
            L.ref_compute_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            L.ref_compute_pixels.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            for f in ("ref_orbit_len", "ref_precision_bits", "ref_use_hardware", "ref_rows", "ref_cols"):
                getattr(L, f).argtypes = [C.c_void_p]
            L.ref_dump_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
            L.ref_dump_orbit_dd.argtypes = [C.c_void_p, C.c_void_p]
            L.ref_orbit_escape.argtypes = [C.c_void_p, C.c_void_p]
            L.ref_dump_eps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.ref_dump_table_2exp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
            L.ref_dump_eps_2exp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.ref_dump_coords.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.ref_in_cardioid.argtypes = [C.c_void_p, C.c_int, C.c_int]
            L.ref_scale.argtypes = [C.c_void_p, C.c_int, C.c_int]
            L.ref_read_grid.argtypes = [C.c_void_p, C.c_void_p]
            L.ref_zoom.argtypes = [C.c_void_p, C.c_float]
            L.ref_translate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
            L.ref_zoom_at.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int]
            L.ref_load_legacy.argtypes = [C.c_void_p, C.c_char_p]
            L.ref_view_string.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, nr, nc, N=256, sz=None, center=None, tol=1e-10):
        L = self.lib()
        self.h = C.c_void_p(L.ref_create(nr, nc))
        self.nr, self.nc, self.N, self.tol = nr, nc, N, tol
        e = lambda s: None if s is None else s.encode()
        L.ref_set_view(self.h, N, e(sz and sz[0]), e(sz and sz[1]), e(center and center[0]), e(center and center[1]), tol)

    def __del__(self):
        try:
            self.lib().ref_destroy(self.h)
        except Exception:
            pass

    def use_hardware(self):
        return bool(self.lib().ref_use_hardware(self.h))

    def precision_bits(self):
        return self.lib().ref_precision_bits(self.h)

    def precompute(self):
        return self.lib().ref_precompute(self.h)

    def precompute_at(self, r, c):
        self.lib().ref_precompute_at(self.h, r, c)

    def orbit_len(self):
        return self.lib().ref_orbit_len(self.h)

    def render_all(self):
        out = np.zeros((self.nr, self.nc), dtype=ESC)
        secs = self.lib().ref_render_all(self.h, vp(out))
        return out, secs

    def compute_rows(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        out = np.zeros((len(rows), self.nc), dtype=ESC)
        secs = self.lib().ref_compute_rows(self.h, vp(rows), len(rows), vp(out))
        return out, secs

    def compute_pixels(self, pix):
        pix = np.ascontiguousarray(pix, dtype=np.int32)
        out = np.zeros(len(pix), dtype=ESC)
        secs = self.lib().ref_compute_pixels(self.h, vp(pix), len(pix), vp(out))
        return out, secs

    def coords(self):
        cre = np.zeros(self.nc); cim = np.zeros(self.nr)
        self.lib().ref_dump_coords(self.h, vp(cre), vp(cim))
        return cre, cim

    def cardioid_mask(self):
        m = np.zeros((self.nr, self.nc), dtype=np.uint8)
        for r in range(self.nr):
            for c in range(self.nc):
                m[r, c] = self.lib().ref_in_cardioid(self.h, r, c)
        return m

    def in_cardioid(self, r, c):
        return bool(self.lib().ref_in_cardioid(self.h, r, c))

    def tables(self, glitch_tol=1e-6):
        L = self.lib()
        M = self.orbit_len()
        dd = np.zeros(4 * M)
        L.ref_dump_orbit_dd(self.h, vp(dd))
        dd = dd.reshape(M, 4)
        esc = np.zeros(2)
        has = L.ref_orbit_escape(self.h, vp(esc))
        x_hi = np.ascontiguousarray(dd[:, [0, 2]]).reshape(-1)
        x_lo = np.ascontiguousarray(dd[:, [1, 3]]).reshape(-1)
        if has:
            x_hi = np.concatenate([x_hi, esc])
        tabs = []
        for which in (1, 2, 3):
            t = np.zeros(2 * M)
            L.ref_dump_table(self.h, which, vp(t))
            tabs.append(t)
        return Tables(x_hi, x_lo, tabs[0], tabs[1], tabs[2], self.N, self.tol, glitch_tol)

    def tables_fe(self, glitch_tol=1e-6, scaled=False):
        """The tables in floatexp form (mantissa + exponent, mpf_get_d_2exp) straight from the compiled reference's mpf
        values — for views beyond the depth where its per-pixel code works (bench.py's CPU port leg). scaled: also
        returns the eps arrays as (mantissa, exponent) pairs. -> Tables[, (eps_re_m, eps_im_m)]"""
        L = self.lib()
        t = self.tables(glitch_tol)
        M = t.M
        ms, es = [], []
        for which in (1, 2, 3):
            m = np.zeros(2 * M); e = np.zeros(2 * M, dtype=np.int32)
            L.ref_dump_table_2exp(self.h, which, vp(m), vp(e))
            ms.append(m); es.append(e)
        tf = Tables(t.x_hi, t.x_lo, ms[0], ms[1], ms[2], self.N, self.tol, glitch_tol, exps=es)
        if not scaled:
            return tf
        mre = np.zeros(self.nc); mim = np.zeros(self.nr)
        ere = np.zeros(self.nc, dtype=np.int32); eim = np.zeros(self.nr, dtype=np.int32)
        L.ref_dump_eps_2exp(self.h, vp(mre), vp(ere), vp(mim), vp(eim))
        tf.eps_exps = [ere, eim]
        return tf, (mre, mim)

    def eps(self):
        er = np.zeros(self.nc); ei = np.zeros(self.nr)
        self.lib().ref_dump_eps(self.h, vp(er), vp(ei))
        return er, ei

    def view_strings(self):
        out = []
        for w in range(4):
            buf = C.create_string_buffer(8192)
            self.lib().ref_view_string(self.h, w, buf, 8192)
            out.append(buf.value.decode())
        return out


# The KAT views of SURVEY.md App. C (inputs only; results are recomputed and compared, and the
# digests this repo pins live in tests/golden/).
KATS = {
    "KAT-1c": dict(nr=48, nc=64, N=256),
    "KAT-1b": dict(nr=600, nc=800, N=256),
    "KAT-1": dict(nr=768, nc=1024, N=1024),
    "KAT-D30": dict(nr=96, nc=128, N=2000, sz=("7.8125e-33", "7.8125e-33"), center=("0", "1")),
    "KAT-D60": dict(nr=96, nc=128, N=2000, sz=("7.8125e-63", "7.8125e-63"), center=("0", "1")),
    "KAT-D90": dict(nr=96, nc=128, N=2000, sz=("7.8125e-93", "7.8125e-93"), center=("0", "1")),
    "KAT-B": dict(nr=96, nc=128, N=2000, sz=("7.8125e-33", "7.8125e-33"), center=("0", "1"), tol=1e9),
    "KAT-T3": dict(nr=96, nc=128, N=2000, sz=("7.8125e-33", "7.8125e-33"), center=("0", "1"), tol=1e-3),
    "KAT-S": dict(nr=30, nc=40, N=20000, sz=("1e-18", "1e-18"),
                  center=("-0.743643887037158704752191506114774", "0.131825904205311970493132056385139")),
}


class OracleDevice:
    """Oracle-P behind the newman_b200.Device frame interface (frame_deep / launch / requeue / stats /
    read_rows): lets the round orchestration (newman_b200.pipeline.render_rounds) and the multi-rank
    logic be exercised on CPU, and gives the GPU tests a whole-frame expectation. TEST ONLY."""

    def __init__(self):
        self.out = None

    def frame_deep(self, tables, eps_re, eps_im, cardioid_mode=0, mask=None, pix_list=None, mode=0):
        keep = tables._keep
        if isinstance(keep[0], dict):     # pipeline.TableSet.tables(): (arrays, row-restricted eps_im exponents)
            arr, eps_im_e = keep
            exps = [np.asarray(arr[k]) for k in ("a_e", "b_e", "c_e")] if "a_e" in arr else None
            eps_exps = None
            if "eps_re_e" in arr:
                eps_exps = [np.asarray(arr["eps_re_e"]), np.asarray(arr["eps_im_e"] if eps_im_e is None else eps_im_e)]
        else:                              # Device.make_tables(): (x_hi, x_lo, a, b, c, exps, eps_exps)
            arr = dict(zip(("x_hi", "x_lo", "a", "b", "c"), keep[:5]))
            exps = None if keep[5][0] is None else [np.asarray(v) for v in keep[5]]
            eps_exps = None if keep[6][0] is None else [np.asarray(v) for v in keep[6]]
        self._t = Tables(np.asarray(arr["x_hi"]), np.asarray(arr["x_lo"]), np.asarray(arr["a"]), np.asarray(arr["b"]),
                         np.asarray(arr["c"]), tables.N, tables.tol, tables.glitch_tol, exps=exps, eps_exps=eps_exps)
        self._args = (np.ascontiguousarray(eps_re), np.ascontiguousarray(eps_im), cardioid_mode, mask,
                      None if pix_list is None else np.ascontiguousarray(pix_list, dtype=np.int32), mode)
        self.nr, self.nc = len(eps_im), len(eps_re)
        if pix_list is None or self.out is None:
            self.out = np.zeros((self.nr, self.nc), dtype=ESC)

    def launch(self):
        er, ei, cm, mask, pl, mode = self._args
        _, self._rq_pix, self._rq_it, self._st = p_render_deep(self._t, er, ei, cm, mask, pl, mode, out=self.out)

    def stats(self):
        d = dict(self._st)
        d.update(pixels=self.nr * self.nc, fixups=0, kernel_launches=0, sweeps=0, checked_steps=0, ms_k1=0.0, ms_k2=0.0, ms_k3=0.0, ms_k4=0.0)
        return d

    def requeue(self):
        return self._rq_pix, self._rq_it

    def read_rows(self, r0=0, r1=None, out=None):
        r1 = self.nr if r1 is None else r1
        return self.out[r0:r1].copy()
