"""Python mirror of the reference's `class Mandelbrot` (reference mandelbrot.h:22-53) over the
view-level C-ABI (nmv_*). Same member names and argument meaning as the C++ drop-in, so the parity
tests read like code written against the reference. Plumbing only."""
import ctypes as C

import numpy as np

from . import _lib as L


class FrameInfo(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("hardware", "precision_bits", "orbit_len", "probe_row", "probe_col",
                                          "references")] +
                [(n, C.c_uint64) for n in ("executed_iters", "series_evals", "skipped_pixels", "glitched", "rebased",
                                           "fixups", "kernel_launches", "ambiguous")] +
                [(n, C.c_double) for n in ("host_precompute_s", "device_ms", "frame_s")] +
                [(n, C.c_uint64) for n in ("probe_iters", "probe_exact")] +
                [(n, C.c_int32) for n in ("probe_consistent", "cancelled")] +
                [("refined", C.c_uint64), ("refine_ms", C.c_double)])

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


VIEW_API = {
    "nmv_create": (C.c_void_p, [C.c_int, C.c_int]),
    "nmv_destroy": (None, [C.c_void_p]),
    "nmv_last_error": (C.c_char_p, [C.c_void_p]),
    "nmv_set_view": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_double]),
    "nmv_set_options": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]),
    "nmv_set_floatexp": (C.c_int, [C.c_void_p, C.c_int]),
    "nmv_set_probe_search": (C.c_int, [C.c_void_p, C.c_int]),
    "nmv_find_probe": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int)]),
    "nmv_get_n": (C.c_int, [C.c_void_p]),
    "nmv_rows": (C.c_int, [C.c_void_p]),
    "nmv_cols": (C.c_int, [C.c_void_p]),
    "nmv_use_hardware": (C.c_int, [C.c_void_p]),
    "nmv_precision_bits": (C.c_int, [C.c_void_p]),
    "nmv_precompute": (C.c_int, [C.c_void_p]),
    "nmv_compute_row": (C.c_int, [C.c_void_p, C.c_int]),
    "nmv_cancel": (C.c_int, [C.c_void_p]),
    "nmv_set_exact": (C.c_int, [C.c_void_p, C.c_int]),
    "nmv_render": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nmv_read_grid": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nmv_write_grid": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nmv_at_sc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(L.Escape)]),
    "nmv_scale": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "nmv_zoom": (C.c_int, [C.c_void_p, C.c_float]),
    "nmv_translate": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "nmv_zoom_at": (C.c_int, [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int]),
    "nmv_load_legacy": (C.c_int, [C.c_void_p, C.c_char_p]),
    "nmv_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "nmv_view_string": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int]),
    "nmv_frame_info_get": (C.c_int, [C.c_void_p, C.POINTER(FrameInfo)]),
    "nmv_resolve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nmv_host_tables": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int)]),
    "nmv_host_table": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "nmv_host_table_exp": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "nmv_host_coords": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nmv_host_cardioid": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nmv_host_in_cardioid": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "nmv_host_selfcheck": (C.c_int, []),
    "nmv_set_devices": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int]),
    # multi-GPU render groups (include/newman_b200.h "multi-GPU"; newman_b200/multigpu.py: RenderGroup)
    "nmm_band_layout": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nmm_unique_id": (C.c_int, [C.c_void_p]),
    "nmm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "nmm_destroy": (None, [C.c_void_p]),
    "nmm_last_error": (C.c_char_p, [C.c_void_p]),
    "nmm_ctx": (C.c_void_p, [C.c_void_p]),
    "nmm_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(FrameInfo)]),
    "nmm_resolve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "nmm_exchange_ms": (C.c_double, [C.c_void_p]),
}

_bound = False


def _lib():
    global _bound
    lib = L.load()
    if not _bound:
        for name, (res, args) in VIEW_API.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _bound = True
    return lib


class Mandelbrot:
    """`Mandelbrot(nr, nc)`; fields N / error_tolerance / center / sz are set through set_view()."""

    def __init__(self, nr, nc, N=None, sz=None, center=None, tol=1e-10, glitch_tol=-1.0, max_secondary=-1, device=-1,
                 host_threads=-1):
        self.lib = _lib()
        self.h = C.c_void_p(self.lib.nmv_create(nr, nc))
        if not self.h:
            raise L.NmError(L.NM_EINVAL, self.lib.nmv_last_error(None).decode())
        self.N = 256
        if N is not None or sz is not None or center is not None:
            self.set_view(N if N is not None else 256, sz, center, tol)
        self.lib.nmv_set_options(self.h, glitch_tol, max_secondary, device, host_threads)

    def close(self):
        if getattr(self, "h", None):
            self.lib.nmv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise L.NmError(rc, self.lib.nmv_last_error(self.h).decode())
        return rc

    def set_view(self, N, sz=None, center=None, tol=1e-10):
        e = lambda s: None if s is None else str(s).encode()
        self.N = N
        self._ck(self.lib.nmv_set_view(self.h, N, e(sz and sz[0]), e(sz and sz[1]), e(center and center[0]),
                                       e(center and center[1]), tol))

    def set_options(self, glitch_tol=-1.0, max_secondary=-1, device=-1, host_threads=-1):
        self.lib.nmv_set_options(self.h, glitch_tol, max_secondary, device, host_threads)

    def set_exact(self, on=True):
        """Mandelbrot::exact: repeat the samples FP64 perturbation cannot resolve in double-double arithmetic."""
        self._ck(self.lib.nmv_set_exact(self.h, 1 if on else 0))

    def cancel(self):
        """Abandon the frame another thread is rendering through this view (nmv_cancel)."""
        self.lib.nmv_cancel(self.h)

    def set_devices(self, devices, band_rows=0):
        """Mandelbrot::devices: split every frame of this view over these GPUs (threads + NCCL inside the library)."""
        arr = (C.c_int * len(devices))(*devices)
        self._ck(self.lib.nmv_set_devices(self.h, arr, len(devices), int(band_rows)))

    def set_floatexp(self, force):
        """0 automatic; 1 floatexp series; 2 also floatexp eps + scaled deltas (even where doubles suffice)."""
        self._ck(self.lib.nmv_set_floatexp(self.h, int(force)))

    def set_probe_search(self, mode):
        """findProbe used by precompute(): 1 GPU-assisted (default), 0 the reference's exhaustive search."""
        self._ck(self.lib.nmv_set_probe_search(self.h, int(mode)))

    def find_probe(self, mode=1):
        """-> (row, col, exact orbit length, candidates measured in arbitrary precision)"""
        r, c, l, n = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.nmv_find_probe(self.h, int(mode), C.byref(r), C.byref(c), C.byref(l), C.byref(n)))
        return r.value, c.value, l.value, n.value

    def rows(self):
        return self.lib.nmv_rows(self.h)

    def cols(self):
        return self.lib.nmv_cols(self.h)

    def useHardware(self):
        return bool(self.lib.nmv_use_hardware(self.h))

    def precision_bits(self):
        return self.lib.nmv_precision_bits(self.h)

    def precompute(self):
        self._ck(self.lib.nmv_precompute(self.h))

    def computeRow(self, r):
        self._ck(self.lib.nmv_compute_row(self.h, r))

    def render(self, out=None):
        if out is None:
            out = np.zeros((self.rows(), self.cols()), dtype=L.ESCAPE_DTYPE)
        self._ck(self.lib.nmv_render(self.h, L.ptr(out)))
        return out

    def grid(self):
        out = np.zeros((self.rows(), self.cols()), dtype=L.ESCAPE_DTYPE)
        self._ck(self.lib.nmv_read_grid(self.h, L.ptr(out)))
        return out

    def set_grid(self, g):
        g = np.ascontiguousarray(g, dtype=L.ESCAPE_DTYPE)
        assert g.shape == (self.rows(), self.cols())
        self._ck(self.lib.nmv_write_grid(self.h, L.ptr(g)))

    def at(self, r, c, sc=None):
        if sc is None:
            return self.grid()[r, c]
        e = L.Escape()
        self._ck(self.lib.nmv_at_sc(self.h, r, c, sc, C.byref(e)))
        return e.iterations, e.smoothing

    def scaleUp(self, sc):
        self._ck(self.lib.nmv_scale(self.h, sc, 1))

    def scaleDown(self, sc):
        self._ck(self.lib.nmv_scale(self.h, sc, 0))

    def zoom(self, scale):
        self._ck(self.lib.nmv_zoom(self.h, scale))

    def translate(self, dr, dc, sc=1):
        self._ck(self.lib.nmv_translate(self.h, dr, dc, sc))

    def zoomAt(self, scale, r, c, sc=1):
        self._ck(self.lib.nmv_zoom_at(self.h, scale, r, c, sc))

    def loadLegacy(self, fn):
        self._ck(self.lib.nmv_load_legacy(self.h, fn.encode()))
        self.N = self.frame_N()

    def frame_N(self):
        """The C++ object's public field N."""
        return self.lib.nmv_get_n(self.h)

    def save(self, fn):
        self._ck(self.lib.nmv_save(self.h, fn.encode()))

    def view_strings(self):
        out = []
        for w in range(4):
            buf = C.create_string_buffer(8192)
            self._ck(self.lib.nmv_view_string(self.h, w, buf, 8192))
            out.append(buf.value.decode())
        return out

    def frame_info(self):
        fi = FrameInfo()
        self._ck(self.lib.nmv_frame_info_get(self.h, C.byref(fi)))
        return fi.asdict()

    def resolve(self, pal_rgb, sc=1, smooth=True):
        pal_rgb = np.ascontiguousarray(pal_rgb, dtype=np.uint8)
        out = np.zeros((self.rows() // sc, self.cols() // sc, 3), dtype=np.uint8)
        self._ck(self.lib.nmv_resolve(self.h, L.ptr(pal_rgb), pal_rgb.size // 3, sc, int(bool(smooth)), L.ptr(out)))
        return out

    # -- host-only pieces (CPU tests) -----------------------------------------------------------
    def host_tables(self, row=-1, col=-1):
        he, pr, pc = C.c_int(), C.c_int(), C.c_int()
        M = self._ck(self.lib.nmv_host_tables(self.h, row, col, C.byref(he), C.byref(pr), C.byref(pc)))
        sizes = [2 * (M + he.value), 2 * M, 2 * M, 2 * M, 2 * M, self.cols(), self.rows()]
        arrs = []
        for which, n in enumerate(sizes):
            a = np.zeros(n)
            self._ck(self.lib.nmv_host_table(self.h, which, L.ptr(a)))
            arrs.append(a)
        keys = ("x_hi", "x_lo", "a", "b", "c", "eps_re", "eps_im")
        d = dict(zip(keys, arrs))
        # floatexp form (mantissa in [0.5, 1) and binary exponent, as mpf_get_d_2exp returns them)
        for which, key, n in ((7, "a", 2 * M), (8, "b", 2 * M), (9, "c", 2 * M), (10, "eps_re", self.cols()),
                              (11, "eps_im", self.rows())):
            m = np.zeros(n)
            e = np.zeros(n, dtype=np.int32)
            self._ck(self.lib.nmv_host_table(self.h, which, L.ptr(m)))
            self._ck(self.lib.nmv_host_table_exp(self.h, which - 5 if which < 10 else which, L.ptr(e)))
            d[key + "_m"], d[key + "_e"] = m, e
        d.update(M=M, has_escape=he.value, probe=(pr.value, pc.value))
        d["finite"] = bool(np.isfinite(d["a"]).all() and np.isfinite(d["b"]).all() and np.isfinite(d["c"]).all())
        return d

    def host_coords(self):
        cre, cim = np.zeros(self.cols()), np.zeros(self.rows())
        self._ck(self.lib.nmv_host_coords(self.h, L.ptr(cre), L.ptr(cim)))
        return cre, cim

    def host_cardioid(self):
        mask = np.zeros((self.rows(), self.cols()), dtype=np.uint8)
        mode = self._ck(self.lib.nmv_host_cardioid(self.h, L.ptr(mask)))
        return mode, mask

    def host_in_cardioid(self, r, c):
        return bool(self._ck(self.lib.nmv_host_in_cardioid(self.h, r, c)))
