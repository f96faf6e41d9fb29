"""Thin object wrapper over the device-level C-ABI (nm_ctx). Plumbing only: every call forwards to
libnewman_b200.so; there is no arithmetic here."""
import ctypes as C

import numpy as np

from . import _lib as L


class Device:
    """One nm_ctx = one GPU + one stream. Mirrors include/newman_b200.h one to one."""

    def __init__(self, device=0, lib=None):
        self.lib = lib if lib is not None else L.load()
        h = C.c_void_p()
        rc = self.lib.nm_create(device, C.byref(h))
        if rc != L.NM_OK:
            raise L.NmError(rc, self.lib.nm_last_error(None).decode())
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.nm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != L.NM_OK:
            raise L.NmError(rc, self.lib.nm_last_error(self.h).decode())

    # -- plumbing -----------------------------------------------------------------------------
    def set_stream(self, cuda_stream_handle):
        self._ck(self.lib.nm_set_stream(self.h, C.c_void_p(cuda_stream_handle or 0)))

    def set_option(self, key, value):
        self._ck(self.lib.nm_set_option(self.h, int(key), int(value)))

    def sync(self):
        self._ck(self.lib.nm_sync(self.h))

    def cancel(self):
        self._ck(self.lib.nm_cancel(self.h))

    def info(self):
        sm, khz, mem = C.c_int(), C.c_int(), C.c_size_t()
        name = C.create_string_buffer(128)
        self._ck(self.lib.nm_device_info(self.h, C.byref(sm), C.byref(khz), C.byref(mem), name, 128))
        return {"sm_count": sm.value, "sm_clock_khz": khz.value, "hbm_bytes": mem.value, "name": name.value.decode()}

    def fp64_peak(self, kind=0, iters=1 << 16):
        ips, ms = C.c_double(), C.c_double()
        self._ck(self.lib.nm_fp64_peak(self.h, kind, iters, C.byref(ips), C.byref(ms)))
        return ips.value, ms.value

    # -- frames -------------------------------------------------------------------------------
    def frame_hw(self, c_re, c_im, N):
        self._keep = [c_re, c_im]
        self.nr, self.nc = len(c_im), len(c_re)
        self._ck(self.lib.nm_frame_hw(self.h, L.ptr(c_re), len(c_re), L.ptr(c_im), len(c_im), N))

    @staticmethod
    def make_tables(x_hi, x_lo, a, b, c, N, tol, glitch_tol=1e-6, has_escape=None, exps=None, eps_exps=None, eps_lo=None,
                    orbit_truncated=False):
        """exps = (a_exp, b_exp, c_exp) int32: floatexp series (a/b/c are then mantissas);
        eps_exps = (eps_re_exp, eps_im_exp) int32: floatexp eps => scaled delta states (the eps arrays
        given to frame_deep/render_deep are then mantissas)."""
        M = len(a) // 2 if getattr(a, "ndim", 1) == 1 else a.shape[0]
        n_x = len(x_hi) // 2 if getattr(x_hi, "ndim", 1) == 1 else x_hi.shape[0]
        if has_escape is None:
            has_escape = 1 if n_x == M + 1 else 0
        pv = lambda x: None if x is None else L.ptr(x).value
        ex = exps or (None, None, None)
        ee = eps_exps or (None, None)
        el = [None if eps_lo is None else np.ascontiguousarray(x, dtype=np.float64) for x in (eps_lo or (None, None))]
        t = L.DeepTables(M=M, N=N, has_escape=has_escape, flags=L.TABLES_ORBIT_TRUNCATED if orbit_truncated else 0, tol=tol,
                         glitch_tol=glitch_tol, eps_re_lo=pv(el[0]), eps_im_lo=pv(el[1]),
                         x_hi=L.ptr(x_hi).value, x_lo=L.ptr(x_lo).value, a=L.ptr(a).value, b=L.ptr(b).value,
                         c=L.ptr(c).value, a_exp=pv(ex[0]), b_exp=pv(ex[1]), c_exp=pv(ex[2]),
                         eps_re_exp=pv(ee[0]), eps_im_exp=pv(ee[1]))
        t._keep = (x_hi, x_lo, a, b, c, ex, ee, el)
        return t

    def frame_deep(self, tables, eps_re, eps_im, cardioid_mode=L.CARDIOID_NONE, mask=None, pix_list=None,
                   mode=L.MODE_REQUEUE):
        self._keep = [tables, eps_re, eps_im, mask, pix_list]
        self.nr, self.nc = len(eps_im), len(eps_re)
        n_list = 0 if pix_list is None else len(pix_list)
        self._ck(self.lib.nm_frame_deep(self.h, C.byref(tables), L.ptr(eps_re), len(eps_re), L.ptr(eps_im), len(eps_im),
                                        cardioid_mode, L.ptr(mask), L.ptr(pix_list), n_list, mode))

    def launch(self):
        self._ck(self.lib.nm_launch(self.h))

    def ambiguous(self):
        n = self.lib.nm_frame_ambiguous(self.h, None, 0)
        if n < 0:
            self._ck(int(n))
        out = np.zeros(n, dtype=np.int32)
        if n:
            self.lib.nm_frame_ambiguous(self.h, L.ptr(out), n)
        return out

    def requeue(self):
        n = self.lib.nm_frame_requeue(self.h, None, None, 0)
        if n < 0:
            self._ck(int(n))
        pix = np.zeros(n, dtype=np.int32)
        it = np.zeros(n, dtype=np.int32)
        if n:
            self.lib.nm_frame_requeue(self.h, L.ptr(pix), L.ptr(it), n)
        return pix, it

    def poke(self, pix, iterations, smoothing):
        self._ck(self.lib.nm_poke(self.h, int(pix), L.Escape(int(iterations), float(smoothing))))

    def read_rows(self, r0=0, r1=None, out=None):
        r1 = self.nr if r1 is None else r1
        if out is None:
            out = np.zeros((r1 - r0, self.nc), dtype=L.ESCAPE_DTYPE)
        self._ck(self.lib.nm_read_rows(self.h, r0, r1, L.ptr(out)))
        return out

    def read_rows_pitched(self, r0, r1, dst_ptr, pitch_bytes):
        """Rows [r0, r1) to host/device address dst_ptr with a row pitch (interleaved bands of a shared raster)."""
        self._ck(self.lib.nm_read_rows_pitched(self.h, r0, r1, C.c_void_p(dst_ptr), pitch_bytes))

    def read_rows_pitched_async(self, r0, r1, dst_ptr, pitch_bytes):
        """Start the copy-out of rows [r0, r1) (snapshot on the device, copy on a second stream) and return."""
        self._ck(self.lib.nm_read_rows_pitched_async(self.h, r0, r1, C.c_void_p(dst_ptr), pitch_bytes))

    def read_wait(self):
        self._ck(self.lib.nm_read_wait(self.h))

    def host_register(self, ptr, nbytes):
        """Page-lock a host buffer; False if the OS refuses (the buffer then stays pageable)."""
        return self.lib.nm_host_register(self.h, C.c_void_p(ptr), nbytes) == L.NM_OK

    def host_unregister(self, ptr):
        self.lib.nm_host_unregister(self.h, C.c_void_p(ptr))

    def read_pixels(self, pix):
        pix = np.ascontiguousarray(pix, dtype=np.int32)
        out = np.zeros(len(pix), dtype=L.ESCAPE_DTYPE)
        self._ck(self.lib.nm_read_pixels(self.h, L.ptr(pix), len(pix), L.ptr(out)))
        return out

    def stats(self):
        s = L.Stats()
        self._ck(self.lib.nm_frame_stats(self.h, C.byref(s)))
        return s.asdict()

    def render_hw(self, c_re, c_im, N, out=None):
        if out is None:
            out = np.zeros((len(c_im), len(c_re)), dtype=L.ESCAPE_DTYPE)
        self.nr, self.nc = len(c_im), len(c_re)
        self._ck(self.lib.nm_render_hw(self.h, L.ptr(c_re), len(c_re), L.ptr(c_im), len(c_im), N, L.ptr(out)))
        return out

    def render_deep(self, tables, eps_re, eps_im, cardioid_mode=L.CARDIOID_NONE, mask=None, pix_list=None,
                    mode=L.MODE_REQUEUE, out=None):
        if out is None:
            out = np.zeros((len(eps_im), len(eps_re)), dtype=L.ESCAPE_DTYPE)
        self.nr, self.nc = len(eps_im), len(eps_re)
        n_list = 0 if pix_list is None else len(pix_list)
        self._ck(self.lib.nm_render_deep(self.h, C.byref(tables), L.ptr(eps_re), len(eps_re), L.ptr(eps_im),
                                         len(eps_im), cardioid_mode, L.ptr(mask), L.ptr(pix_list), n_list, mode,
                                         L.ptr(out)))
        return out

    def resolve(self, pal_rgb, N, sc=1, smooth=True, out=None):
        n_pal = len(pal_rgb) // 3 if pal_rgb.ndim == 1 else pal_rgb.shape[0]
        if out is None:
            out = np.zeros((self.nr // sc, self.nc // sc, 3), dtype=np.uint8)
        self._ck(self.lib.nm_resolve(self.h, L.ptr(pal_rgb), n_pal, N, sc, int(bool(smooth)), L.ptr(out)))
        return out

    def video_inbetween(self, prev_rgb, next_rgb, nr, nc, rate=45, out=None):
        """VideoZoom::nextFrame (video.cpp:14-34): the `rate` canvases between two key frames -> (rate, nr, nc, 3) uint8."""
        H, W = prev_rgb.shape[0], prev_rgb.shape[1]
        assert tuple(next_rgb.shape) == tuple(prev_rgb.shape)
        if out is None:
            out = np.zeros((rate, nr, nc, 3), dtype=np.uint8)
        self._ck(self.lib.nm_video_inbetween(self.h, L.ptr(prev_rgb), L.ptr(next_rgb), H, W, nr, nc, rate, L.ptr(out)))
        return out

    def resolve_device_palette(self, N, sc=1, smooth=True, out=None):
        """Colour resolve of the resident raster with the palette MultiWaveGenerator.cache_device() left on this GPU."""
        if out is None:
            out = np.zeros((self.nr // sc, self.nc // sc, 3), dtype=np.uint8)
        self._ck(self.lib.nm_resolve_device_palette(self.h, N, sc, int(bool(smooth)), L.ptr(out)))
        return out

    def resolve_grid(self, grid, pal_rgb, N, sc=1, smooth=True, out=None):
        nr, nc = grid.shape
        n_pal = len(pal_rgb) // 3 if pal_rgb.ndim == 1 else pal_rgb.shape[0]
        if out is None:
            out = np.zeros((nr // sc, nc // sc, 3), dtype=np.uint8)
        self._ck(self.lib.nm_resolve_grid(self.h, L.ptr(grid), nr, nc, L.ptr(pal_rgb), n_pal, N, sc, int(bool(smooth)),
                                          L.ptr(out)))
        return out
