"""Python mirror of the reference's MultiWaveGenerator (reference multiwave.h:8-37) over the nmp_* C-ABI."""
import ctypes as C

import numpy as np

from . import _lib as L

PALETTE_API = {
    "nmp_create": (C.c_void_p, []),
    "nmp_destroy": (None, [C.c_void_p]),
    "nmp_load_file": (C.c_int, [C.c_void_p, C.c_char_p]),
    "nmp_save_file": (C.c_int, [C.c_void_p, C.c_char_p]),
    "nmp_clear": (C.c_int, [C.c_void_p]),
    "nmp_add_hue_cycle": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "nmp_set_hue_period": (C.c_int, [C.c_void_p, C.c_int]),
    "nmp_set_sat_cycle": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "nmp_add_lum_wave": (C.c_int, [C.c_void_p, C.c_float, C.c_int]),
    "nmp_cache": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "nmp_cache_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
}
_bound = False


def _lib():
    global _bound
    lib = L.load()
    if not _bound:
        for name, (res, args) in PALETTE_API.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return lib


class MultiWaveGenerator:
    def __init__(self, filename=None):
        self.lib = _lib()
        self.h = C.c_void_p(self.lib.nmp_create())
        if filename:
            self.load_filename(filename)

    def __del__(self):
        try:
            self.lib.nmp_destroy(self.h)
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != L.NM_OK:
            raise L.NmError(rc, what)

    def load_filename(self, fn):
        self._ck(self.lib.nmp_load_file(self.h, str(fn).encode()), f"cannot load palette {fn}")

    def save_filename(self, fn):
        self._ck(self.lib.nmp_save_file(self.h, str(fn).encode()), f"cannot save palette {fn}")

    def clear(self):
        self.lib.nmp_clear(self.h)

    def add_hue_cycle(self, hues_deg, period):
        a = np.ascontiguousarray(hues_deg, dtype=np.float32)
        self._ck(self.lib.nmp_add_hue_cycle(self.h, L.ptr(a), len(a), int(period)), "add_hue_cycle")

    def set_hue_period(self, period):
        self._ck(self.lib.nmp_set_hue_period(self.h, int(period)), "set_hue_period")

    def set_sat_cycle(self, sats, period):
        a = np.ascontiguousarray(sats, dtype=np.float32)
        self._ck(self.lib.nmp_set_sat_cycle(self.h, L.ptr(a), len(a), int(period)), "set_sat_cycle")

    def add_lum_wave(self, amplitude, period):
        self._ck(self.lib.nmp_add_lum_wave(self.h, float(amplitude), int(period)), "add_lum_wave")

    def cache_device(self, dev, N):
        """cache(N) generated on the GPU of `dev` (K6) and kept there for dev.resolve_device_palette();
        returns a host copy (N, 3) uint8."""
        out = np.zeros((N, 3), dtype=np.uint8)
        rc = self.lib.nmp_cache_device(self.h, dev.h, int(N), L.ptr(out))
        if rc != L.NM_OK:
            raise L.NmError(rc, dev.lib.nm_last_error(dev.h).decode() or "palette is empty")
        return out

    def cache(self, N):
        """cache(N) -> (N, 3) uint8 RGB table (multiwave.cpp:75-116)."""
        out = np.zeros((N, 3), dtype=np.uint8)
        self._ck(self.lib.nmp_cache(self.h, int(N), L.ptr(out)), "palette is empty")
        return out
