"""ctypes binding of libnewman_b200.so (the C-ABI in include/newman_b200.h).

The library is built in-tree by ``newman_b200/csrc/Makefile`` (``__graft_entry__.build()``).
There is deliberately no fallback: if the shared object is missing or a CUDA device is absent the
import / context creation raises — nothing in this package computes on the CPU.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NEWMAN_B200_LIB: an alternative build of the same library (kernel-variant experiments); default in-tree
LIB_PATH = os.environ.get("NEWMAN_B200_LIB") or os.path.join(_HERE, "libnewman_b200.so")

ESCAPE_DTYPE = np.dtype([("iterations", "<i4"), ("smoothing", "<f4")])  # grid.h:8-16

NM_OK, NM_EINVAL, NM_ENODEV, NM_ECUDA, NM_ENOMEM, NM_ESTATE, NM_ERANGE, NM_ECANCELLED = 0, -1, -2, -3, -4, -5, -6, -7
CARDIOID_NONE, CARDIOID_ALL, CARDIOID_MASK = 0, 1, 2
MODE_REQUEUE, MODE_REBASE, MODE_DD = 0, 1, 2
TABLES_ORBIT_TRUNCATED = 1
OPT_K2_LITERAL = 1
OPT_K3_GROUP = 2
OPT_K3_FINISH_MAX = 3
OPT_K3_SPLIT = 4


class NmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"newman_b200 error {code}: {msg}")
        self.code = code


class DeepTables(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("has_escape", C.c_int32), ("flags", C.c_int32),
        ("tol", C.c_double), ("glitch_tol", C.c_double),
        ("x_hi", C.c_void_p), ("x_lo", C.c_void_p), ("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p),
        ("a_exp", C.c_void_p), ("b_exp", C.c_void_p), ("c_exp", C.c_void_p),      # floatexp series (or NULL)
        ("eps_re_exp", C.c_void_p), ("eps_im_exp", C.c_void_p),                    # floatexp eps => scaled deltas
        ("eps_re_lo", C.c_void_p), ("eps_im_lo", C.c_void_p),                      # NM_MODE_DD: low parts of eps (or NULL)
    ]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "pixels", "executed_iters", "series_evals", "skipped_pixels", "glitched", "rebased", "fixups",
        "kernel_launches", "sweeps", "checked_steps")] + [(n, C.c_float) for n in ("ms_k1", "ms_k2", "ms_k3", "ms_k4")]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Escape(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("smoothing", C.c_float)]


# name -> (restype, argtypes); every symbol include/newman_b200.h declares
DEVICE_API = {
    "nm_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "nm_destroy": (None, [C.c_void_p]),
    "nm_last_error": (C.c_char_p, [C.c_void_p]),
    "nm_version": (C.c_char_p, []),
    "nm_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nm_get_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "nm_sync": (C.c_int, [C.c_void_p]),
    "nm_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "nm_cancel": (C.c_int, [C.c_void_p]),
    "nm_frame_hw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "nm_frame_deep": (C.c_int, [C.c_void_p, C.POINTER(DeepTables), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "nm_launch": (C.c_int, [C.c_void_p]),
    "nm_frame_ambiguous": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64]),
    "nm_frame_requeue": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "nm_poke": (C.c_int, [C.c_void_p, C.c_int64, Escape]),
    "nm_read_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "nm_read_rows_pitched": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "nm_raster_keep": (C.c_int, [C.c_void_p]),
    "nm_raster_diff": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64]),
    "nm_raster_restore": (C.c_int, [C.c_void_p]),
    "nm_read_rows_pitched_async": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "nm_read_wait": (C.c_int, [C.c_void_p]),
    "nm_host_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "nm_host_unregister": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nm_read_pixels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "nm_frame_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "nm_render_hw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "nm_render_deep": (C.c_int, [C.c_void_p, C.POINTER(DeepTables), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                 C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "nm_resolve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nm_resolve_grid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p]),
    "nm_video_inbetween": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p]),
    "nm_palette_cache": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "nm_resolve_device_palette": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nm_fp64_peak": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "nm_k3_filter_entry": (None, [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_uint32)]),
    "nm_k3_seg_bound": (C.c_int32, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_double]),
    "nm_k3_filter_fires": (C.c_int, [C.POINTER(C.c_uint32), C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int),
                                     C.POINTER(C.c_int)]),
    "nm_device_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t),
                                 C.c_char_p, C.c_int]),
}

_lib = None


def load():
    """Load libnewman_b200.so (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C newman_b200/csrc`). newman_b200 has no CPU fallback.")
    _lib = load_from(LIB_PATH)
    return _lib


def load_from(path):
    """A separate handle on another build of the library (tests compare builds side by side); not cached."""
    lib = C.CDLL(path)
    for name, (res, args) in DEVICE_API.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def ptr(a):
    """void* of a numpy array (host) or of anything exposing data_ptr() (torch tensor, host or device)."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)
