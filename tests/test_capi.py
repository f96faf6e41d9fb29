"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/newman_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import newman_b200
from newman_b200 import _lib as L
from newman_b200 import view as V
from newman_b200 import palette as PAL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "newman_b200.h")).read()
    return sorted(set(re.findall(r"NM_API[^;(]*?\b(nm[vpm]?_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = C.CDLL(L.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/newman_b200.h but not exported"


def test_bindings_cover_header():
    names = set(declared_symbols())
    bound = set(L.DEVICE_API) | set(V.VIEW_API) | set(PAL.PALETTE_API)
    assert names == bound, (names - bound, bound - names)


def test_version_and_no_cpu_fallback():
    lib = newman_b200.load()
    assert b"sm_100a" in lib.nm_version()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the refusal path is exercised on the CPU box")
    with pytest.raises(newman_b200.NmError) as ei:
        newman_b200.Device(0)
    assert ei.value.code == L.NM_ENODEV and "no CPU fallback" in str(ei.value)
    m = newman_b200.Mandelbrot(8, 8, N=16)
    with pytest.raises(newman_b200.NmError) as ei:
        m.precompute()
    assert ei.value.code == L.NM_ENODEV


def test_product_does_not_reference_oracle():
    """Nothing under newman_b200/ or include/ may include, link or load oracle/ code."""
    bad = []
    for base in ("newman_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            if "_build" in dp or "__pycache__" in dp:
                continue
            for fn in fns:
                if fn.endswith((".so", ".o", ".pyc")):
                    continue
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r'(#include\s*[<"][^>"]*oracle|liboracle_p|libnewman_ref|import\s+oracles|from\s+oracles|-loracle)', txt):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad
