"""newman_b200 — B200-native (sm_100a CUDA) implementation of axnjaxn/newman's per-pixel
Mandelbrot hot path. The product is libnewman_b200.so (C-ABI: include/newman_b200.h, C++ drop-in
class: include/newman_b200/mandelbrot.h); this package is the Python binding used by the tests and
bench. Importing it does not load CUDA; creating a Device or View does, and raises without a GPU."""
from ._lib import (CARDIOID_ALL, CARDIOID_MASK, CARDIOID_NONE, ESCAPE_DTYPE, MODE_REBASE, MODE_REQUEUE, NmError,
                   load)
from .device import Device
from .view import Mandelbrot
from .palette import MultiWaveGenerator

__all__ = ["Device", "Mandelbrot", "MultiWaveGenerator", "NmError", "load", "ESCAPE_DTYPE", "CARDIOID_NONE", "CARDIOID_ALL", "CARDIOID_MASK",
           "MODE_REQUEUE", "MODE_REBASE"]
