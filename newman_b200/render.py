"""Headless renderer: a saved newman location (the 5-line view file FractalViewer::save writes,
reference viewer.cpp:12-23, read by Mandelbrot::loadLegacy, mandelbrot.cpp:19-35) + a `.pal` palette
(multiwave.cpp:19-48) -> PNG, without the SDL viewer. Covers what the reference's interactive render,
screenshot (viewer.cpp:52-60) and beauty render (viewer.cpp:186-253: 1920x1080 at 3x3 multisampling)
produce, on the GPU path.

  python -m newman_b200.render --view loc.txt --palette my.pal --size 1920x1080 --sc 3 -o out.png
  python -m newman_b200.render --center -0.75 0.1 --depth 1e-3 -N 4096 -o out.png

Plumbing only: the raster comes from the drop-in class (libnewman_b200.so), the colours from K4.
"""
import argparse
import os
import struct
import sys
import time
import zlib

import numpy as np

from . import Mandelbrot, MultiWaveGenerator

DEFAULT_PAL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "default.pal")


def write_png(path, rgb):
    """Minimal 8-bit RGB PNG writer (zlib only)."""
    h, w, _ = rgb.shape
    raw = np.concatenate([np.zeros((h, 1), dtype=np.uint8), rgb.reshape(h, w * 3)], axis=1).tobytes()

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--view", help="5-line view file: N, centre.re, centre.im, sz.re, sz.im (for an 800x600 window)")
    ap.add_argument("--center", nargs=2, metavar=("RE", "IM"), help="centre as decimal strings")
    ap.add_argument("--depth", default=None, help="scale on the default view: sz = 4*depth/cols, 3*depth/rows")
    ap.add_argument("-N", type=int, default=None, help="iteration limit")
    ap.add_argument("--size", default="800x600", help="output image WxH (default: the viewer's window)")
    ap.add_argument("--sc", type=int, default=1, help="multisampling per axis (the viewer's 1-4 keys; beauty render: 3)")
    ap.add_argument("--palette", default=DEFAULT_PAL, help=".pal file (default: the reference's default.pal parameters)")
    ap.add_argument("--no-smooth", action="store_true", help="S key: smoothing off")
    ap.add_argument("--tolerance", type=float, default=1e-10, help="series error tolerance (E key)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--devices", default=None, help="comma-separated GPU ordinals: split the frame over these GPUs (Mandelbrot::devices)")
    ap.add_argument("-o", "--output", default="newman.png")
    a = ap.parse_args(argv)

    w, h = [int(x) for x in a.size.lower().split("x")]
    nr, nc = h * a.sc, w * a.sc
    m = Mandelbrot(nr, nc, device=a.device)
    if a.devices:
        m.set_devices([int(x) for x in a.devices.split(",")], band_rows=a.sc * (2 if a.sc < 4 else 1))
    if a.view:
        m.loadLegacy(a.view)             # rescales sz from the 800x600 window to this grid
        N = m.frame_N()
        if a.N:
            N = a.N
        m.set_view(N, None, None, a.tolerance)
    else:
        from fractions import Fraction
        from .workloads import _dec
        d = Fraction(a.depth) if a.depth else Fraction(1)
        sz = (_dec(4 * d / nc), _dec(3 * d / nr))
        m.set_view(a.N or 256, sz, tuple(a.center) if a.center else ("-0.5", "0"), a.tolerance)
    t0 = time.perf_counter()
    m.precompute()
    info = m.frame_info()
    pal = MultiWaveGenerator(a.palette).cache(max(m.N, 1))
    rgb = m.resolve(pal, sc=a.sc, smooth=not a.no_smooth)
    write_png(a.output, rgb)
    print(f"{a.output}: {w}x{h} (grid {nc}x{nr}), N={m.N}, {'plain double' if info['hardware'] == 1 else 'series+perturbation'}, "
          f"orbit {info['orbit_len']}, {info['executed_iters']:.3g} iterations, device {info['device_ms']:.1f} ms, "
          f"host precompute {info['host_precompute_s']:.2f} s, total {time.perf_counter() - t0:.2f} s", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
