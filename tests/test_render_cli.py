"""newman_b200/render.py — the headless renderer of saved locations (reference viewer.cpp:12-23,
52-60, 186-253; mandelbrot.cpp:19-35). CPU: the PNG writer and the no-GPU failure; GPU: the PNG holds exactly the
pixels the drop-in class + K4 produce for the saved view."""
import os
import struct
import zlib

import numpy as np
import pytest

import newman_b200
from newman_b200 import render, workloads
from newman_b200 import palette as PAL


def read_png(path):
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(b):
        n, tag = struct.unpack(">I", b[pos:pos + 4])[0], b[pos + 4:pos + 8]
        data = b[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", b[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + data) & 0xFFFFFFFF
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", data[:10])
            assert (depth, ctype) == (8, 2)
        elif tag == b"IDAT":
            idat += data
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert not raw[:, 0].any()      # filter type 0 on every row
    return raw[:, 1:].reshape(h, w, 3)


def test_png_writer_round_trip(tmp_path):
    rng = np.random.default_rng(7)
    rgb = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    fn = str(tmp_path / "x.png")
    render.write_png(fn, rgb)
    assert np.array_equal(read_png(fn), rgb)


def test_cli_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(newman_b200.NmError):
        render.main(["--size", "16x12", "-N", "32", "-o", str(tmp_path / "x.png")])


@pytest.mark.gpu
@pytest.mark.parametrize("sc", [1, 3])
def test_cli_renders_a_saved_view(tmp_path, sc):
    cfg = workloads.config("cfg2", scale=40)
    src = newman_b200.Mandelbrot(600, 800, N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
    view_fn, png = str(tmp_path / "loc.txt"), str(tmp_path / "out.png")
    src.save(view_fn)                                            # what the viewer's F2 writes
    assert render.main(["--view", view_fn, "--size", "96x54", "--sc", str(sc), "-o", png]) == 0
    m = newman_b200.Mandelbrot(54 * sc, 96 * sc)
    m.loadLegacy(view_fn)
    m.set_view(m.frame_N(), None, None, 1e-10)
    m.precompute()
    pal = PAL.MultiWaveGenerator(os.path.join(os.path.dirname(render.__file__), "default.pal")).cache(m.N)
    assert np.array_equal(read_png(png), m.resolve(pal, sc=sc, smooth=True))
    # the same frame split over a render group (one rank on a one-GPU box)
    png2 = str(tmp_path / "out2.png")
    assert render.main(["--view", view_fn, "--size", "96x54", "--sc", str(sc), "--devices", "0", "-o", png2]) == 0
    assert np.array_equal(read_png(png2), read_png(png))
