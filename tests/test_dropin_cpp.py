"""The drop-in at the C++ source level: tests/dropin/headless_viewer.cpp is a caller written the way
the reference's viewer uses `class Mandelbrot` / `MultiWaveGenerator` (viewer.cpp:71-124, 157-253, 271-285, 329-354).
It is compiled against include/newman_b200/ and linked with libnewman_b200.so only — which proves the headers are
self-sufficient and the library exports the classes — and what it computes must equal the C-ABI path bit for bit."""
import os
import subprocess

import numpy as np
import pytest

import newman_b200
from newman_b200 import _lib as L
from newman_b200 import palette as PAL
from newman_b200 import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "newman_b200")


def fnv(data, h=0xcbf29ce484222325):
    for b in bytes(data):
        h = ((h ^ b) * 0x100000001b3) & 0xffffffffffffffff
    return h


def build_caller(tmp_path_factory, name):
    exe = str(tmp_path_factory.mktemp("dropin") / name)
    cmd = ["g++", "-std=c++11", "-O2", "-Wall", "-Werror",
           "-I" + os.path.join(ROOT, "include", "newman_b200"),       # "mandelbrot.h" / "video.h" resolve to the drop-in
           "-I" + os.path.join(PKG, "csrc", "compat"),               # <gmpxx.h> stand-in: this image has no GMP headers
           os.path.join(ROOT, "tests", "dropin", name + ".cpp"), "-o", exe,
           "-L" + os.path.dirname(L.LIB_PATH), "-l:" + os.path.basename(L.LIB_PATH),
           "-Wl,-rpath," + os.path.dirname(L.LIB_PATH), "-l:libgmp.so.10"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.fixture(scope="module")
def viewer(tmp_path_factory):
    return build_caller(tmp_path_factory, "headless_viewer")


@pytest.fixture(scope="module")
def video(tmp_path_factory):
    return build_caller(tmp_path_factory, "headless_video")


def run(exe, *args):
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, cwd=PKG, timeout=600)  # default.pal lives in PKG
    assert r.returncode == 0, r.stderr
    out = {}
    for line in r.stdout.splitlines():
        tag, *kv = line.split()
        out[tag] = dict(x.split("=", 1) for x in kv) if kv and "=" in kv[0] else kv
    return out


def view_of(m):
    c_re, c_im, s_re, s_im = m.view_strings()
    return {"rows": str(m.rows()), "cols": str(m.cols()), "N": str(m.frame_N()), "hw": str(int(m.useHardware())),
            "c_re": c_re, "c_im": c_im, "sz_re": s_re, "sz_im": s_im}


def test_cpp_caller_links_and_view_state_matches_capi(viewer):
    got = run(viewer, "state")
    m = newman_b200.Mandelbrot(48, 64)
    assert got["reset"] == view_of(m)
    pal = PAL.MultiWaveGenerator(os.path.join(PKG, "default.pal")).cache(256)
    assert got["pal"] == {"n": "256", "hash": "%016x" % fnv(pal.tobytes())}
    m.zoomAt(2.0, 10, 40)
    assert got["zoomAt"] == view_of(m)
    m.translate(3, -5)
    assert got["translate"] == view_of(m)
    for _ in range(100):
        m.zoom(1.5)
    assert got["zoom100"] == view_of(m) and got["zoom100"]["hw"] == "0"
    assert got["restored"] == got["zoom100"] and got["scaleDown"] == got["zoom100"]   # copies are values
    assert got["beauty"]["rows"] == "96" and got["beauty"]["c_re"] == got["zoom100"]["c_re"]
    m.scaleUp(2)
    assert got["scaleUp"] == view_of(m)
    # no CPU fallback behind the C++ class either
    import torch
    assert got["precompute"] == ["ok" if torch.cuda.is_available() else "runtime_error"]


def test_library_exports_the_dropin_classes():
    out = subprocess.run("nm -D --defined-only %s | c++filt" % L.LIB_PATH, shell=True, capture_output=True, text=True).stdout
    for sym in ("Mandelbrot::Mandelbrot(int, int)", "Mandelbrot::precompute()", "Mandelbrot::computeRow(int)",
                "Mandelbrot::at(int, int)", "Mandelbrot::at(int, int, int)", "Mandelbrot::useHardware()",
                "Mandelbrot::pointAt(int, int, int) const", "Mandelbrot::translate(int, int, int)", "Mandelbrot::zoom(float)",
                "Mandelbrot::zoomAt(float, int, int, int)", "Mandelbrot::scaleUp(int)", "Mandelbrot::scaleDown(int)",
                "Mandelbrot::loadLegacy(char const*)", "MultiWaveGenerator::cache(int) const",
                "MultiWaveGenerator::load_filename(char const*)", "MultiWaveGenerator::save_filename(char const*) const",
                "VideoZoom::VideoZoom()", "VideoZoom::nextFrame(byteimage::ByteImage const&)"):
        assert sym in out, f"{sym} is not exported: a C++ caller of the reference could not link"


@pytest.mark.gpu
def test_cpp_caller_renders_the_same_frames_as_the_capi(viewer, tmp_path):
    cfg = workloads.config("cfg2", scale=40)
    src = newman_b200.Mandelbrot(600, 800, N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
    fn = str(tmp_path / "loc.txt")
    src.save(fn)                                   # the viewer's F2 file; loadLegacy rescales sz by 800 / cols
    h, w = cfg["nr"], cfg["nc"]
    got = run(viewer, "render", h, w, 256, fn)

    m = newman_b200.Mandelbrot(h, w, N=256)
    g0 = m.render()
    assert got["frame0"]["hw"] == "1" and got["frame0"]["raster"] == "%016x" % fnv(g0.tobytes())
    gen = PAL.MultiWaveGenerator(os.path.join(PKG, "default.pal"))
    assert got["frame0"]["rgb"] == "%016x" % fnv(m.resolve(gen.cache(256), sc=1, smooth=True).tobytes())

    m.loadLegacy(fn)
    assert got["loaded"] == view_of(m)
    g1 = m.render()
    assert got["frame1"]["hw"] == "0" and got["frame1"]["raster"] == "%016x" % fnv(g1.tobytes())
    assert int(got["frame1"]["M"]) == m.frame_info()["orbit_len"]
    pal = gen.cache(m.frame_N())
    assert got["frame1"]["rgb"] == "%016x" % fnv(m.resolve(pal, sc=1, smooth=True).tobytes())

    m.scaleUp(2)
    g2 = m.render()
    assert got["frame2"]["rows"] == str(2 * h) and got["frame2"]["raster"] == "%016x" % fnv(g2.tobytes())
    assert got["frame2"]["rgb"] == "%016x" % fnv(m.resolve(pal, sc=2, smooth=True).tobytes())
    assert got["k4"] == {"equal": "1"} and got["copy"] == {"at": "1"}
    assert got["rerender"]["over"] == "0" and int(got["rerender"]["N"]) == m.frame_N() // 2
    assert (g1["iterations"] < m.frame_N()).any() and (g1["iterations"] > 256).any()   # the deep frame is not trivial


@pytest.mark.gpu
def test_cpp_beauty_render_on_all_gpus(viewer, tmp_path):
    """FractalViewer::beautyRender (viewer.cpp:186-253) written against the drop-in, with `mandel.devices` = every GPU of
    the box: the frame must be identical, raster and RGB, to the one-GPU frame. (On a one-GPU box the group has one
    rank: the collective path still runs.)"""
    import torch
    n = max(1, torch.cuda.device_count())
    cfg = workloads.config("cfg2", scale=40)
    src = newman_b200.Mandelbrot(600, 800, N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
    fn = str(tmp_path / "loc.txt")
    src.save(fn)
    r = subprocess.run([viewer, "beauty", "54", "96", "3", str(cfg["N"]), str(n), fn], capture_output=True, text=True, cwd=PKG,
                       timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "beauty identical=1" in r.stdout
    lines = [l for l in r.stdout.splitlines() if l.startswith("beauty pass=")]
    assert len(lines) == 2 and f"gpus={n}" in lines[1] and "rows=162 cols=288" in lines[1]


def key_frame(k, H, W):
    r, c, ch = np.meshgrid(np.arange(H), np.arange(W), np.arange(3), indexing="ij")
    return ((7 * r + 13 * c + 29 * ch + 101 * k + r * c * k) & 255).astype(np.uint8)


def test_cpp_video_caller_links_and_needs_a_gpu_only_for_frames(video, tmp_path):
    """VideoZoom (reference video.h:13-26) from the library: the first key frame is only stored (video.cpp:15, 33); the
    in-between frames of a pair come from K5, so without a device the second nextFrame fails loudly."""
    out = str(tmp_path / "one.raw")
    r = subprocess.run([video, out, "40", "64", "5", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "ok" and os.path.getsize(out) == 0
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([video, out, "40", "64", "5", "2"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 2 and "runtime_error" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_video_caller_writes_the_frames_of_k5(video, tmp_path, dev):
    nr, nc, rate, keys = 40, 64, 5, 3
    out = str(tmp_path / "zoom.raw")
    r = subprocess.run([video, out, str(nr), str(nc), str(rate), str(keys)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(out, dtype=np.uint8)
    assert got.size == (keys - 1) * rate * nr * nc * 3
    got = got.reshape(keys - 1, rate, nr, nc, 3)
    H, W = nr * 3 // 2, nc * 3 // 2
    for k in range(1, keys):
        want = dev.video_inbetween(key_frame(k - 1, H, W), key_frame(k, H, W), nr, nc, rate)
        assert np.array_equal(got[k - 1], want), k
