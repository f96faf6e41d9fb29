"""Multiwave palette (reference multiwave.cpp): .pal I/O and cache(N) against the C restatement.
libbyteimage's hsl2rgb/interp are not vendored by the reference, so RGB parity with the original
binary is UNPINNED; what is pinned here is our C++ == our oracle, plus structural properties."""
import ctypes as C
import os

import numpy as np
import pytest

import newman_b200
import oracles

HERE = os.path.dirname(os.path.abspath(__file__))
PAL = os.path.join(HERE, "golden", "default.pal")  # the reference's default palette parameters (default.pal:1-13)


def parse_pal(fn):
    tok = open(fn).read().split()
    it = iter(tok)
    n = int(next(it))
    cycles = []
    for _ in range(n):
        m = int(next(it))
        vals = [float(next(it)) for _ in range(m)]
        cycles.append((vals, int(next(it))))
    hue_period = int(next(it))
    m = int(next(it))
    sat = [float(next(it)) for _ in range(m)]
    sat_period = int(next(it))
    m = int(next(it))
    lum = [(float(next(it)), int(next(it))) for _ in range(m)]
    return cycles, hue_period, sat, sat_period, lum


def oracle_cache(cycles, hue_period, sat, sat_period, lum, N):
    P = oracles.oraclep()
    counts = np.array([len(v) for v, _ in cycles], dtype=np.int32)
    values = np.array([x for v, _ in cycles for x in v], dtype=np.float32)
    periods = np.array([p for _, p in cycles], dtype=np.int32)
    satv = np.array(sat, dtype=np.float32)
    amp = np.array([a for a, _ in lum], dtype=np.float32)
    lper = np.array([p for _, p in lum], dtype=np.int32)
    out = np.zeros((N, 3), dtype=np.uint8)
    P.oraclep_palette_cache.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    P.oraclep_palette_cache(len(cycles), oracles.vp(counts), oracles.vp(values), oracles.vp(periods), hue_period,
                            len(sat), oracles.vp(satv), sat_period, len(lum), oracles.vp(amp), oracles.vp(lper), N,
                            oracles.vp(out))
    return out


def test_default_palette_matches_restatement():
    mw = newman_b200.MultiWaveGenerator(PAL)
    spec = parse_pal(PAL)
    for N in (1, 256, 4096, 65536):
        got = mw.cache(N)
        want = oracle_cache(*spec, N)
        assert np.array_equal(got, want), N
    pal = mw.cache(4096)
    assert len(np.unique(pal, axis=0)) > 500          # a real gradient, not a constant
    assert pal.min() >= 0 and pal.max() <= 255


def test_pal_roundtrip_and_builder(tmp_path):
    mw = newman_b200.MultiWaveGenerator(PAL)
    fn = str(tmp_path / "out.pal")
    mw.save_filename(fn)
    assert parse_pal(fn) == parse_pal(PAL)             # %f formatting keeps every default.pal digit
    mw2 = newman_b200.MultiWaveGenerator(fn)
    assert np.array_equal(mw2.cache(1000), mw.cache(1000))
    # same palette assembled through the builder calls
    cycles, hue_period, sat, sat_period, lum = parse_pal(PAL)
    mw3 = newman_b200.MultiWaveGenerator()
    for vals, per in cycles:
        mw3.add_hue_cycle(vals, per)
    mw3.set_hue_period(hue_period)
    mw3.set_sat_cycle(sat, sat_period)
    for a, p in lum:
        mw3.add_lum_wave(a, p)
    assert np.array_equal(mw3.cache(1000), mw.cache(1000))


def test_hue_extremes():
    """Pure hue nodes at full saturation, lum wave sum 0 -> logistic(0) = 0.5: primary colours."""
    for hue, rgb in ((0.0, (255, 0, 0)), (120.0, (0, 255, 0)), (240.0, (0, 0, 255))):
        mw = newman_b200.MultiWaveGenerator()
        mw.add_hue_cycle([hue], 10)
        mw.set_hue_period(10)
        mw.set_sat_cycle([1.0], 10)
        pal = mw.cache(4)
        assert (pal == np.array(rgb, dtype=np.uint8)).all(), (hue, pal)
    mw = newman_b200.MultiWaveGenerator()
    try:
        mw.cache(4)
        assert False, "empty palette must be refused"
    except newman_b200.NmError:
        pass


@pytest.mark.gpu
def test_device_palette_k6_vs_host_builder(dev):
    """K6 (MultiWaveGenerator::cache as a kernel) against the host builder: <= 1 LSB per
    channel (north-star tolerance; device libm), nearly always equal; and the recolour of a resident
    raster from the device table equals the resolve with that table passed from the host."""
    import newman_b200
    mw = newman_b200.MultiWaveGenerator(os.path.join(HERE, "golden", "default.pal"))
    for N in (1, 256, 65536):
        host = mw.cache(N)
        gpu = mw.cache_device(dev, N)
        diff = np.abs(host.astype(np.int16) - gpu.astype(np.int16))
        assert diff.max() <= 1, (N, int(diff.max()))
        print("N", N, "entries differing by 1 LSB:", int((diff.max(axis=1) > 0).sum()))
        assert (diff.max(axis=1) > 0).mean() < 0.01
    # a second generator with several hue cycles / waves built through the API
    g = newman_b200.MultiWaveGenerator()
    g.add_hue_cycle([0.0, 120.0, 240.0], 97)
    g.add_hue_cycle([30.0, 300.0], 41)
    g.add_hue_cycle([200.0], 13)
    g.set_hue_period(1000)
    g.set_sat_cycle([1.0, 0.2, 0.7], 333)
    g.add_lum_wave(0.9, 77)
    g.add_lum_wave(0.4, 11)
    g.add_lum_wave(1.3, 1000)
    host, gpu = g.cache(5000), g.cache_device(dev, 5000)
    assert np.abs(host.astype(np.int16) - gpu.astype(np.int16)).max() <= 1
    # recolour without re-rendering
    m = newman_b200.Mandelbrot(48, 64, N=256)
    grid = m.render()
    cre, cim = m.host_coords()
    dev.render_hw(cre, cim, 256)
    table = mw.cache_device(dev, 256)
    a = dev.resolve_device_palette(256, sc=2, smooth=True)
    b = dev.resolve(table, 256, sc=2, smooth=True)
    assert np.array_equal(a, b)
