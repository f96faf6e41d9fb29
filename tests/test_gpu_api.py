"""-m gpu: behaviour of the C-ABI around the kernels — argument validation, call-order errors, cancelling a
frame in flight (the viewer abandons renders on user input, viewer.cpp:177, 221-231), context reuse."""
import threading
import time

import numpy as np
import pytest

import newman_b200
import oracles
from newman_b200 import _lib as L
from newman_b200 import workloads

pytestmark = pytest.mark.gpu


def test_argument_validation(dev):
    cre, cim = np.zeros(8), np.zeros(4)
    with pytest.raises(newman_b200.NmError) as e:
        dev.frame_hw(cre, cim, -1)
    assert e.value.code == L.NM_EINVAL
    z = np.load(oracles.ROOT + "/tests/golden/kat_d30.npz")
    good = dict(x_hi=z["x_hi"], x_lo=z["x_lo"], a=z["a"], b=z["b"], c=z["c"], N=int(z["N"]), tol=float(z["tol"]))
    # orbit longer than N
    bad = dev.make_tables(good["x_hi"], good["x_lo"], good["a"], good["b"], good["c"], 10, good["tol"])
    with pytest.raises(newman_b200.NmError) as e:
        dev.frame_deep(bad, z["eps_re"], z["eps_im"])
    assert e.value.code == L.NM_EINVAL
    # floatexp eps without floatexp series tables
    ex = np.zeros(len(z["eps_re"]), dtype=np.int32), np.zeros(len(z["eps_im"]), dtype=np.int32)
    bad = dev.make_tables(good["x_hi"], good["x_lo"], good["a"], good["b"], good["c"], good["N"], good["tol"], eps_exps=ex)
    with pytest.raises(newman_b200.NmError) as e:
        dev.frame_deep(bad, z["eps_re"], z["eps_im"])
    assert e.value.code == L.NM_EINVAL
    # bad mode / bad option
    ok = dev.make_tables(good["x_hi"], good["x_lo"], good["a"], good["b"], good["c"], good["N"], good["tol"])
    with pytest.raises(newman_b200.NmError):
        dev.frame_deep(ok, z["eps_re"], z["eps_im"], mode=7)
    with pytest.raises(newman_b200.NmError):
        dev.set_option(L.OPT_K3_GROUP, 3)
    with pytest.raises(newman_b200.NmError):
        dev.set_option(99, 1)
    # the context is still usable
    out = dev.render_deep(ok, z["eps_re"], z["eps_im"])
    assert np.array_equal(out["iterations"], z["oraclep"]["iterations"])


def test_call_order_errors():
    d = newman_b200.Device(0)
    try:
        with pytest.raises(newman_b200.NmError) as e:
            d.launch()
        assert e.value.code == L.NM_ESTATE
        with pytest.raises(newman_b200.NmError):
            d.nr, d.nc = 1, 1
            d.read_rows(0, 1)
    finally:
        d.close()


def test_cancel_frame_in_flight_then_reuse():
    """A long frame is abandoned from another thread; the frame reports NM_ECANCELLED promptly and the
    next frame on the same context renders normally (persistent CTAs poll the flag)."""
    cfg = workloads.config("cfg2", scale=4)
    view = newman_b200.Mandelbrot(cfg["nr"], cfg["nc"], N=cfg["N"], sz=cfg["sz"], center=cfg["center"], tol=cfg["tol"])
    h = view.host_tables(cfg["nr"] // 2, cfg["nc"] // 2)
    d = newman_b200.Device(0)
    try:
        tabs = d.make_tables(h["x_hi"], h["x_lo"], h["a"], h["b"], h["c"], cfg["N"], cfg["tol"])
        d.frame_deep(tabs, h["eps_re"], h["eps_im"])
        t0 = time.perf_counter()
        d.launch()
        d.stats()
        full_s = time.perf_counter() - t0
        ref = d.read_rows()
        # same frame again, cancelled right away from a second thread
        d.frame_deep(tabs, h["eps_re"], h["eps_im"])
        killer = threading.Thread(target=lambda: d.cancel())
        killer.start()
        try:
            d.launch()
            d.stats()
            cancelled = False
        except newman_b200.NmError as e:
            cancelled = e.code == L.NM_ECANCELLED
        killer.join()
        # (the cancel may land after the last level was queued on a fast frame; either way the context must recover)
        d.frame_deep(tabs, h["eps_re"], h["eps_im"])
        d.launch()
        again = d.read_rows()
        assert np.array_equal(again["iterations"], ref["iterations"])
        print("frame", round(full_s * 1e3, 1), "ms; cancelled:", cancelled)
    finally:
        d.close()


def test_class_copy_semantics_and_shared_engine():
    """Python twin of what viewer.cpp does with the value type (copies at 193, 251): two views on one
    process render independently and do not disturb each other's rasters."""
    a = newman_b200.Mandelbrot(48, 64, N=256)
    b = newman_b200.Mandelbrot(30, 40, N=300, sz=("1e-18", "1e-18"),
                               center=("-0.743643887037158704752191506114774", "0.131825904205311970493132056385139"))
    ga = a.render()
    gb = b.render()
    ga2 = a.grid()
    assert np.array_equal(ga["iterations"], ga2["iterations"])
    assert gb.shape == (30, 40) and (gb["iterations"] >= 0).all()


def test_async_copy_out_overlaps_and_matches(dev):
    """nm_read_rows_pitched_async / nm_read_wait: the raster leaves from a device-side snapshot on a second stream, so a
    new frame may be started before the copy has landed; what lands is the frame the call was made for."""
    import torch
    z = np.load(oracles.ROOT + "/tests/golden/kat_d30.npz")
    t = dev.make_tables(z["x_hi"], z["x_lo"], z["a"], z["b"], z["c"], int(z["N"]), float(z["tol"]))
    nr, nc = len(z["eps_im"]), len(z["eps_re"])
    want_a = dev.render_deep(t, z["eps_re"], z["eps_im"]).copy()
    host_a = torch.zeros((nr, nc, 2), dtype=torch.int32).pin_memory()
    host_b = torch.zeros((nr, 2 * nc, 2), dtype=torch.int32).pin_memory()     # pitched destination: every other column block
    dev.frame_deep(t, z["eps_re"], z["eps_im"])
    dev.launch()
    dev.read_rows_pitched_async(0, nr, host_a.data_ptr(), nc * 8)
    # a different frame right away (rebasing mode: other records for the glitched samples), read with a pitch
    dev.frame_deep(t, z["eps_re"], z["eps_im"], mode=L.MODE_REBASE)
    dev.launch()
    want_b = dev.read_rows().copy()
    dev.read_rows_pitched_async(0, nr, host_b.data_ptr(), 2 * nc * 8)
    dev.read_wait()
    got_a = host_a.numpy().view(newman_b200.ESCAPE_DTYPE).reshape(nr, nc)
    got_b = host_b.numpy()[:, :nc].copy().view(newman_b200.ESCAPE_DTYPE).reshape(nr, nc)
    assert np.array_equal(got_a.view(np.uint8), want_a.view(np.uint8))
    assert np.array_equal(got_b.view(np.uint8), want_b.view(np.uint8))
    assert not host_b.numpy()[:, nc:].any()
