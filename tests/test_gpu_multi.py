"""-m gpu: one frame over several GPUs behind the C-ABI (nmm_* / Mandelbrot::devices; reference viewer.cpp:193-238:
the beauty render's row loop). The N-GPU raster must be byte-identical to the 1-GPU one (last row).
On a one-GPU box the collective code path still runs with a group of one rank; the 2+-GPU tests skip."""
import os
import subprocess
import sys

import numpy as np
import pytest

import newman_b200
from newman_b200 import multigpu, workloads
from oracles import KATS

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    import torch
    return torch.cuda.device_count()


def same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


@pytest.mark.parametrize("name,band", [("KAT-1c", 1), ("KAT-D30", 2), ("KAT-S", 3), ("KAT-D90", 4), ("cfg2/8", 2), ("cfg4/32", 5)])
def test_group_of_one_equals_class_render(name, band):
    """render_collective (mandelbrot_host.cpp) with world = 1: same raster, counters and RGB as Mandelbrot::renderFrame."""
    if "/" in name:
        c = workloads.config(name.split("/")[0], scale=int(name.split("/")[1]))
        k = dict(nr=c["nr"], nc=c["nc"], N=c["N"], sz=c["sz"], center=c["center"], tol=c["tol"])
    else:
        k = dict(KATS[name])
    grp = multigpu.RenderGroup(0)
    try:
        v = newman_b200.Mandelbrot(**k)
        out = np.zeros((k["nr"], k["nc"]), dtype=newman_b200.ESCAPE_DTYPE)
        info = grp.render(v, band, out)
        single = newman_b200.Mandelbrot(**k)
        want = single.render()
        si = single.frame_info()
        assert same(out, want)
        for key in ("executed_iters", "references", "glitched", "rebased", "orbit_len", "hardware"):
            assert info[key] == si[key], key
        if k["nr"] % band == 0 and k["nc"] % band == 0:
            pal = (np.arange(3 * k["N"]) * 7 % 256).astype(np.uint8).reshape(-1, 3)
            rgb = np.zeros((k["nr"] // band, k["nc"] // band, 3), dtype=np.uint8)
            grp.resolve(pal, band, True, rgb)
            assert np.array_equal(rgb, single.resolve(pal, sc=band, smooth=True))
    finally:
        grp.close()


def test_band_rows_must_divide():
    grp = multigpu.RenderGroup(0)
    try:
        v = newman_b200.Mandelbrot(48, 64, N=64)
        with pytest.raises(newman_b200.NmError):
            grp.render(v, 5, np.zeros((48, 64), dtype=newman_b200.ESCAPE_DTYPE))
    finally:
        grp.close()


@pytest.mark.parametrize("name", ["KAT-1c", "KAT-S", "cfg2/8"])
def test_devices_threads_equal_single_gpu(name):
    """Mandelbrot::devices: host threads + NCCL ranks inside one process."""
    if n_gpus() < 2:
        pytest.skip("needs 2+ GPUs")
    if "/" in name:
        c = workloads.config(name.split("/")[0], scale=int(name.split("/")[1]))
        k = dict(nr=c["nr"], nc=c["nc"], N=c["N"], sz=c["sz"], center=c["center"], tol=c["tol"])
    else:
        k = dict(KATS[name])
    want = newman_b200.Mandelbrot(**k).render()
    v = newman_b200.Mandelbrot(**k)
    v.set_devices(list(range(n_gpus())), band_rows=2)
    got = v.render()
    assert same(got, want)
    got2 = v.render()     # the group is reused
    assert same(got2, want)


def test_process_group_equals_single_gpu():
    """One process per GPU (torchrun), the frame split inside the library: tests/multirank_nccl_worker.py."""
    n = n_gpus()
    if n < 2:
        pytest.skip("needs 2+ GPUs")
    n = min(n, 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
                        "127.0.0.1", "--master-port", "29631", os.path.join(ROOT, "tests", "multirank_nccl_worker.py")],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "MULTI-RANK OK" in r.stdout
