"""world_size-2 (and 3) CPU test of the N>1 path over gloo: row-interleaved bands, rank-0 tables
broadcast, the cross-rank secondary-reference pick and the band gather must reproduce the single
process raster byte for byte. Oracle-P stands in for the device (tests may use the oracle)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kat, outdir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import newman_b200
    from newman_b200 import multigpu, pipeline
    from oracles import KATS, OracleDevice
    k = KATS[kat]
    nr, nc, N, tol = k["nr"], k["nc"], k["N"], k.get("tol", 1e-10)
    view = newman_b200.Mandelbrot(nr, nc, N=N, sz=k.get("sz"), center=k.get("center"), tol=tol) if rank == 0 else None
    dev_t = torch.device("cpu")

    def tables(row, col):
        h = view.host_tables(row, col) if rank == 0 else None
        meta = [h["M"], h["has_escape"], h["probe"][0], h["probe"][1]] if rank == 0 else [0, 0, 0, 0]
        sizes = lambda m: [("x_hi", 2 * (m[0] + m[1])), ("x_lo", 2 * m[0]), ("a", 2 * m[0]), ("b", 2 * m[0]),
                           ("c", 2 * m[0]), ("eps_re", nc), ("eps_im", nr)]
        arrs, meta = multigpu.broadcast_arrays(h, meta, rank, world, dev_t, sizes)
        d = {k2: v.numpy() for k2, v in arrs.items()}
        d.update(M=meta[0], has_escape=meta[1], probe=(meta[2], meta[3]))
        return pipeline.TableSet(d, N, tol, 1e-6)

    rows = pipeline.local_rows(nr, rank, world)
    od = OracleDevice()
    res = pipeline.render_rounds(od, tables(-1, -1), lambda gp: tables(gp // nc, gp % nc), nc, rows,
                                 reduce_pick=multigpu.make_reduce_pick(world, dev_t))
    band = torch.from_numpy(od.out.view(np.int32).reshape(len(rows), nc, 2).copy())
    full = multigpu.gather_bands(band, nr, rank, world)
    if rank == 0:
        np.save(os.path.join(outdir, "raster.npy"), full.numpy())
        np.save(os.path.join(outdir, "refs.npy"), np.array(res["refs"], dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kat", [(2, "KAT-D60"), (3, "KAT-D90"), (2, "KAT-S")])
def test_bands_equal_single_process(tmp_path, world, kat):
    sys.path.insert(0, HERE)
    import newman_b200
    from newman_b200 import pipeline
    from oracles import KATS, OracleDevice
    k = KATS[kat]
    if k["nr"] % world:
        pytest.skip("rows must divide evenly")
    nr, nc, N, tol = k["nr"], k["nc"], k["N"], k.get("tol", 1e-10)
    # single process expectation
    view = newman_b200.Mandelbrot(nr, nc, N=N, sz=k.get("sz"), center=k.get("center"), tol=tol)
    mk = lambda d: pipeline.TableSet(d, N, tol, 1e-6)
    od = OracleDevice()
    res = pipeline.render_rounds(od, mk(view.host_tables()), lambda gp: mk(view.host_tables(gp // nc, gp % nc)), nc,
                                 np.arange(nr))
    port = _free_port()
    mp.spawn(_worker, args=(world, port, kat, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "raster.npy")
    refs = np.load(tmp_path / "refs.npy").tolist()
    assert refs == res["refs"]
    assert np.array_equal(got.reshape(nr, nc, 2), od.out.view(np.int32).reshape(nr, nc, 2))
    assert (od.out["iterations"] >= 0).all()


def _banded_worker(rank, world, port, kat, band, outdir):
    """The band layout of the C++ render group (nmm_band_layout: blocks of `band` rows dealt round-robin, ragged when the
    block count does not divide) driving the same round loop: every rank renders its blocks, the cross-rank MIN picks the
    next reference by GLOBAL sample id, rank 0 assembles the raster from the layout."""
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import newman_b200
    from newman_b200 import multigpu, pipeline
    from newman_b200 import view as V
    from oracles import KATS, OracleDevice
    k = KATS[kat]
    nr, nc, N, tol = k["nr"], k["nc"], k["N"], k.get("tol", 1e-10)
    lib = V._lib()
    n = lib.nmm_band_layout(nr, band, rank, world, None)
    rows = np.zeros(max(n, 1), dtype=np.int32)
    assert lib.nmm_band_layout(nr, band, rank, world, rows.ctypes.data) == n
    rows = rows[:n].astype(np.int64)
    view = newman_b200.Mandelbrot(nr, nc, N=N, sz=k.get("sz"), center=k.get("center"), tol=tol) if rank == 0 else None
    dev_t = torch.device("cpu")

    def tables(row, col):
        h = view.host_tables(row, col) if rank == 0 else None
        meta = [h["M"], h["has_escape"], h["probe"][0], h["probe"][1]] if rank == 0 else [0, 0, 0, 0]
        sizes = lambda m: [("x_hi", 2 * (m[0] + m[1])), ("x_lo", 2 * m[0]), ("a", 2 * m[0]), ("b", 2 * m[0]),
                           ("c", 2 * m[0]), ("eps_re", nc), ("eps_im", nr)]
        arrs, meta = multigpu.broadcast_arrays(h, meta, rank, world, dev_t, sizes)
        d = {k2: v.numpy() for k2, v in arrs.items()}
        d.update(M=meta[0], has_escape=meta[1], probe=(meta[2], meta[3]))
        return pipeline.TableSet(d, N, tol, 1e-6)

    od = OracleDevice()
    res = pipeline.render_rounds(od, tables(-1, -1), lambda gp: tables(gp // nc, gp % nc), nc, rows,
                                 reduce_pick=multigpu.make_reduce_pick(world, dev_t))
    parts = [None] * world
    dist.all_gather_object(parts, (rows, od.out.copy()))
    if rank == 0:
        full = np.zeros((nr, nc), dtype=od.out.dtype)
        seen = np.zeros(nr, dtype=np.int32)
        for r_rows, r_out in parts:
            full[r_rows] = r_out
            seen[r_rows] += 1
        assert (seen == 1).all(), "the bands must partition the rows"
        for r_rows, _ in parts:     # every band starts on a multiple of `band` and is `band` rows long
            assert (r_rows.reshape(-1, band)[:, 0] % band == 0).all() and (np.diff(r_rows.reshape(-1, band), axis=1) == 1).all()
        np.save(os.path.join(outdir, "raster.npy"), full.view(np.int32).reshape(nr, nc, 2))
        np.save(os.path.join(outdir, "refs.npy"), np.array(res["refs"], dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kat,band", [(2, "KAT-S", 3), (3, "KAT-S", 3), (3, "KAT-D60", 4)])
def test_banded_layout_equals_single_process(tmp_path, world, kat, band):
    sys.path.insert(0, HERE)
    import newman_b200
    from newman_b200 import pipeline
    from oracles import KATS, OracleDevice
    k = KATS[kat]
    nr, nc, N, tol = k["nr"], k["nc"], k["N"], k.get("tol", 1e-10)
    view = newman_b200.Mandelbrot(nr, nc, N=N, sz=k.get("sz"), center=k.get("center"), tol=tol)
    mk = lambda d: pipeline.TableSet(d, N, tol, 1e-6)
    od = OracleDevice()
    res = pipeline.render_rounds(od, mk(view.host_tables()), lambda gp: mk(view.host_tables(gp // nc, gp % nc)), nc,
                                 np.arange(nr))
    port = _free_port()
    mp.spawn(_banded_worker, args=(world, port, kat, band, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "raster.npy")
    assert np.load(tmp_path / "refs.npy").tolist() == res["refs"]
    assert np.array_equal(got, od.out.view(np.int32).reshape(nr, nc, 2))
