"""Multi-rank plumbing (one process per GPU, torch.distributed): the path shards by row-interleaved
bands with NO collective on the per-pixel data path. What is exchanged:
  * tables of a reference orbit: computed on rank 0 (host GMP), broadcast            (broadcast_arrays)
  * which glitched sample becomes the next secondary reference: one MIN all-reduce     (make_reduce_pick)
  * finished raster bands: gathered to rank 0, the host-facing rank                   (gather_bands)
Backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist

_BIG = np.iinfo(np.int64).max


def make_reduce_pick(world, device):
    """reduce_pick(best_iter, best_global_pix, n_local) -> global pixel of the next reference or None.
    Rule (same as Mandelbrot::renderFrame): earliest flagged iteration, lowest pixel id on ties."""
    def reduce_pick(best_iter, best_pix, n_local):
        key = _BIG if best_iter == _BIG else (int(best_iter) << 40) | int(best_pix)
        if world > 1:
            t = torch.tensor([key], dtype=torch.int64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            key = int(t.item())
        return None if key == _BIG else key & ((1 << 40) - 1)
    return reduce_pick


def broadcast_arrays(arrays, meta, rank, world, device, sizes_from_meta):
    """rank 0 passes dict name->np.float64 array and an int64 meta list; others pass None.
    Returns (dict name->torch tensor on `device`, meta list)."""
    m = torch.tensor(meta if rank == 0 else [0] * len(meta), dtype=torch.int64, device=device)
    if world > 1:
        dist.broadcast(m, 0)
    meta = [int(x) for x in m.tolist()]
    out = {}
    for name, n in sizes_from_meta(meta):
        t = (torch.from_numpy(np.ascontiguousarray(arrays[name])).to(device) if rank == 0
             else torch.empty(n, dtype=torch.float64, device=device))
        if world > 1:
            dist.broadcast(t, 0)
        out[name] = t
    return out, meta


def gather_bands(band, nr, rank, world):
    """band: (rows_local, nc, 2) int32 tensor of this rank's interleaved rows. Returns the assembled
    (nr, nc, 2) tensor on rank 0 (None elsewhere)."""
    if world == 1:
        return band
    # rows r, r + world, ...: ceil(nr / world) rows on the first nr % world ranks, one fewer on the rest — every rank
    # pads its band to the longest one so the gather sees equal shapes, and rank 0 drops the padding
    longest = -(-nr // world)
    if band.shape[0] < longest:
        pad = torch.zeros((longest - band.shape[0],) + tuple(band.shape[1:]), dtype=band.dtype, device=band.device)
        band = torch.cat([band, pad], dim=0)
    elif band.shape[0] > longest:
        raise ValueError(f"gather_bands: band of {band.shape[0]} rows, but rank {rank} of {world} owns at most {longest} of {nr}")
    bufs = [torch.empty_like(band) for _ in range(world)] if rank == 0 else None
    dist.gather(band, bufs, dst=0)
    if rank != 0:
        return None
    full = torch.empty((nr,) + tuple(band.shape[1:]), dtype=band.dtype, device=band.device)
    for r in range(world):
        n_r = len(range(r, nr, world))
        full[r::world] = bufs[r][:n_r]
    return full


class SharedHostRaster:
    """The host-facing raster of an N-rank frame as ONE pinned buffer every rank can DMA into: a POSIX
    shared-memory file mapped by all ranks of the node and page-locked in each (nm_host_register), so
    rank r copies its interleaved rows r, r+N, ... device -> host over its own PCIe link
    (Device.read_rows_pitched) instead of funnelling the whole frame through rank 0's. Rank 0 reads the
    assembled raster after the barrier. (The NCCL alternative is gather_bands.)

    create() returns None — on every rank alike — when /dev/shm cannot hold the raster or any rank fails
    to page-lock it; the caller then uses gather_bands."""

    MAX_BYTES = 4 << 30

    @staticmethod
    def create(dev, nr, nc, rank, world, device, tag="raster"):
        import mmap
        import os
        nbytes = nr * nc * 8
        ok = [False, None]
        if rank == 0:
            try:
                st = os.statvfs("/dev/shm")
                # (page-locking one 8.5 GB mapping in each of 8 processes was refused by the OS on the B200 box)
                if st.f_bavail * st.f_frsize > nbytes * 1.25 and nbytes <= SharedHostRaster.MAX_BYTES:
                    path = f"/dev/shm/newman_b200_{os.getpid()}_{tag}"
                    with open(path, "wb") as f:
                        f.truncate(nbytes)
                    ok = [True, path]
            except OSError:
                ok = [False, None]
        if world > 1:
            dist.broadcast_object_list(ok, src=0)
        if not ok[0]:
            return None
        self = SharedHostRaster.__new__(SharedHostRaster)
        self.nr, self.nc, self.rank, self.world, self.dev, self.path = nr, nc, rank, world, dev, ok[1]
        self.nbytes = nbytes
        self._f = open(self.path, "r+b")
        self._mm = mmap.mmap(self._f.fileno(), nbytes)
        self.array = np.frombuffer(self._mm, dtype=np.int32).reshape(nr, nc, 2)
        self.ptr = self.array.ctypes.data
        self._registered = dev.host_register(self.ptr, nbytes)
        good = torch.tensor([1 if self._registered else 0], dtype=torch.int32, device=device)
        if world > 1:
            dist.all_reduce(good, op=dist.ReduceOp.MIN)
        if rank == 0:
            try:
                os.unlink(self.path)      # the mappings keep it alive; nothing is left behind in /dev/shm
            except OSError:
                pass
        if int(good.item()) == 0:
            self.close()
            return None
        return self

    def band_ptr(self):
        """Address of this rank's first row; rows are self.world * nc * 8 bytes apart."""
        return self.ptr + self.rank * self.nc * 8, self.world * self.nc * 8

    def close(self):
        if self._registered:
            self.dev.host_unregister(self.ptr)
            self._registered = False
        self.array = None
        try:
            self._mm.close()
        except BufferError:
            pass
        self._f.close()


class RenderGroup:
    """One rank of a multi-GPU render group behind the C-ABI (nmm_*, include/newman_b200.h): the round loop, the table
    broadcasts (ncclBroadcast), the 8-byte next-reference MIN (ncclAllReduce) and the band return all run inside
    libnewman_b200.so; this class only carries the 128-byte NCCL id from rank 0 to the other processes (through
    torch.distributed, which the launcher has already set up) and hands over views and output buffers.

    One process per GPU (torchrun): RenderGroup(device=LOCAL_RANK, rank=RANK, world=WORLD_SIZE).
    (Threads of one process need none of this: Mandelbrot.set_devices([...]).)"""

    RETURN_LOCAL, RETURN_ROOT = 0, 1

    def __init__(self, device, rank=0, world=1, backend_device=None):
        import ctypes as C
        from . import _lib as L
        from . import view as V
        self.C, self.L, self.V = C, L, V
        self.lib = V._lib()
        self.rank, self.world, self.device = rank, world, device
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if world > 1:
            if rank == 0:
                raw = (C.c_uint8 * 128)()
                rc = self.lib.nmm_unique_id(raw)
                if rc < 0:
                    raise L.NmError(rc, self.lib.nmm_last_error(None).decode())
                idbuf = torch.tensor(list(raw), dtype=torch.uint8)
            dev = backend_device if backend_device is not None else (
                torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu"))
            t = idbuf.to(dev)
            dist.broadcast(t, 0)
            idbuf = t.cpu()
        raw = (C.c_uint8 * 128)(*idbuf.tolist())
        h = C.c_void_p()
        rc = self.lib.nmm_create(device, rank, world, raw, C.byref(h))
        if rc < 0:
            raise L.NmError(rc, self.lib.nmm_last_error(None).decode())
        self.h = h

    def _ck(self, rc):
        if rc < 0:
            raise self.L.NmError(rc, self.lib.nmm_last_error(self.h).decode())
        return rc

    def render(self, view, band_rows, out=None, return_mode=1):
        """Collective. view: newman_b200.Mandelbrot in the same state on every rank. out: (nr, nc) ESCAPE_DTYPE / (nr, nc, 2)
        int32 host array or torch tensor (rank 0 for RETURN_ROOT; every rank for RETURN_LOCAL) or None. -> frame info dict"""
        info = self.V.FrameInfo()
        ptr = None if out is None else self.L.ptr(out)
        self._ck(self.lib.nmm_render(self.h, view.h, int(band_rows), ptr, int(return_mode), self.C.byref(info)))
        return info.asdict()

    def resolve(self, pal_rgb, sc, smooth, out=None, return_mode=1):
        pal = np.ascontiguousarray(pal_rgb, dtype=np.uint8)
        ptr = None if out is None else self.L.ptr(out)
        self._ck(self.lib.nmm_resolve(self.h, self.L.ptr(pal), len(pal.reshape(-1)) // 3, int(sc), 1 if smooth else 0, ptr,
                                      int(return_mode)))

    def exchange_ms(self):
        return float(self.lib.nmm_exchange_ms(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.nmm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
